#!/usr/bin/env python
"""Stand-alone timings of cl_conv_igemm at the bench shapes (32 frames, 60x90 cells), with and without the GroupNorm
statistics of the epilogue.  One line per configuration.

    python tools/bench_conv.py [batch]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from crossloc_b200 import _lib  # noqa: E402
from crossloc_b200.cnn import PackedConv, _Geometry, _PF, _taps  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    cases = ((512, 512, 1, 2, 1, 60, 90), (512, 512, 1, 3, 1, 60, 90), (512, 512, 3, 2, 1, 60, 90), (256, 256, 3, 2, 1, 60, 90),
             (256, 512, 1, 3, 1, 60, 90), (32, 64, 3, 3, 2, 240, 360), (64, 128, 3, 3, 2, 120, 180), (128, 256, 3, 3, 2, 60, 90))
    for (cin, cout, k, nterms, stride, ho, wo) in cases:
        geo = _Geometry(batch, ho, wo)
        conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).to(dev)
        pack = PackedConv(conv.weight, conv.bias, stride, nterms)
        act = _PF(geo, cin, 4 if stride == 2 else 1, 2, dev)
        act.h16.normal_()
        f8 = act.f8 if nterms == 2 else None
        raw = torch.empty(geo.Mp, cout, dtype=torch.float32, device=dev)
        stats = torch.zeros(batch, 32, 2, dtype=torch.float64, device=dev)
        taps = _taps(pack, geo)
        arr = (ctypes.c_int32 * len(taps))(*taps)
        for group_ch in (cout // 32, 0):
            def run():
                _lib.check(lib.cl_conv_igemm(act.h16.data_ptr(), act.h16.size(0), act.phases * geo.Mp, cin, pack.weights.data_ptr(), cout,
                                             len(taps), arr, nterms, geo.Mp, geo.Hp, geo.Wp, group_ch, pack.out_scale,
                                             raw.data_ptr(), pack.bias.data_ptr(), stats.data_ptr() if group_ch else 0,
                                             f8.data_ptr() if nterms == 2 else 0, f8.size(0) if nterms == 2 else 0,
                                             geo.Mp if nterms == 2 else 0, pack.weights8.data_ptr() if nterms == 2 else 0,
                                             stream))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            flops = 2.0 * batch * ho * wo * cin * cout * k * k
            print('%dx%d k%d s%d nterms%d group_ch %2d: %.4f ms  %.0f TFLOP/s useful' % (cin, cout, k, stride, nterms, group_ch, ms, flops / ms / 1e9), flush=True)


if __name__ == '__main__':
    main()
