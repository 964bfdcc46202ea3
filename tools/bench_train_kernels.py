#!/usr/bin/env python
"""Stand-alone timings of the training-path kernels at BASELINE config 4 shapes (12 frames of 480x720 -> 60x90 cells):
cl_gn_backward (both passes), cl_conv_wgrad_pf and the data gradient through cl_conv_igemm.  One JSON line.

    python tools/bench_train_kernels.py [batch]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from crossloc_b200 import _lib  # noqa: E402
from crossloc_b200.cnn import _Geometry, _PF  # noqa: E402
from crossloc_b200.train_plan import TrainPlan, _Src, _TrainPack  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    geo = _Geometry(batch, 60, 90)
    plan = TrainPlan.__new__(TrainPlan)
    plan._pool, plan._zero_bias, plan.bwd_terms = {}, {}, 3
    out = {'batch': batch}
    for c in (512, 256):
        norm = torch.nn.GroupNorm(32, c).to(dev)
        raw = torch.randn(geo.Mp, c, device=dev)
        stats = torch.stack([torch.zeros(batch, 32, device=dev, dtype=torch.float64),
                             torch.full((batch, 32), float(c // 32 * 5400), device=dev, dtype=torch.float64)], -1).contiguous()
        g1, g2 = torch.randn(geo.Mp, c, device=dev) * 1e-3, torch.randn(geo.Mp, c, device=dev) * 1e-3
        mask = torch.randn(geo.Mp, c, device=dev).half()
        rec = {'raw': raw, 'stats': stats}
        bytes_plain = geo.B * 5400 * c * 4 * 5.0          # g, raw (pass 0) + g, raw, d_raw hi/lo (pass 1)
        ms = timed(lambda: plan._gn_backward(lib, stream, geo, c, rec, norm, True, [_Src(g1, c)], None, False))
        out['gn_bwd_plain_%d' % c] = {'ms': ms, 'GBps': bytes_plain / ms / 1e6}
        bytes_merge = geo.B * 5400 * c * (4 * 3 + 2 + 4 + 4 * 3)   # 2 g + raw + mask + g_out; g_out + raw + d_raw
        ms = timed(lambda: plan._gn_backward(lib, stream, geo, c, rec, norm, True, [_Src(g1, c), _Src(g2, c)], mask, True))
        out['gn_bwd_merge_%d' % c] = {'ms': ms, 'GBps': bytes_merge / ms / 1e6}
    for (cin, cout, k) in ((512, 512, 3), (512, 512, 1), (256, 256, 3)):
        conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2).to(dev)
        pack = _TrainPack(conv, 4)
        act = _PF(geo, cin, 1, 2, dev)
        act.h16.normal_()
        d_raw = torch.randn(2 * geo.Mp, cout, device=dev).half()
        scale_out = torch.ones(2, device=dev)
        taps = [0] if k == 1 else [(kh - 1) * geo.Wp + (kw - 1) for kh in range(3) for kw in range(3)]
        rec = {'pack': pack, 'geo': geo, 'act': act, 'taps': taps}
        gw = torch.zeros_like(conv.weight)
        flops = 2.0 * batch * 5400 * cin * cout * k * k
        for terms in (3, 1):
            plan.bwd_terms = terms
            ms = timed(lambda: plan._conv_backward(lib, stream, rec, d_raw, scale_out, False, gw))
            out['wgrad_%dx%d_k%d_terms%d' % (cin, cout, k, terms)] = {'ms': ms, 'TFLOPs_useful': flops / ms / 1e9}
            ms_both = timed(lambda: plan._conv_backward(lib, stream, rec, d_raw, scale_out, True, gw))
            out['dgrad_%dx%d_k%d_terms%d' % (cin, cout, k, terms)] = {'ms': ms_both - ms, 'TFLOPs_useful': flops / (ms_both - ms) / 1e9}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
