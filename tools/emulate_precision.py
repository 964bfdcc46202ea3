#!/usr/bin/env python
"""CPU emulation of the convolution arithmetic schemes: relative error of the regressed coordinate map against an fp64
evaluation of the same network, with every convolution's operands rounded the way a tensor-core scheme would round them
(products and sums themselves are exact here; the hardware accumulates in fp32, which is far below these errors).

    python tools/emulate_precision.py [height] [width] [seeds]

Schemes ("Precision decision"):
  fp16x1      a_hi * w_hi                                          one fp16 pass (TF32 has the same 10-bit mantissa)
  fp16x3      a_hi*w_hi + a_lo*w_hi + a_hi*w_lo                    three fp16 passes
  fp16+fp8    a_hi*w_hi + 2^-14 (e4m3(a_lo 2^14) e4m3(w_hi) + e4m3(a_hi 2^2) e4m3(w_lo 2^12))   the shipped default
  fp16+fp4    the same corrections in e2m1 with one power-of-two scale per 32 input channels (block-scaled MX FP4): what
              section 9 item 1 proposes -- 1.5 instead of 2 fp16-MMA equivalents per product
  fp16+fp8a / fp16+fp8w   only the activation / only the weight correction term (1.5 equivalents)
The e4m3 / e2m1 schemes apply to the layers the engine runs in the fp16 + fp8 scheme (3x3 stride 1 with Cin >= 256 and
1x1 with Cin >= 512); every other convolution is fp16x3, as on the GPU.  One JSON line.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import networks.networks as nets  # noqa: E402
from crossloc_b200.cnn import _nterms_for  # noqa: E402


def h16(x):
    return x.to(torch.float16).to(torch.float64)


def e4m3(x):
    return x.clamp(-448.0, 448.0).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)


_E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)


def e2m1_blocks(x, dim):
    """Block-scaled FP4: along `dim` (the input-channel axis) every 32 values share a power-of-two scale (UE8M0) chosen so
    that the block maximum lands in (3, 6]; values round to the nearest e2m1 magnitude."""
    x = x.movedim(dim, -1)
    shape = x.shape
    pad = (-shape[-1]) % 32
    if pad:
        x = F.pad(x, (0, pad))
    blk = x.reshape(*x.shape[:-1], -1, 32)
    amax = blk.abs().amax(-1, keepdim=True)
    scale = torch.exp2(torch.ceil(torch.log2(amax.clamp_min(1e-300) / 6.0)))
    scale = torch.where(amax > 0, scale, torch.ones_like(scale))
    v = (blk / scale).abs().clamp(max=6.0)
    idx = torch.bucketize(v.contiguous(), (_E2M1[:-1] + _E2M1[1:]) / 2)      # nearest grid point
    q = _E2M1[idx] * torch.sign(blk) * scale
    q = q.reshape(*x.shape)[..., :shape[-1]]
    return q.movedim(-1, dim)


def make_conv(scheme):
    def conv(m, x):
        k, stride = m.kernel_size[0], m.stride[0]
        a = x.to(torch.float64)
        w = m.weight.detach().to(torch.float64)
        b = None if m.bias is None else m.bias.detach().to(torch.float64)

        def cv(aa, ww):
            return F.conv2d(aa, ww, None, stride, k // 2)
        if scheme == 'fp64' or m.in_channels % 32 != 0 or m.out_channels < 32:      # stem / head: as on the GPU (fp16x3 / fp32)
            y = cv(a, w)
        else:
            amax = float(w.abs().max())
            sc = 2.0 ** int(torch.floor(torch.log2(torch.tensor(128.0 / amax)))) if amax > 0 else 1.0
            ws = w * sc
            ah, wh = h16(a), h16(ws)
            al, wl = a - ah, ws - wh
            big = _nterms_for('fp16+fp8', m.in_channels, k, stride) == 2
            if scheme == 'fp16x1':
                y = cv(ah, wh)
            elif scheme == 'fp16x3' or not big:
                y = cv(ah, wh) + cv(h16(al), wh) + cv(ah, h16(wl))
            elif scheme in ('fp16+fp8', 'fp16+fp8a', 'fp16+fp8w'):
                y = cv(ah, wh)
                if scheme != 'fp16+fp8w':
                    y = y + cv(e4m3(al * 2.0 ** 14), e4m3(wh)) * 2.0 ** -14
                if scheme != 'fp16+fp8a':
                    y = y + cv(e4m3(ah * 4.0), e4m3(wl * 2.0 ** 12)) * 2.0 ** -14
            elif scheme == 'fp16+fp4':
                y = cv(ah, wh) + cv(e2m1_blocks(al, 1), e2m1_blocks(wh, 1)) + cv(e2m1_blocks(ah, 1), e2m1_blocks(wl, 1))
            else:
                raise ValueError(scheme)
            y = y / sc
        if b is not None:
            y = y + b[None, :, None, None]
        return y
    return conv


SCHEMES = ['fp16x1', 'fp16x3', 'fp16+fp8', 'fp16+fp8a', 'fp16+fp8w', 'fp16+fp4']


def run(height, width, seeds, schemes=SCHEMES):
    errs = {s: [] for s in schemes}
    for seed in range(seeds):
        torch.manual_seed(2021 + seed)
        net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1).double().eval()
        x = torch.rand(1, 3, height, width, generator=torch.Generator().manual_seed(seed), dtype=torch.float32).double()
        with torch.no_grad():
            ref = net.forward_reference(x, conv=make_conv('fp64'))[:, :3]
            for s in schemes:
                out = net.forward_reference(x, conv=make_conv(s))[:, :3]
                errs[s].append(float((out - ref).norm() / ref.norm()))
    return errs


def main():
    height = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    schemes = SCHEMES
    errs = run(height, width, seeds)
    print(json.dumps({'input': [1, 3, height, width], 'seeds': seeds, 'network': 'TransPoseNet enc+2/dec+2, default init',
                      'rel_l2_error_of_coordinate_map': {s: errs[s] for s in schemes},
                      'mma_equivalents_per_product_on_the_large_layers': {'fp16x1': 1, 'fp16x3': 3, 'fp16+fp8': 2, 'fp16+fp8a': 1.5,
                                                                          'fp16+fp8w': 1.5, 'fp16+fp4': 1.5}}))


if __name__ == '__main__':
    main()
