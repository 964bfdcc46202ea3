"""Training-step convolutions: forward, data gradient and weight gradient on the tensor-core kernels, as a
torch.autograd.Function (BASELINE config 4 / SURVEY.md section 8a row a19).

The reference trains through stock autograd (/root/reference/train_single_task.py:262-299): every nn.Conv2d
contributes a cuDNN forward, dgrad and wgrad kernel.  Here
  forward  = cl_conv_igemm on the padded-flat layout (same kernel as inference, fp16x3, no GroupNorm statistics),
  dgrad    = cl_conv_igemm on the output gradient with the transposed, flipped filter (a stride-2 convolution
             decomposes into one small stride-1 convolution per input parity phase),
  wgrad    = cl_conv_wgrad on channel-major operands (split-K over images, fp32 atomics).
GroupNorm / ReLU / residual adds and the loss stay stock torch ops in this round, and the NCHW <-> kernel-layout
conversions are torch ops as well: this is a first correct native training path, not yet a fast one.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib, layout
from .cnn import _Geometry

_NTERMS = 3


def eligible(conv):
    """Convolutions the tensor-core path covers: 3x3 / 1x1, stride 1 or 2, Cin % 32 == 0, Cout in {64, 128k}."""
    k = conv.kernel_size[0]
    return (k in (1, 3) and conv.kernel_size[1] == k and conv.stride[0] in (1, 2) and conv.padding[0] == k // 2
            and conv.in_channels % 32 == 0 and (conv.out_channels == 64 or conv.out_channels % 128 == 0)
            and conv.dilation[0] == 1 and conv.groups == 1)


def _pack_taps(w_taps):
    """[taps][N][K] fp32 -> (fp16 [2][taps][N][K] hi/lo planes scaled by 2^k, 2^-k)."""
    amax = w_taps.abs().amax().clamp_min(1e-30)
    scale = torch.exp2(torch.floor(torch.log2(128.0 / amax)))       # device scalar, no host sync
    w = w_taps * scale
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    return torch.stack([hi, lo], 0).contiguous(), 1.0 / scale


def _igemm(act_pf, in_phases, geo, packed, taps, cin, cout):
    """Sum over taps of shifted GEMMs on a PF activation; returns the raw fp32 PF matrix [Mp][cout]."""
    lib = _lib.load()
    raw = torch.empty(geo.Mp, cout, dtype=torch.float32, device=act_pf.device)
    zero_bias = torch.zeros(cout, dtype=torch.float32, device=act_pf.device)
    arr = (ctypes.c_int32 * len(taps))(*taps)
    _lib.check(lib.cl_conv_igemm(act_pf.data_ptr(), act_pf.size(0), in_phases * geo.Mp, cin, packed.data_ptr(), cout,
                                 len(taps), arr, _NTERMS, geo.Mp, geo.Hp, geo.Wp, 0, 1.0, raw.data_ptr(),
                                 zero_bias.data_ptr(), 0, 0, 0, 0, 0,
                                 torch.cuda.current_stream(act_pf.device).cuda_stream))
    return raw


def _amax_scale(t):
    """Power-of-two factor (device scalar) that brings max|t| to about 2^8: keeps gradients inside fp16's range."""
    amax = t.abs().amax().clamp_min(1e-30)
    return torch.exp2(torch.floor(torch.log2(256.0 / amax)))


def conv_forward(x, weight, stride):
    """y = conv2d(x, weight, stride, padding = k // 2) without bias; NCHW fp32 in and out."""
    cout, cin, k, _ = weight.shape
    b, _, h, w = x.shape
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
    geo = _Geometry(b, ho, wo)
    phases = 4 if stride == 2 else 1
    act = layout.to_pf(x, phases=phases, terms=2)
    packed, inv = _pack_taps(weight.permute(2, 3, 0, 1).reshape(k * k, cout, cin).to(torch.float32))
    taps = _forward_taps(k, stride, geo)
    raw = _igemm(act, phases, geo, packed, taps, cin, cout)
    return layout.raw_to_nchw(raw, b, ho, wo) * inv


def _forward_taps(k, stride, geo):
    if k == 1:
        return [0]
    out = []
    for kh in range(3):
        for kw in range(3):
            if stride == 1:
                out.append((kh - 1) * geo.Wp + (kw - 1))
            else:
                a, dy = (1, -1) if kh == 0 else ((0, 0) if kh == 1 else (1, 0))
                bb, dx = (1, -1) if kw == 0 else ((0, 0) if kw == 1 else (1, 0))
                out.append((a * 2 + bb) * geo.Mp + dy * geo.Wp + dx)
    return out


def conv_dgrad(gy, weight, in_hw, stride):
    """dL/dx of y = conv2d(x, weight, stride, padding = k // 2); gy NCHW fp32 [B, Cout, Ho, Wo]."""
    cout, cin, k, _ = weight.shape
    b, _, ho, wo = gy.shape
    h, w = in_hw
    s = _amax_scale(gy)
    geo = _Geometry(b, ho, wo)
    g_pf = layout.to_pf(gy * s, phases=1, terms=2)
    n_out = cin if cin % 64 == 0 else ((cin + 63) // 64) * 64      # the kernel wants Cout' % 64 == 0: zero-pad
    wt = weight.to(torch.float32)

    def taps_weight(pairs):
        """[taps][cin'][cout] for (kh, kw) pairs: dX[ci] = sum_co dY[co] * W[co][ci][kh][kw]."""
        mats = torch.stack([wt[:, :, kh, kw].t() for kh, kw in pairs], 0)      # [taps][cin][cout]
        if n_out != cin:
            mats = F.pad(mats, (0, 0, 0, n_out - cin))
        return mats.contiguous()

    if stride == 1:
        if k == 1:
            pairs, shifts = [(0, 0)], [0]
        else:
            # x[p] feeds y[p - (kh-1, kw-1)] through tap (kh, kw): dX[p] = sum dY[p + (1-kh, 1-kw)] W[kh][kw]
            pairs = [(kh, kw) for kh in range(3) for kw in range(3)]
            shifts = [(1 - kh) * geo.Wp + (1 - kw) for kh, kw in pairs]
        packed, inv = _pack_taps(taps_weight(pairs))
        raw = _igemm(g_pf, 1, geo, packed, shifts, cout, n_out)
        gx = layout.raw_to_nchw(raw, b, ho, wo)[:, :cin]
        return gx * (inv / s)
    # stride 2: input rows of parity a receive from kh in {1} (a = 0: oy = i) or {0 (oy = i + 1), 2 (oy = i)} (a = 1)
    per_parity = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}
    gx = torch.zeros(b, cin, 2 * ho, 2 * wo, dtype=torch.float32, device=gy.device)
    for a in (0, 1):
        for bb in (0, 1):
            pairs = [(kh, kw) for kh, _ in per_parity[a] for kw, _ in per_parity[bb]]
            shifts = [dy * geo.Wp + dx for _, dy in per_parity[a] for _, dx in per_parity[bb]]
            packed, inv = _pack_taps(taps_weight(pairs))
            raw = _igemm(g_pf, 1, geo, packed, shifts, cout, n_out)
            gx[:, :, a::2, bb::2] = layout.raw_to_nchw(raw, b, ho, wo)[:, :cin] * (inv / s)
    return gx[:, :, :h, :w].contiguous()


def _to_cm(x, hp, wp, col0):
    """NCHW fp32 -> channel-major fp16 hi/lo planes [2][B][C][hp * wp]: x placed at rows 1.., columns col0.., zeros elsewhere."""
    b, c, h, w = x.shape
    p = F.pad(x, (col0, wp - w - col0, 1, hp - h - 1)).reshape(b, c, hp * wp)
    hi = p.to(torch.float16)
    lo = (p - hi.to(torch.float32)).to(torch.float16)
    return torch.stack([hi, lo], 0)


def conv_wgrad(gy, x, weight_shape, stride):
    """dL/dweight of y = conv2d(x, weight, stride, padding = k // 2); returns [Cout, Cin, k, k] fp32.

    The kernel shifts whole rows only (TMA box starts must be 16-byte aligned), so the row pitch is padded to a
    multiple of 8 pixels and the horizontal neighbours of a 3x3 filter are supplied as column-shifted copies of x.
    """
    lib = _lib.load()
    cout, cin, k, _ = weight_shape
    b, _, ho, wo = gy.shape
    hp = ho + 2
    wp = (wo + 3 + 7) // 8 * 8                                         # room for the -1 column shift, pitch % 8 == 0
    plane = hp * wp
    s = _amax_scale(gy)
    g_cm = _to_cm(gy * s, hp, wp, 1).contiguous()                      # [2][B][Cout][plane]
    groups, shifts, tphase = [], [], []
    if k == 1:
        groups.append(_to_cm(x if stride == 1 else x[:, :, ::2, ::2], hp, wp, 1))
        shifts, tphase = [0], [0]
    elif stride == 1:
        # group kw holds x shifted by (kw - 1) columns: reading it at row shift (kh - 1) gives x[p + (kh-1, kw-1)]
        for kw in range(3):
            groups.append(_to_cm(x, hp, wp, 1 - (kw - 1)))
        for kh in range(3):
            for kw in range(3):
                shifts.append((kh - 1) * wp); tphase.append(kw)
    else:
        # parity phases of x at the output resolution, each with column shifts 0 and -1 (dx of the tap)
        index = {}
        for a in (0, 1):
            for bb in (0, 1):
                sub = x[:, :, a::2, bb::2]
                sub = F.pad(sub, (0, wo - sub.size(3), 0, ho - sub.size(2)))
                for dx in (0, -1):
                    index[(a, bb, dx)] = len(groups)
                    groups.append(_to_cm(sub, hp, wp, 1 - dx))
        for kh in range(3):
            for kw in range(3):
                a, dy = (1, -1) if kh == 0 else ((0, 0) if kh == 1 else (1, 0))
                bb, dx = (1, -1) if kw == 0 else ((0, 0) if kw == 1 else (1, 0))
                shifts.append(dy * wp); tphase.append(index[(a, bb, dx)])
    x_cm = torch.stack(groups, 1).contiguous()                        # [2][groups][B][Cin][plane]
    dw = torch.zeros(k * k, cout, cin, dtype=torch.float32, device=gy.device)
    a_shift = (ctypes.c_int32 * len(shifts))(*shifts)
    a_phase = (ctypes.c_int32 * len(tphase))(*tphase)
    _lib.check(lib.cl_conv_wgrad(g_cm.data_ptr(), x_cm.data_ptr(), b, cout, cin, plane, plane, len(groups), k * k, a_shift,
                                 a_phase, _NTERMS, 1.0, dw.data_ptr(), torch.cuda.current_stream(gy.device).cuda_stream))
    return (dw / s).reshape(k, k, cout, cin).permute(2, 3, 0, 1).contiguous()


class NativeConv2d(torch.autograd.Function):
    """conv2d (padding k // 2, stride 1 | 2) whose forward, dgrad and wgrad run on the sm_100a tensor-core kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        y = conv_forward(x, weight, stride)
        if bias is not None:
            y = y + bias[None, :, None, None]
        ctx.save_for_backward(x, weight)
        ctx.stride = stride
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = conv_dgrad(gy, weight, x.shape[2:], ctx.stride) if ctx.needs_input_grad[0] else None
        gw = conv_wgrad(gy, x, weight.shape, ctx.stride) if ctx.needs_input_grad[1] else None
        gb = gy.sum((0, 2, 3)) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


def conv2d(conv, x):
    """Apply an nn.Conv2d through the native autograd function when it is eligible, else through torch."""
    if x.is_cuda and eligible(conv):
        return NativeConv2d.apply(x, conv.weight, conv.bias, conv.stride[0])
    return conv(x)
