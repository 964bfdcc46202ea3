"""Training-step convolutions: forward, data gradient and weight gradient on the tensor-core kernels, as a
torch.autograd.Function (BASELINE config 4 / SURVEY.md section 8a row a19).

The reference trains through stock autograd (/root/reference/train_single_task.py:262-299): every nn.Conv2d
contributes a cuDNN forward, dgrad and wgrad kernel.  Here
  forward  = cl_conv_igemm on the padded-flat layout (same kernel as inference, fp16x3, no GroupNorm statistics),
  dgrad    = cl_conv_igemm on the output gradient with the transposed filter and negated tap shifts (a stride-2
             convolution decomposes into one small stride-1 problem per input parity phase),
  wgrad    = cl_conv_wgrad_pf: the padded-flat operands of the forward input and of the output gradient read as
             MN-major tensor-core operands (split-K over pixel rows, fp32 atomics),
with the NCHW <-> operand-layout conversions and the filter packing done by cl_nchw_to_pf / cl_pf_to_nchw /
cl_pack_filter.  GroupNorm / ReLU / residual adds and the loss stay stock torch ops in this round.
Gradients are rescaled by a power of two computed on the device (no host synchronisation) so that they stay
inside fp16's range; filters are scaled the same way.
"""
import ctypes

import torch

from . import _lib
from .cnn import _Geometry

_NTERMS = 3


def eligible(conv):
    """Convolutions the tensor-core path covers: 3x3 / 1x1, stride 1 or 2, Cin % 32 == 0, Cout in {64, 128k}."""
    k = conv.kernel_size[0]
    return (k in (1, 3) and conv.kernel_size[1] == k and conv.stride[0] in (1, 2) and conv.padding[0] == k // 2
            and (conv.in_channels == 32 or conv.in_channels % 64 == 0)
            and (conv.out_channels == 64 or conv.out_channels % 128 == 0)
            and conv.dilation[0] == 1 and conv.groups == 1)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _i32(values):
    return (ctypes.c_int32 * len(values))(*values)


def _pow2_scale(t, target):
    """Device scalars (2^k, 2^-k) bringing max|t| to about `target`, computed on the GPU (no host sync)."""
    lib = _lib.load()
    ws = torch.empty(4, dtype=torch.float32, device=t.device)
    _lib.check(lib.cl_pow2_scale(t.data_ptr(), t.numel(), float(target), ws.data_ptr(), _stream(t)))
    return ws[0:1], ws[1:2]


def _pack(weight, scale, pairs, transpose, n, k):
    """fp16 hi/lo [2][taps][n][k] operand of the filter taps `pairs` ((kh, kw) list), x scale."""
    lib = _lib.load()
    cout, cin, ks, _ = weight.shape
    out = torch.empty(2, len(pairs), n, k, dtype=torch.float16, device=weight.device)
    _lib.check(lib.cl_pack_filter(weight.data_ptr(), scale.data_ptr(), out.data_ptr(), cout, cin, ks, len(pairs),
                                  _i32([p[0] for p in pairs]), _i32([p[1] for p in pairs]), 1 if transpose else 0, n, k,
                                  _stream(weight)))
    return out


def _to_pf(x, phases, geo, scale=None):
    lib = _lib.load()
    b, c, h, w = x.shape
    out = torch.zeros(2 * phases * geo.Mp, c, dtype=torch.float16, device=x.device)
    _lib.check(lib.cl_nchw_to_pf(x.data_ptr(), 0 if scale is None else scale.data_ptr(), out.data_ptr(), b, c, h, w,
                                 phases, _stream(x)))
    return out


def _igemm(act_pf, in_phases, geo, packed, taps, cin, cout):
    """Sum over taps of shifted GEMMs on a PF activation; returns the raw fp32 PF matrix [Mp][cout]."""
    lib = _lib.load()
    raw = torch.empty(geo.Mp, cout, dtype=torch.float32, device=act_pf.device)
    zero_bias = torch.zeros(cout, dtype=torch.float32, device=act_pf.device)
    _lib.check(lib.cl_conv_igemm(act_pf.data_ptr(), act_pf.size(0), in_phases * geo.Mp, cin, packed.data_ptr(), cout,
                                 len(taps), _i32(taps), _NTERMS, geo.Mp, geo.Hp, geo.Wp, 0, 1.0, raw.data_ptr(),
                                 zero_bias.data_ptr(), 0, 0, 0, 0, 0, _stream(act_pf)))
    return raw


def _from_raw(raw, geo, craw, out, step=1, off=(0, 0), scale=None, bias=None):
    lib = _lib.load()
    b, c, hout, wout = out.shape
    _lib.check(lib.cl_pf_to_nchw(raw.data_ptr(), geo.B, geo.H, geo.W, craw, out.data_ptr(), c, hout, wout, step, off[0],
                                 off[1], 0 if scale is None else scale.data_ptr(), 0 if bias is None else bias.data_ptr(),
                                 _stream(raw)))
    return out


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


def conv_forward(x, weight, bias, stride, return_operand=False):
    """y = conv2d(x, weight, bias, stride, padding = k // 2); NCHW fp32 in and out.
    With return_operand the fp16 hi/lo PF operand of x is returned too (the weight gradient reads it again)."""
    cout, cin, k, _ = weight.shape
    b, _, h, w = x.shape
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
    geo = _Geometry(b, ho, wo)
    phases = 4 if stride == 2 else 1
    act = _to_pf(x, phases, geo)
    ws, ws_inv = _pow2_scale(weight, 128.0)
    pairs = [(kh, kw) for kh in range(k) for kw in range(k)]
    packed = _pack(weight, ws, pairs, False, cout, cin)
    raw = _igemm(act, phases, geo, packed, _forward_taps(k, stride, geo), cin, cout)
    y = torch.empty(b, cout, ho, wo, dtype=torch.float32, device=x.device)
    y = _from_raw(raw, geo, cout, y, scale=ws_inv, bias=bias)
    return (y, act) if return_operand else y


def grad_operand(gy):
    """fp16 hi/lo PF operand of an output gradient, rescaled by a device-side power of two: (g_pf, 2^-k, geometry)."""
    b, _, ho, wo = gy.shape
    gs, gs_inv = _pow2_scale(gy, 256.0)
    geo = _Geometry(b, ho, wo)
    return _to_pf(gy, 1, geo, gs), gs_inv, geo


def conv_dgrad(gy, weight, in_hw, stride, operand=None):
    """dL/dx of y = conv2d(x, weight, stride, padding = k // 2); gy NCHW fp32 [B, Cout, Ho, Wo]."""
    cout, cin, k, _ = weight.shape
    b, _, ho, wo = gy.shape
    h, w = in_hw
    g_pf, gs_inv, geo = operand if operand is not None else grad_operand(gy)
    ws, ws_inv = _pow2_scale(weight, 128.0)
    out_scale = gs_inv * ws_inv
    n_out = (cin + 63) // 64 * 64          # the kernel wants Cout' % 64 == 0: zero-padded filter rows
    gx = torch.empty(b, cin, h, w, dtype=torch.float32, device=gy.device)
    if stride == 1:
        if k == 1:
            pairs, shifts = [(0, 0)], [0]
        else:
            # x[p] feeds y[p - (kh-1, kw-1)] through tap (kh, kw): dX[p] = sum dY[p + (1-kh, 1-kw)] W[kh][kw]
            pairs = [(kh, kw) for kh in range(3) for kw in range(3)]
            shifts = [(1 - kh) * geo.Wp + (1 - kw) for kh, kw in pairs]
        raw = _igemm(g_pf, 1, geo, _pack(weight, ws, pairs, True, n_out, cout), shifts, cout, n_out)
        return _from_raw(raw, geo, n_out, gx, scale=out_scale)
    # stride 2: input rows of parity a receive from kh = 1 (a = 0: oy = i) or kh = 0 (oy = i + 1), 2 (oy = i) (a = 1)
    per_parity = {0: [(1, 0)], 1: [(0, 1), (2, 0)]} if k == 3 else {0: [(0, 0)], 1: []}
    if h % 2 or w % 2 or k == 1:
        gx.zero_()                          # phases a 1x1 stride-2 filter never touches / cropped odd rows
    for a in (0, 1):
        for bb in (0, 1):
            pairs = [(kh, kw) for kh, _ in per_parity[a] for kw, _ in per_parity[bb]]
            if not pairs:
                continue
            shifts = [dy * geo.Wp + dx for _, dy in per_parity[a] for _, dx in per_parity[bb]]
            raw = _igemm(g_pf, 1, geo, _pack(weight, ws, pairs, True, n_out, cout), shifts, cout, n_out)
            _from_raw(raw, geo, n_out, gx, step=2, off=(a, bb), scale=out_scale)
    return gx


def conv_wgrad(gy, x, weight_shape, stride, operand=None, act=None):
    """dL/dweight of y = conv2d(x, weight, stride, padding = k // 2); returns [Cout, Cin, k, k] fp32.

    Both GEMM operands are the padded-flat matrices the forward / data-gradient kernels already use (cl_conv_wgrad_pf
    reads them as MN-major tensor-core operands): `operand` = grad_operand(gy), `act` = the forward's operand of x.
    """
    lib = _lib.load()
    cout, cin, k, _ = weight_shape
    g_pf, gs_inv, geo = operand if operand is not None else grad_operand(gy)
    phases = 4 if stride == 2 else 1
    if act is None:
        act = _to_pf(x, phases, geo)
    shifts, tphase = [], []
    for t in _forward_taps(k, stride, geo):   # split the forward tap rows into (phase plane, in-plane shift)
        ph = (t + geo.Mp // 2) // geo.Mp if stride == 2 else 0
        tphase.append(ph)
        shifts.append(t - ph * geo.Mp)
    dw = torch.zeros(cout, cin, k, k, dtype=torch.float32, device=g_pf.device)   # written in OIHW, x 2^-k, by the kernel
    _lib.check(lib.cl_conv_wgrad_pf(g_pf.data_ptr(), geo.Mp, act.data_ptr(), geo.Mp, geo.Mp, cout, cin, phases, k * k,
                                    _i32(shifts), _i32(tphase), _NTERMS, 1.0, gs_inv.data_ptr(), 1, dw.data_ptr(),
                                    _stream(g_pf)))
    return dw


class NativeConv2d(torch.autograd.Function):
    """conv2d (padding k // 2, stride 1 | 2) whose forward, dgrad and wgrad run on the sm_100a tensor-core kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        x = x.contiguous()
        y, act = conv_forward(x, weight.contiguous(), bias, stride, return_operand=True)
        ctx.save_for_backward(act, weight)   # the PF operand of x is all the weight gradient needs
        ctx.stride = stride
        ctx.in_hw = tuple(x.shape[2:])
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        act, weight = ctx.saved_tensors
        gy = gy.contiguous()
        weight = weight.contiguous()
        operand = grad_operand(gy)   # shared by the data and the weight gradient
        gx = conv_dgrad(gy, weight, ctx.in_hw, ctx.stride, operand) if ctx.needs_input_grad[0] else None
        gw = conv_wgrad(gy, None, weight.shape, ctx.stride, operand, act) if ctx.needs_input_grad[1] else None
        gb = gy.sum((0, 2, 3)) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


def conv2d(conv, x):
    """Apply an nn.Conv2d through the native autograd function when it is eligible, else through torch."""
    if x.is_cuda and x.dtype == torch.float32 and eligible(conv):
        return NativeConv2d.apply(x, conv.weight, conv.bias, conv.stride[0])
    return conv(x)
