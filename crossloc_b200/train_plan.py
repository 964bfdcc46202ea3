"""Fused training step of the coordinate network (BASELINE config 4).

The reference trains through stock autograd (/root/reference/train_single_task.py:262-299): every nn.Conv2d /
nn.GroupNorm / F.relu / residual add contributes its own forward and backward library kernels on NCHW fp32 tensors.
Here the whole network is ONE autograd node:

  forward   the inference plan of crossloc_b200.cnn (fp16 + fp8 scheme by default) with every raw convolution
            output, GroupNorm statistic and operand kept (the engine's `tape`);
  backward  walks the tape in reverse.  Per stage: cl_gn_backward pass 0 (sum of the incoming gradients, residual
            mask, GroupNorm / ReLU reductions) and pass 1 (gradient of the raw convolution output as fp16 hi / lo
            padded-flat planes, power-of-two scaled on the device), then the data gradient (cl_conv_igemm with the
            transposed filter) and the weight gradient (cl_conv_wgrad_pf) read those planes directly.

Activations and gradients never leave the padded-flat layout between layers, so the NCHW <-> operand conversions of
the per-layer path (crossloc_b200.train) disappear.  The stem's GroupNorm / ReLU backward runs on the same kernels from the
raw conv1 output the forward keeps (its 864-entry weight gradient is one cuDNN wgrad call), the 4-channel head has its own
backward kernel (cl_head_backward).  Covers TransPoseNet / Network without MLR encoders or the full-size head;
the other variants keep using the per-layer path.
"""
import ctypes
import os

import torch

from . import _lib
from .cnn import CoordNetEngine, _nterms_for, _W8_LO_SCALE, pack_fp4
from .train import _i32, _pack

_NTERMS = 3
# arithmetic of the data gradient GEMMs: 'fp16+fp4' (default: a_hi*w_hi in fp16 plus the two correction products as
# block-scaled e2m1 MMAs, the inference scheme, on the layers whose channel counts are multiples of 256; the others stay
# fp16x3), 'fp16x3' (three fp16 MMAs per product) or 'fp16x1' (one fp16 pass with fp32 accumulation: the 10-bit mantissa
# of the TF32 kernels stock PyTorch trains with by default)
BACKWARD = os.environ.get('CROSSLOC_B200_TRAIN_BACKWARD', 'fp16+fp4')
# arithmetic of the forward convolutions: the inference default ('fp16+fp4': block-scaled e2m1 correction terms on the
# 256/512-channel layers, 1.7e-4 relative on the coordinate map), 'fp16+fp8' (e4m3 correction terms, 3e-5) or 'fp16x3'
FORWARD = os.environ.get('CROSSLOC_B200_TRAIN_FORWARD', 'fp16+fp4')
# arithmetic of the weight gradient GEMMs alone: one fp16 pass by default.  A weight gradient is one sum over every pixel of
# the batch (5,400 x batch terms per entry at 60 x 90 cells): the operand rounding averages out and, unlike in the data gradient,
# does not travel on through the layers below.  Measured on the BASELINE config-4 step (tools/dbg_wgrad_precision.py, batch
# 12, same forward and data gradients): all gradients move by 3.1e-5 relative L2 (worst convolution weight 8.1e-5, median
# 1.3e-5) against fp16x3 weight gradients, for 4.7 ms less per step.  'fp16x3' restores the three-term products.
WGRAD = os.environ.get('CROSSLOC_B200_TRAIN_WGRAD', 'fp16x1')
_RESCALE_EVERY = 64   # steps between host-side refreshes of the filters' power-of-two scales


class _Src:
    """One gradient contribution to an activation: fp32 padded-flat buffer x two optional device scalars."""

    def __init__(self, g, stride, scale_a=None, scale_b=None, phased=False):
        self.g, self.stride, self.scale_a, self.scale_b, self.phased = g, stride, scale_a, scale_b, phased


class _TrainPack:
    """Filter of one convolution in the tensor-core layout, packed on the device (no host synchronisation)."""

    def __init__(self, conv, exp, nterms=_NTERMS):
        w = conv.weight.detach().contiguous()
        cout, cin, k, _ = w.shape
        self.cin, self.cout, self.ksize, self.stride, self.nterms = cin, cout, k, conv.stride[0], nterms
        self.out_scale = float(2.0 ** (-exp))
        self.scale = torch.full((1,), float(2.0 ** exp), dtype=torch.float32, device=w.device)
        self.inv_scale = torch.full((1,), self.out_scale, dtype=torch.float32, device=w.device)
        self.weight = w
        self.weight_param, self.bias_param = conv.weight, conv.bias
        self.weights = _pack(w, self.scale, [(kh, kw) for kh in range(k) for kw in range(k)], False, cout, cin)
        self.weights8 = None
        if nterms == 2:   # forward in the fp16 + fp8 scheme: e4m3 planes fp8(w_hi), fp8(w_lo * 2^12)
            self.weights8 = torch.stack([self.weights[0].to(torch.float32).to(torch.float8_e4m3fn),
                                         (self.weights[1].to(torch.float32) * _W8_LO_SCALE).to(torch.float8_e4m3fn)], 0).contiguous()
        self.weights4 = self.w_sf = None
        if nterms == 4:   # forward in the fp16 + fp4 scheme: block-scaled e2m1 planes of the scaled filter
            self.weights4, self.w_sf = pack_fp4(w.to(torch.float32), 2.0 ** exp)
        self.bias = (conv.bias.detach().to(torch.float32) if conv.bias is not None
                     else torch.zeros(cout, dtype=torch.float32, device=w.device)).contiguous()


def supported(net):
    return not getattr(net, 'num_mlr', 0) and not getattr(net, 'full_size_output', False)


class TrainPlan:
    def __init__(self, net, backward=None, forward=None, wgrad=None):
        self.net = net
        forward = forward or FORWARD
        backward = backward or BACKWARD
        wgrad = wgrad or ('fp16x1' if backward == 'fp16x1' else WGRAD)
        if backward not in ('fp16+fp4', 'fp16x3', 'fp16x1') or wgrad not in ('fp16x3', 'fp16x1'):
            raise ValueError('unknown backward arithmetic %r / %r (fp16+fp4 | fp16x3 | fp16x1)' % (backward, wgrad))
        self.backward_mode = backward
        self.bwd_terms = 1 if backward == 'fp16x1' else 3      # layers outside the fp4 scheme
        self.bwd_fp4 = backward == 'fp16+fp4'
        self.wgrad_terms = 3 if wgrad == 'fp16x3' else 1
        self._zero_bias = {}
        self.engine = CoordNetEngine(precision=forward, fp4=forward == 'fp16+fp4')
        self.engine.packer = self._packer
        self._exps = {}
        self._step = 0
        self._pool = {}

    # ------------------------------------------------------------------ filters
    def _packer(self, name, conv):
        if name not in self._exps or self._step % _RESCALE_EVERY == 0:
            amax = float(conv.weight.detach().abs().max())   # host sync, once every _RESCALE_EVERY steps
            import math
            exp = 0 if amax == 0.0 or not math.isfinite(amax) else int(math.floor(math.log2(128.0 / amax)))
            self._exps[name] = max(-24, min(24, exp))
        nterms = _nterms_for(self.engine.precision, conv.in_channels, conv.kernel_size[0], conv.stride[0], conv.out_channels)
        return _TrainPack(conv, self._exps[name], nterms)

    # ------------------------------------------------------------------ forward
    def forward(self, image):
        eng = self.engine
        eng.tape = []
        spec = self.net._spec()
        out = eng.forward(spec, image)
        # the tape points into workspace buffers the next forward overwrites: remember which forward a state belongs to
        state = {'tape': eng.tape, 'head_in': eng.head_in, 'stem_out': eng.stem_out, 'stem_rec': eng.stem_rec, 'spec': spec,
                 'image': image, 'out': out.detach(),
                 'geo3': eng.head_in.geo, 'generation': self._step + 1}
        eng.tape = None
        self._step += 1
        return out, state

    # ------------------------------------------------------------------ helpers
    def _zeros_pf(self, geo, channels, device):
        """fp16 hi/lo planes with zero borders, recycled between layers of one (resolution, width)."""
        key = (geo.B, geo.H, geo.W, channels, str(device))
        buf = self._pool.get(key)
        if buf is None:
            if len(self._pool) > 16:
                self._pool.clear()
            buf = self._pool[key] = torch.zeros(2 * geo.Mp, channels, dtype=torch.float16, device=device)
        return buf

    def _fp4_planes(self, geo, channels, device):
        """e2m1 planes + scale words of a scaled gradient (zero borders, written once), recycled like `_zeros_pf`."""
        key = ('fp4', geo.B, geo.H, geo.W, channels, str(device))
        buf = self._pool.get(key)
        if buf is None:
            buf = self._pool[key] = (torch.zeros(2 * geo.Mp, channels // 2, dtype=torch.uint8, device=device),
                                     torch.zeros(channels // 256 * geo.Mp, dtype=torch.int32, device=device))
        return buf

    def _fp4_dgrad(self, pack):
        """The data gradient of this convolution runs in fp16 + fp4 mode: stride 1, both channel counts multiples of 256."""
        return self.bwd_fp4 and pack.stride == 1 and pack.cin % 256 == 0 and pack.cout % 256 == 0

    def _gn_backward(self, lib, stream, geo, channels, rec, norm, relu_inner, srcs, mask, want_g, d_raw_f32=None, acc=None):
        """Both passes of cl_gn_backward for one stage; returns (d_raw, scale_out, ab, dbias, g_buf).
        `acc` = (ab pointer, channels per image of the shared ab buffer, dbias view, two-double scratch view) places the
        accumulators of this stage inside buffers shared by the whole backward pass (one fill, one reduction at the end)."""
        dev = rec['raw'].device
        if acc is None:
            n_ab = geo.B * channels * 2
            z = torch.zeros(n_ab + channels + 2, dtype=torch.float64, device=dev)   # one fill for all the accumulators
            ab, ab_ptr, ab_C = z[:n_ab].view(geo.B, channels, 2), z.data_ptr(), 0
            dbias = z[n_ab:n_ab + channels]
            misc = z[n_ab + channels:]
        else:
            ab_ptr, ab_C, dbias, misc = acc
            ab = None
        gmax = misc[0:1].view(torch.int32)
        scale_out = misc[1:2].view(torch.float32)
        g_buf = torch.empty(geo.Mp, channels, dtype=torch.float32, device=dev) if want_g else None
        d_raw = self._zeros_pf(geo, channels, dev)
        fp4 = self._fp4_planes(geo, channels, dev) if ('pack' in rec and self._fp4_dgrad(rec['pack'])) else None
        rec['d_raw4'] = fp4
        group_ch = 0 if norm is None else channels // norm.num_groups

        def call(pass_id, sources):
            n = len(sources)
            ptrs = (ctypes.c_void_p * 3)(*[s.g.data_ptr() for s in sources] + [None] * (3 - n))
            sa = (ctypes.c_void_p * 3)(*[None if s.scale_a is None else s.scale_a.data_ptr() for s in sources] + [None] * (3 - n))
            sb = (ctypes.c_void_p * 3)(*[None if s.scale_b is None else s.scale_b.data_ptr() for s in sources] + [None] * (3 - n))
            _lib.check(lib.cl_gn_backward_fp4(
                pass_id, geo.B, geo.H, geo.W, channels, group_ch, rec['raw'].data_ptr(),
                None if norm is None else rec['stats'].data_ptr(), None if norm is None else norm.weight.data_ptr(),
                None if norm is None else norm.bias.data_ptr(), 1e-5 if norm is None else float(norm.eps),
                1 if relu_inner else 0, n, ptrs, sa, sb, _i32([s.stride for s in sources] + [0] * (3 - n)),
                _i32([1 if s.phased else 0 for s in sources] + [0] * (3 - n)),
                None if (mask is None or pass_id == 1) else mask.data_ptr(),
                None if (g_buf is None or pass_id == 1) else g_buf.data_ptr(), ab_ptr, ab_C, gmax.data_ptr(),
                d_raw.data_ptr(), geo.Mp, scale_out.data_ptr(), dbias.data_ptr(),
                None if (d_raw_f32 is None or pass_id == 0) else d_raw_f32.data_ptr(),
                None if fp4 is None else fp4[0].data_ptr(), geo.Mp, None if fp4 is None else fp4[1].data_ptr(), stream))

        call(0, srcs)
        call(1, [_Src(g_buf, channels)] if want_g else srcs[:1])
        return d_raw, scale_out, ab, dbias, g_buf

    def _conv_backward(self, lib, stream, rec, d_raw, scale_out, need_dgrad, gw):
        """Weight gradient into `gw` (zeroed OIHW tensor) and data gradient (a source for the producer of the input)."""
        pack, geo, act = rec['pack'], rec['geo'], rec['act']
        k, stride, cin, cout = pack.ksize, pack.stride, pack.cin, pack.cout
        dev = d_raw.device
        # ---- weight gradient: dW[tap][co][ci] from the PF planes of d_raw and of the forward operand
        shifts, tphase = [], []
        for t in rec['taps']:
            ph = (t + geo.Mp // 2) // geo.Mp if stride == 2 else 0
            tphase.append(ph)
            shifts.append(t - ph * geo.Mp)
        _lib.check(lib.cl_conv_wgrad_pf(d_raw.data_ptr(), geo.Mp, act.h16.data_ptr(), geo.Mp, geo.Mp, cout, cin, act.phases,
                                        k * k, _i32(shifts), _i32(tphase), self.wgrad_terms, 1.0, scale_out[1:2].data_ptr(), 1,
                                        gw.data_ptr(), stream))
        if not need_dgrad:
            return None
        # ---- data gradient: transposed filter, negated tap shifts; one small problem per input parity for stride 2
        n_out = (cin + 63) // 64 * 64
        zero_bias = self._zero_bias.get((n_out, str(dev)))
        if zero_bias is None:
            zero_bias = self._zero_bias[(n_out, str(dev))] = torch.zeros(n_out, dtype=torch.float32, device=dev)

        def igemm(pairs, shifts, raw):
            packed = _pack(pack.weight, pack.scale, pairs, True, n_out, cout)
            fp4 = rec.get('d_raw4')
            if fp4 is not None and stride == 1 and len(pairs) == k * k:
                # fp16 + fp4: e2m1 planes of the transposed filter, packed on the device from its OIHW view
                w_t = pack.weight.permute(1, 0, 2, 3).contiguous()
                w4 = torch.empty(2 * len(pairs) * n_out, cout // 2, dtype=torch.uint8, device=dev)
                w_sf = torch.empty(len(pairs) * (cout // 256) * n_out, dtype=torch.int32, device=dev)
                _lib.check(lib.cl_pack_conv_fp4(w_t.data_ptr(), n_out, cout, len(pairs), 1.0 / pack.out_scale, w4.data_ptr(),
                                                w_sf.data_ptr(), stream))
                _lib.check(lib.cl_conv_igemm_fp4(d_raw.data_ptr(), d_raw.size(0), geo.Mp, cout, packed.data_ptr(), n_out,
                                                 len(shifts), _i32(shifts), geo.Mp, geo.Hp, geo.Wp, 0, 1.0, raw.data_ptr(),
                                                 zero_bias.data_ptr(), None, fp4[0].data_ptr(), fp4[0].size(0), geo.Mp,
                                                 fp4[1].data_ptr(), w4.data_ptr(), w_sf.data_ptr(), stream))
                return
            _lib.check(lib.cl_conv_igemm(d_raw.data_ptr(), d_raw.size(0), geo.Mp, cout, packed.data_ptr(), n_out,
                                         len(shifts), _i32(shifts), self.bwd_terms, geo.Mp, geo.Hp, geo.Wp, 0, 1.0,
                                         raw.data_ptr(), zero_bias.data_ptr(), 0, 0, 0, 0, 0, stream))

        if stride == 1:
            raw = torch.empty(geo.Mp, n_out, dtype=torch.float32, device=dev)
            if k == 1:
                igemm([(0, 0)], [0], raw)
            else:
                pairs = [(kh, kw) for kh in range(3) for kw in range(3)]
                igemm(pairs, [(1 - kh) * geo.Wp + (1 - kw) for kh, kw in pairs], raw)
            return _Src(raw, n_out, scale_out[1:2], pack.inv_scale, phased=False)
        raw = (torch.empty if k == 3 else torch.zeros)(4 * geo.Mp, n_out, dtype=torch.float32, device=dev)
        per_parity = {0: [(1, 0)], 1: [(0, 1), (2, 0)]} if k == 3 else {0: [(0, 0)], 1: []}
        for a in (0, 1):
            for bb in (0, 1):
                pairs = [(kh, kw) for kh, _ in per_parity[a] for kw, _ in per_parity[bb]]
                if pairs:
                    shifts = [dy * geo.Wp + dx for _, dy in per_parity[a] for _, dx in per_parity[bb]]
                    igemm(pairs, shifts, raw[(a * 2 + bb) * geo.Mp:(a * 2 + bb + 1) * geo.Mp])
        return _Src(raw, n_out, scale_out[1:2], pack.inv_scale, phased=True)

    # ------------------------------------------------------------------ backward
    def backward(self, state, g_out):
        """Gradients of every parameter given dL/d(output); returns {parameter: gradient}."""
        if state['generation'] != self._step:
            raise RuntimeError('crossloc_b200: backward through a forward whose activations have been overwritten -- the '
                               'fused training plan keeps ONE forward per network (call backward before the next forward, '
                               'or use forward_train(x, fused=False))')
        lib = _lib.load()
        image = state['image']
        dev = image.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        grads = {}
        spec = state['spec']
        head = spec['head']
        geo3 = state['geo3']
        res = state['head_in']

        # ---- head (1x1 C -> Co, mean offset, exp(clamp)): the derivative of the output maps on the small [B,Co,Hc,Wc] tensor
        # in torch (d exp(clamp(s)) = out where the clamp is inactive), then cl_head_backward on the padded-flat activation
        hconv = head['conv']
        co, k_task = hconv.out_channels, head['num_task']
        g_sc = g_out.to(torch.float32).contiguous().clone()
        if co > k_task:
            pos = state['out'][:, k_task:]
            lo = torch.exp(torch.tensor(head['clamp'][0], dtype=torch.float32, device=dev))
            hi = torch.exp(torch.tensor(head['clamp'][1], dtype=torch.float32, device=dev))
            g_sc[:, k_task:] = g_sc[:, k_task:] * pos * ((pos > lo) & (pos < hi))
        w2d = hconv.weight.detach().reshape(co, -1).to(torch.float32).contiguous()
        g_w = torch.zeros_like(w2d)
        g_pf = torch.empty(geo3.Mp, w2d.size(1), dtype=torch.float32, device=dev)
        _lib.check(lib.cl_head_backward(res.h16.data_ptr(), geo3.Mp, geo3.B, geo3.H, geo3.W, w2d.size(1), co, w2d.data_ptr(),
                                        g_sc.data_ptr(), g_pf.data_ptr(), g_w.data_ptr(), stream))
        grads[id(hconv.weight)] = g_w.reshape(hconv.weight.shape)
        grads[id(hconv.bias)] = g_sc.sum((0, 2, 3))
        sources = {id(res): [_Src(g_pf, g_pf.size(1))]}

        # one zeroed buffer for every convolution's weight gradient (the kernels accumulate into OIHW views of it)
        recs = [e['conv'] for e in state['tape']] + [e['skip'] for e in state['tape'] if e['skip'] is not None]
        flat = torch.zeros(sum(r['pack'].weight.numel() for r in recs), dtype=torch.float32, device=dev)
        offset = 0
        for r in recs:
            n = r['pack'].weight.numel()
            r['gw'] = flat[offset:offset + n].view(r['pack'].weight.shape)
            offset += n

        # shared accumulators of all stages: ab [B][sumC][2], dbias [sumC], two scratch doubles per stage -- one fill now, one
        # reduction over the batch and one fp32 cast at the end instead of three small launches per stage
        stages = [(e['conv'], e['norm']) for e in state['tape']] + \
                 [(e['skip'], e['norm2']) for e in state['tape'] if e['skip'] is not None]
        if state['stem_out'] is not None:
            stages.append((state['stem_rec'], state['stem_rec']['norm']))
        batch = geo3.B
        sum_c = sum(32 if 'pack' not in r else r['pack'].cout for r, _ in stages)
        zall = torch.zeros(batch * sum_c * 2 + sum_c + 2 * len(stages), dtype=torch.float64, device=dev)
        z_ab, z_db, z_misc = zall[:batch * sum_c * 2], zall[batch * sum_c * 2:batch * sum_c * 2 + sum_c], zall[batch * sum_c * 2 + sum_c:]
        offs, c_off = {}, 0
        for i, (r, _) in enumerate(stages):
            ch = 32 if 'pack' not in r else r['pack'].cout
            offs[id(r)] = (z_ab.data_ptr() + c_off * 16, sum_c, z_db[c_off:c_off + ch], z_misc[2 * i:2 * i + 2], c_off, ch)
            c_off += ch

        stem_out = state['stem_out']
        for e in reversed(state['tape']):
            rec, out = e['conv'], e['out']
            geo, channels = rec['geo'], rec['pack'].cout
            srcs = sources.pop(id(out))
            merge = e['add_kind'] != 0
            mask = out.h16 if (merge and e['relu_outer']) else None
            want_g = merge or len(srcs) > 1
            d_raw, scale_out, _, _, g_buf = self._gn_backward(lib, stream, geo, channels, rec, e['norm'], e['relu_inner'], srcs,
                                                              mask, want_g, acc=offs[id(rec)][:4])
            if e['add_kind'] == 1:
                sources.setdefault(id(e['res']), []).append(_Src(g_buf, channels))
            act = rec['act']
            src = self._conv_backward(lib, stream, rec, d_raw, scale_out, True, rec['gw'])
            grads[id(rec['pack'].weight_param)] = rec['gw']
            sources.setdefault(id(act), []).append(src)
            if e['add_kind'] == 2:
                # skip branch: out = [relu](GroupNorm(skip_conv(res)) + main): its gradient is the merged gradient g_buf
                srec = e['skip']
                d_raw_s, scale_s, _, _, _ = self._gn_backward(lib, stream, geo, channels, srec, e['norm2'], False,
                                                              [_Src(g_buf, channels)], None, False, acc=offs[id(srec)][:4])
                src = self._conv_backward(lib, stream, srec, d_raw_s, scale_s, True, srec['gw'])
                grads[id(srec['pack'].weight_param)] = srec['gw']
                sources.setdefault(id(srec['act']), []).append(src)

        # ---- stem (conv1 + norm1 + relu on the 1- or 3-channel frame): GroupNorm / ReLU backward on the native kernels from
        # the raw conv1 output the forward kept (the four-phase data gradient of conv2 is read in place); the 864-entry
        # weight gradient is left to torch (cuDNN wgrad on the fp32 NCHW copy of the result)
        if stem_out is not None:
            srcs = sources.pop(id(stem_out))
            rec = state['stem_rec']
            conv1, norm1, geo0 = rec['conv'], rec['norm'], rec['geo']
            d_f32 = torch.empty(geo0.Mp, 32, dtype=torch.float32, device=dev)
            self._gn_backward(lib, stream, geo0, 32, rec, norm1, True, srcs, None, False, d_raw_f32=d_f32, acc=offs[id(rec)][:4])
            g_conv = torch.empty(geo0.B, 32, geo0.H, geo0.W, dtype=torch.float32, device=dev)
            _lib.check(lib.cl_pf_to_nchw(d_f32.data_ptr(), geo0.B, geo0.H, geo0.W, 32, g_conv.data_ptr(), 32, geo0.H, geo0.W, 1, 0,
                                         0, None, None, stream))
            grads[id(conv1.weight)] = torch.nn.grad.conv2d_weight(image, conv1.weight.shape, g_conv, padding=1)

        # ---- per-channel parameter gradients of every stage from the shared accumulators
        ab_sum = z_ab.view(batch, sum_c, 2).sum(0).to(torch.float32)   # [sumC][2]: (d_beta, d_gamma)
        db = z_db.to(torch.float32)
        for r, norm in stages:
            _, _, _, _, c0, ch = offs[id(r)]
            conv = r['conv'] if 'pack' not in r else None
            bias_param = conv.bias if conv is not None else r['pack'].bias_param
            if bias_param is not None:
                grads[id(bias_param)] = db[c0:c0 + ch]
            if norm is not None:
                grads[id(norm.bias)] = ab_sum[c0:c0 + ch, 0]
                grads[id(norm.weight)] = ab_sum[c0:c0 + ch, 1]
        return grads


class _FusedStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, image, *params):
        out, state = plan.forward(image)
        ctx.plan, ctx.state, ctx.params = plan, state, params
        return out

    @staticmethod
    def backward(ctx, g_out):
        grads = ctx.plan.backward(ctx.state, g_out)
        ctx.state = None
        return (None, None) + tuple(grads.get(id(p)) if ctx.needs_input_grad[2 + i] else None for i, p in enumerate(ctx.params))


def forward_train(net, image, backward=None, forward=None, wgrad=None):
    """Differentiable forward of `net` through the fused plan (one autograd node for the whole network)."""
    plan = getattr(net, '_train_plan', None)
    if (plan is None or (backward is not None and plan.backward_mode != backward)
            or (wgrad is not None and plan.wgrad_terms != (3 if wgrad == 'fp16x3' else 1))
            or (forward is not None and plan.engine.precision != forward)):
        plan = TrainPlan(net, backward, forward, wgrad)
        object.__setattr__(net, '_train_plan', plan)
    params = tuple(p for p in net.parameters())
    return _FusedStep.apply(plan, image.contiguous().to(torch.float32), *params)
