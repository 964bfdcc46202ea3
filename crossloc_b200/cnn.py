"""Host side of the scene-coordinate CNN: weight packing, workspace and the layer plan.

The arithmetic runs in libcrossloc_b200.so (cl_stem_forward, cl_conv_igemm, cl_gn_apply, cl_head_forward);
this file only lays tensors out in HBM and sequences the launches for the two network families of the
reference (/root/reference/networks/networks.py): `TransPoseNet` (:363-502, GroupNorm) and the vanilla
DSAC* `Network` (:43-130, no normalisation).  PyTorch is used for device memory and streams only.
"""
import ctypes
import math
import os

import torch

from . import _lib

# Convolution arithmetic (CROSSLOC_B200_CONV_PRECISION):
#   'fp16+fp8'            a_hi*w_hi on the fp16 tensor pipe + the two 2^-11 correction products as e4m3 MMAs for the
#                         large 3x3 layers (85 % of the FLOPs), fp16x3 elsewhere: 2e-5 relative on the coordinate map
#   'fp16+fp4' (default)  the same with the corrections as block-scaled e2m1 MMAs (kind::mxf4, 4x the fp16 rate): 1.5 instead
#                         of 2 fp16-MMA equivalents per product, ~1.5e-4 relative (C++ runtime only)
#   'fp16x3'              every product as three fp16 MMAs: 1e-5 relative
#   'fp16x1'              one pass: 1.1e-3 relative -- misses the 1e-3 parity bar, offered for speed comparisons only
PRECISION = os.environ.get('CROSSLOC_B200_CONV_PRECISION', 'fp16+fp4')
_PRECISIONS = ('fp16+fp8', 'fp16x3', 'fp16x1')
_W8_LO_SCALE = 4096.0   # csrc/conv.h kW8LoScale
_FP8_1X1 = os.environ.get('CROSSLOC_B200_FP8_1X1', '1') != '0'   # 1x1 512->512 layers in the fp16 + fp8 scheme too


def _nterms_for(precision, cin, ksize, stride, cout=None):
    """MMA scheme of one convolution: 1 = fp16, 2 = fp16 + e4m3 corrections, 3 = fp16x3, 4 = fp16 + block-scaled e2m1
    corrections (only with `cout`: both channel counts must be multiples of 256)."""
    if precision == 'fp16x1':
        return 1
    wide = stride == 1 and cin % 128 == 0 and cin >= 256 and (ksize == 3 or (_FP8_1X1 and cin >= 512))
    if precision == 'fp16+fp4' and wide:
        return 4 if (cout is not None and cin % 256 == 0 and cout % 256 == 0) else 2
    if precision == 'fp16+fp8' and wide:
        return 2
    return 3


class PackedConv:
    """Weights of one convolution in the tensor-core layout: fp16 [term][tap][Cout][Cin] (+ e4m3 planes), scaled by 2^k."""

    def __init__(self, weight, bias, stride, nterms):
        cout, cin, kh, kw = weight.shape
        assert kh == kw and kh in (1, 3) and stride in (1, 2)
        self.cin, self.cout, self.ksize, self.stride, self.nterms = cin, cout, kh, stride, nterms
        w = weight.detach().to(torch.float32)
        amax = float(w.abs().max())
        # power-of-two pre-scale keeps the low-order terms inside the fp16 / e4m3 normal ranges; undone in the epilogue
        exp = 0 if amax == 0.0 else int(math.floor(math.log2(128.0 / amax)))
        exp = max(-24, min(24, exp))
        self.out_scale = float(2.0 ** (-exp))
        w = (w * (2.0 ** exp)).permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)   # [tap][Cout][Cin]
        hi = w.to(torch.float16)
        lo = w - hi.to(torch.float32)
        planes = [hi]
        if nterms == 3:
            planes.append(lo.to(torch.float16))
        self.weights = torch.stack(planes, 0).contiguous()
        self.weights8 = None
        if nterms == 2:
            self.weights8 = torch.stack([hi.to(torch.float32).to(torch.float8_e4m3fn),
                                         (lo * _W8_LO_SCALE).to(torch.float8_e4m3fn)], 0).contiguous()
        self.weights4 = self.w_sf = None
        if nterms == 4:
            self.weights4, self.w_sf = pack_fp4(weight.detach().to(torch.float32).contiguous(), 2.0 ** exp)
        self.bias = (bias.detach().to(torch.float32) if bias is not None
                     else torch.zeros(cout, dtype=torch.float32, device=weight.device)).contiguous()


def pack_fp4(weight, scale):
    """OIHW fp32 filter (x scale, a power of two) -> (e2m1 planes [2 * taps * Cout][Cin / 2], scale words) of cl_pack_conv_fp4."""
    cout, cin, kh, kw = weight.shape
    taps = kh * kw
    w4 = torch.empty(2 * taps * cout, cin // 2, dtype=torch.uint8, device=weight.device)
    w_sf = torch.empty(taps * (cin // 256) * cout, dtype=torch.int32, device=weight.device)
    _lib.check(_lib.load().cl_pack_conv_fp4(weight.data_ptr(), cout, cin, taps, float(scale), w4.data_ptr(), w_sf.data_ptr(),
                                            torch.cuda.current_stream(weight.device).cuda_stream))
    return w4, w_sf


class _Geometry:
    """Padded-flat geometry of one resolution level."""

    def __init__(self, batch, h, w):
        self.B, self.H, self.W = batch, h, w
        self.Hp, self.Wp = h + 2, w + 2
        self.plane = self.Hp * self.Wp
        self.Mp = batch * self.plane


class _PF:
    """One padded-flat activation: fp16 hi/lo planes and, on demand, the e4m3 planes of the fp16 + fp8 mode."""

    def __init__(self, geo, channels, phases, terms, device):
        self.geo, self.channels, self.phases, self.terms = geo, channels, phases, terms
        # zero-initialised once: kernels only ever write interior pixels, so the borders stay zero
        self.h16 = torch.zeros(terms * phases * geo.Mp, channels, dtype=torch.float16, device=device)
        self._f8 = None
        self._f4 = None
        self._device = device

    @property
    def f8(self):
        if self._f8 is None:
            self._f8 = torch.zeros(2 * self.phases * self.geo.Mp, self.channels, dtype=torch.uint8, device=self._device)
        return self._f8

    @property
    def f4(self):
        """(e2m1 planes [2 * Mp][C / 2], scale words [C / 256 * Mp]) of the fp16 + fp4 mode (same-resolution activations)."""
        if self._f4 is None:
            assert self.phases == 1 and self.channels % 256 == 0
            self._f4 = (torch.zeros(2 * self.geo.Mp, self.channels // 2, dtype=torch.uint8, device=self._device),
                        torch.zeros(self.channels // 256 * self.geo.Mp, dtype=torch.int32, device=self._device))
        return self._f4


def _taps(pack, geo):
    """Activation row shift of every filter tap in the padded-flat layout of the OUTPUT resolution."""
    if pack.ksize == 1:
        return [0]
    out = []
    for kh in range(3):
        for kw in range(3):
            if pack.stride == 1:
                out.append((kh - 1) * geo.Wp + (kw - 1))
            else:
                a, dy = (1, -1) if kh == 0 else ((0, 0) if kh == 1 else (1, 0))
                b, dx = (1, -1) if kw == 0 else ((0, 0) if kw == 1 else (1, 0))
                out.append((a * 2 + b) * geo.Mp + dy * geo.Wp + dx)
    return out


class CoordNetEngine:
    """Runs the coordinate network of an nn.Module twin (networks.networks.TransPoseNet / Network)."""

    def __init__(self, precision=None, fp4=False):
        self.precision = precision or PRECISION
        if self.precision == 'fp16+fp4' and not fp4:
            # inference runs the e2m1 corrections in the C++ runtime (csrc/net.cu); this plan keeps e4m3 unless asked
            # (fp4=True: the training forward of crossloc_b200.train_plan, plans without MLR merge / full-size head)
            self.precision = 'fp16+fp8'
        if self.precision not in _PRECISIONS and self.precision != 'fp16+fp4':
            raise ValueError('unknown conv precision %r (%s)' % (self.precision, ' | '.join(_PRECISIONS)))
        self.terms = 1 if self.precision == 'fp16x1' else 2
        self.nterms = {'fp16x1': 1, 'fp16x3': 3, 'fp16+fp8': 2, 'fp16+fp4': 1.5}[self.precision]   # scheme of the dominant 3x3 layers
        self._packs = {}
        self._pack_versions = {}
        self._ws = {}
        self.launches = 0
        self.events = None   # set to a list to record (layer, shape, flops, start, end) CUDA events around every launch
        # training mode (crossloc_b200.train_plan): a list that receives one record per GroupNorm/merge stage; every
        # layer then keeps its own raw / activation buffers (nothing is recycled) and `packer` supplies the filters
        self.tape = None
        self.packer = None
        self.head_in = None
        self._conv_rec = {}
        self._keep_i = 0
        self._shared = {}
        self.no_fp4 = set()   # ids of convolutions that must not use the fp16 + fp4 scheme (see nterms_of)

    # ------------------------------------------------------------------ parameters
    def _pack(self, name, conv, force_split=False):
        if self.packer is not None:
            return self.packer(name, conv)
        ver = (conv.weight._version, conv.weight.data_ptr(), None if conv.bias is None else conv.bias._version, force_split)
        if self._pack_versions.get(name) != ver:
            nterms = self.nterms_of(conv)
            if force_split and nterms in (2, 4):
                nterms = 3   # the operand comes from outside the plan and has no e4m3 / e2m1 planes
            self._packs[name] = PackedConv(conv.weight, conv.bias, conv.stride[0], nterms)
            self._pack_versions[name] = ver
        return self._packs[name]

    # ------------------------------------------------------------------ workspace
    def _workspace(self, device, batch, h, w, channels):
        key = (str(device), batch, h, w, self.terms, tuple(sorted(channels.items())))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        ws = {'geo': {}, 'act': {}, 'raw': {}}
        while len(self._ws) >= 3:   # training frames change size every step (dataloader.py:520-524): keep a few workspaces only
            self._ws.pop(next(iter(self._ws)))
        hh, wwid = h, w
        for level in range(4):   # level 0 = input resolution, 3 = output resolution
            ws['geo'][level] = _Geometry(batch, hh, wwid)
            if level < 3:
                hh, wwid = (hh + 1) // 2, (wwid + 1) // 2
        self._ws[key] = ws
        ws['device'] = device
        return ws

    def _act(self, ws, tag, level, channels, phases):
        key = (tag, level, channels, phases)
        buf = ws['act'].get(key)
        if buf is None:
            buf = _PF(ws['geo'][level], channels, phases, self.terms, ws['device'])
            ws['act'][key] = buf
        return buf

    def _raw(self, ws, tag, level, channels, name=None):
        if self.tape is not None and name is not None:
            tag = 'keep:' + name
        key = (tag, level, channels)
        buf = ws['raw'].get(key)
        if buf is None:
            geo = ws['geo'][level]
            buf = torch.empty(geo.Mp, channels, dtype=torch.float32, device=ws['device'])
            ws['raw'][key] = buf
        return buf

    # ------------------------------------------------------------------ operators
    def _tick(self):
        if self.events is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _tock(self, e0, name, shape, flops):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.events.append((name, shape, flops, e0, e1))

    def _conv(self, stream, pack, act, geo, raw, stats, group_ch, name=None):
        taps = _taps(pack, geo)
        tap_arr = (ctypes.c_int32 * len(taps))(*taps)
        lo_rows = act.phases * geo.Mp
        use8 = pack.nterms == 2
        e0 = self._tick()
        if pack.nterms == 4:
            f4, sf = act.f4
            _lib.check(self._lib.cl_conv_igemm_fp4(
                act.h16.data_ptr(), act.h16.size(0), lo_rows, pack.cin, pack.weights.data_ptr(), pack.cout, len(taps),
                tap_arr, geo.Mp, geo.Hp, geo.Wp, group_ch, pack.out_scale, raw.data_ptr(), pack.bias.data_ptr(),
                0 if stats is None else stats.data_ptr(), f4.data_ptr(), f4.size(0), geo.Mp, sf.data_ptr(),
                pack.weights4.data_ptr(), pack.w_sf.data_ptr(), stream))
        else:
            _lib.check(self._lib.cl_conv_igemm(
                act.h16.data_ptr(), act.h16.size(0), lo_rows, pack.cin, pack.weights.data_ptr(), pack.cout, len(taps),
                tap_arr, pack.nterms, geo.Mp, geo.Hp, geo.Wp, group_ch, pack.out_scale, raw.data_ptr(),
                pack.bias.data_ptr(), 0 if stats is None else stats.data_ptr(),
                act.f8.data_ptr() if use8 else 0, act.f8.size(0) if use8 else 0, lo_rows if use8 else 0,
                pack.weights8.data_ptr() if use8 else 0, stream))
        # algorithmic FLOPs: 2 * (real output pixels) * Cout * Cin * taps (borders and split terms excluded)
        self._tock(e0, name, (pack.cin, pack.cout, pack.ksize, pack.stride),
                   2.0 * geo.B * geo.H * geo.W * pack.cout * pack.cin * len(taps))
        self.launches += 1
        if self.tape is not None:
            self._conv_rec[id(raw)] = {'name': name, 'pack': pack, 'act': act, 'geo': geo, 'taps': taps, 'raw': raw,
                                       'stats': stats, 'group_ch': group_ch}

    def _apply(self, stream, raw, geo, channels, norm, stats, out, relu_inner=True, res=None, raw2=None, norm2=None,
               stats2=None, relu_outer=False, want_lo=True, want8=False, out_c0=None):
        group_ch = 0 if norm is None else channels // norm.num_groups
        add_kind = 1 if res is not None else (2 if raw2 is not None else 0)
        e0 = self._tick()
        want4 = bool(want8 & 2) if not isinstance(want8, bool) else False
        want8 = bool(want8 & 1) if not isinstance(want8, bool) else want8
        _lib.check(self._lib.cl_gn_apply_fp4(
            raw.data_ptr(), geo.B, geo.H, geo.W, channels, group_ch,
            0 if stats is None else stats.data_ptr(),
            0 if norm is None else norm.weight.data_ptr(), 0 if norm is None else norm.bias.data_ptr(),
            1e-5 if norm is None else float(norm.eps), 1 if relu_inner else 0, add_kind,
            0 if res is None else res.h16.data_ptr(), geo.Mp if self.terms == 2 else 0,
            0 if raw2 is None else raw2.data_ptr(), 0 if stats2 is None else stats2.data_ptr(),
            0 if norm2 is None else norm2.weight.data_ptr(), 0 if norm2 is None else norm2.bias.data_ptr(),
            1 if relu_outer else 0, out.h16.data_ptr(), out.phases, 2 if (want_lo and self.terms == 2) else 1,
            out.f8.data_ptr() if want8 else 0, out.channels if out_c0 is not None else 0, out_c0 or 0,
            out.f4[0].data_ptr() if want4 else 0, out.f4[1].data_ptr() if want4 else 0, stream))
        self._tock(e0, 'gn_apply', ('gn_apply', channels, out.phases, add_kind), 0.0)
        self.launches += 1
        if self.tape is not None:
            self.tape.append({'conv': self._conv_rec[id(raw)], 'norm': norm, 'out': out, 'relu_inner': relu_inner,
                              'add_kind': add_kind, 'res': res, 'relu_outer': relu_outer,
                              'skip': None if raw2 is None else self._conv_rec[id(raw2)], 'norm2': norm2})

    def shared_pf(self, key, device, batch, h, w, channels):
        """A padded-flat activation at the output resolution of an (h, w) frame that outlives single forward() calls
        (the concatenated encoder outputs of the MLR model)."""
        for _ in range(3):
            h, w = (h + 1) // 2, (w + 1) // 2
        k = (key, str(device), batch, h, w, channels, self.terms)
        buf = self._shared.get(k)
        if buf is None:
            if len(self._shared) >= 4:
                self._shared.pop(next(iter(self._shared)))
            buf = self._shared[k] = _PF(_Geometry(batch, h, w), channels, 1, self.terms, device)
        return buf

    def nterms_of(self, conv):
        n = _nterms_for(self.precision, conv.in_channels, conv.kernel_size[0], conv.stride[0], conv.out_channels)
        if n == 4 and id(conv) in self.no_fp4:
            n = 2   # its operand is written by a pass without e2m1 planes (MLR: encoder slices, cl_pf_groupnorm)
        return n

    # ------------------------------------------------------------------ plans
    def forward(self, spec, image):
        """spec: dict produced by networks.networks (layer modules + head description). image: NCHW fp32 CUDA.

        Optional spec keys: 'roles' maps conv1..conv4 to layer names (several encoders can share one engine);
        'input': 'image' (default) | 'activation' (`image` is an NCHW fp32 activation at the output resolution: the
        stem and the strided ladder are skipped) | 'pf' (spec['in_pf'] is a padded-flat activation produced by an earlier
        plan; `image` is ignored); 'output': 'head' (default) | 'activation' (returns the final residual stream as an
        NCHW fp32 tensor instead of running a head) | 'pf' (returns it as a padded-flat activation; with
        spec['out_pf'] = (buffer, first channel, want e4m3 planes) the last stage writes into that channel slice)."""
        self._lib = _lib.load()
        in_pf = spec.get('in_pf') if spec.get('input') == 'pf' else None
        if in_pf is not None:
            dev = in_pf.h16.device
            batch, cin, h, w = in_pf.geo.B, in_pf.channels, in_pf.geo.H, in_pf.geo.W
        else:
            if not image.is_cuda:
                raise RuntimeError('crossloc_b200: the coordinate network runs on a CUDA device only (no CPU fallback)')
            image = image.contiguous().to(torch.float32)
            batch, cin, h, w = image.shape
            dev = image.device
        if batch == 0:
            raise RuntimeError('crossloc_b200: empty batch')
        torch.cuda.set_device(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        from_activation = spec.get('input', 'image') in ('activation', 'pf')
        ws = self._workspace(dev, batch, h, w, {'cin': cin, 'act': int(from_activation)})
        geo = ws['geo']
        if from_activation:
            geo[3] = in_pf.geo if in_pf is not None else _Geometry(batch, h, w)
        roles = spec.get('roles', {'conv1': 'conv1', 'conv2': 'conv2', 'conv3': 'conv3', 'conv4': 'conv4'})
        gn = spec['group_norm']
        layers = spec['layers']          # ordered list of (name, conv, norm or None)
        n_stat = len(layers) + 1
        stats_all = ws.get('stats')
        # one [batch][groups][2] slice per layer, sized for the widest grouping in the plan (the convolution epilogue
        # indexes its slice by image * groups + group)
        max_groups = max([32] + [norm.num_groups for _, _, norm in layers if norm is not None])
        if stats_all is None or stats_all.size(0) < n_stat or stats_all.size(2) < max_groups:
            stats_all = torch.zeros(n_stat, batch, max_groups, 2, dtype=torch.float64, device=dev)
            ws['stats'] = stats_all
        else:
            stats_all.zero_()
        self._stat_i = 0
        self._keep_i = 0
        self._conv_rec = {}

        def next_stats():
            s = stats_all[self._stat_i]
            self._stat_i += 1
            return s

        convs = {name: (conv, norm) for name, conv, norm in layers}
        blocks = spec['blocks']
        entry = set()
        if from_activation and in_pf is None and blocks:   # convolutions reading the externally supplied activation
            entry = {blocks[0]['convs'][0]} | ({blocks[0]['skip']} if blocks[0]['kind'] == 'residual_skip' else set())
        packs = {name: self._pack(name, conv, name in entry) for name, conv, _ in layers if name != roles.get('conv1')}

        def groups_of(norm, channels):
            return 0 if norm is None else channels // norm.num_groups

        def planes_for(consumers, also_lo=False):
            """Which operand planes the consumers of an activation need: (fp16 lo plane, e4m3 planes)."""
            # bit 0: e4m3 planes, bit 1: block-scaled e2m1 planes (an int, so that `if want8` keeps its meaning)
            want8 = (1 if any(packs[c].nterms == 2 for c in consumers) else 0) | (2 if any(packs[c].nterms == 4 for c in consumers) else 0)
            # training: the weight gradient reads the fp16 lo plane of every operand, whatever the forward scheme
            want_lo = also_lo or self.tape is not None or any(packs[c].nterms == 3 for c in consumers)
            return want_lo, want8

        duc = spec.get('head', {}).get('duc') if spec.get('output', 'head') == 'head' else None

        def first_conv_of(block_index):
            """Name(s) of the convolution(s) that read the residual stream entering block `block_index`."""
            if block_index >= len(blocks):
                return [duc['name']] if duc else []
            blk = blocks[block_index]
            return [blk['convs'][0]] + ([blk['skip']] if blk['kind'] == 'residual_skip' else [])

        g3 = geo[3]
        out_pf = spec.get('out_pf') if spec.get('output') == 'pf' else None
        if in_pf is not None:
            res = in_pf
        elif from_activation:
            # externally produced activation (e.g. the MLR merge): NCHW fp32 -> fp16 hi/lo padded-flat planes
            res = ws['act'].get('ext')
            if res is None:   # cl_nchw_to_pf always writes both planes, whatever the precision mode
                res = ws['act']['ext'] = _PF(g3, cin, 1, 2, dev)
            _lib.check(self._lib.cl_nchw_to_pf(image.data_ptr(), 0, res.h16.data_ptr(), batch, cin, h, w, 1, stream))
            self.launches += 1
        else:
            # ---- stem: conv1 (+ norm1) + relu, written as the 4-phase input of conv2
            conv1, norm1 = convs[roles['conv1']]
            if conv1.out_channels != 32 or (norm1 is not None and norm1.num_groups != 32):
                raise RuntimeError('crossloc_b200: the stem kernel is built for 32 channels / 32 groups')
            a = self._act(ws, 'stem', 1, 32, 4)
            st = next_stats() if norm1 is not None else None
            stem_raw = None
            if self.tape is not None:   # training: keep the raw conv1 output for the GroupNorm / ReLU backward
                stem_raw = self._raw(ws, 'stem', 0, 32, 'stem')
            e0 = self._tick()
            _lib.check(self._lib.cl_stem_forward(
                image.data_ptr(), batch, cin, h, w, conv1.weight.detach().contiguous().data_ptr(),
                conv1.bias.detach().contiguous().data_ptr(), 1 if norm1 is not None else 0,
                0 if st is None else st.data_ptr(), 0 if norm1 is None else norm1.weight.data_ptr(),
                0 if norm1 is None else norm1.bias.data_ptr(), 1e-5 if norm1 is None else float(norm1.eps),
                a.h16.data_ptr(), self.terms, 0 if stem_raw is None else stem_raw.data_ptr(), stream))
            self._tock(e0, 'stem', ('stem',), 2.0 * batch * h * w * 32 * cin * 9)
            self.launches += 2 if norm1 is not None else 1
            self.stem_out = a if self.tape is not None else None
            self.stem_rec = None if self.tape is None else {'raw': stem_raw, 'stats': st, 'norm': norm1, 'geo': geo[0],
                                                            'conv': conv1}

            # ---- strided ladder conv2..conv4
            for level, role in ((1, 'conv2'), (2, 'conv3'), (3, 'conv4')):
                name = roles[role]
                conv, norm = convs[name]
                pack = packs[name]
                raw = self._raw(ws, 'ladder', level, pack.cout, name)
                st = next_stats() if norm is not None else None
                self._conv(stream, pack, a, geo[level], raw, st, groups_of(norm, pack.cout), name)
                if level < 3:
                    out = self._act(ws, 'ladder', level + 1, pack.cout, 4)
                    self._apply(stream, raw, geo[level], pack.cout, norm, st, out)
                else:
                    out = self._act(ws, 'res', 3, pack.cout, 1)
                    want_lo, want8 = planes_for(first_conv_of(0), also_lo=True)   # residual stream: always hi + lo
                    self._apply(stream, raw, geo[level], pack.cout, norm, st, out, want_lo=want_lo, want8=want8)
                a = out
            res = a

        rot = {}

        def scratch(channels, avoid):
            """A PF buffer at the output resolution that is none of `avoid`."""
            if self.tape is not None:   # training: every activation is kept for the backward pass
                self._keep_i += 1
                return self._act(ws, 'keep%d' % self._keep_i, 3, channels, 1)
            pool = rot.setdefault(channels, [self._act(ws, 'pool%d' % i, 3, channels, 1) for i in range(4)])
            for buf in pool:
                if all(buf is not o for o in avoid):
                    return buf
            raise AssertionError

        def chain(names, x, res_in, outer_relu, next_readers, merge_raw2=None, target=None):
            """conv -> [GN] -> relu per layer; the last layer merges the residual stream:
            out = [relu](res + relu(gn(conv)))  or, with merge_raw2 = (raw, norm, stats), the GroupNorm'ed skip path."""
            for i, name in enumerate(names):
                conv, norm = convs[name]
                pack = packs[name]
                raw = self._raw(ws, 'r%d' % (i % 2), 3, pack.cout, name)
                st = next_stats() if norm is not None else None
                self._conv(stream, pack, x, g3, raw, st, groups_of(norm, pack.cout), name)
                out = scratch(pack.cout, (x, res_in))
                last = i == len(names) - 1
                if not last:
                    want_lo, want8 = planes_for([names[i + 1]])
                    self._apply(stream, raw, g3, pack.cout, norm, st, out, want_lo=want_lo, want8=want8)
                else:
                    want_lo, want8 = planes_for(next_readers, also_lo=True)
                    c0 = None
                    if target is not None:   # the plan's result goes into a channel slice of a caller-owned activation
                        out, c0, want8 = target
                        want_lo = True
                    if merge_raw2 is None:
                        self._apply(stream, raw, g3, pack.cout, norm, st, out, res=res_in, relu_outer=outer_relu,
                                    want_lo=want_lo, want8=want8, out_c0=c0)
                    else:
                        raw_s, snorm, st_s = merge_raw2
                        self._apply(stream, raw, g3, pack.cout, norm, st, out, raw2=raw_s, norm2=snorm, stats2=st_s,
                                    relu_outer=outer_relu, want_lo=want_lo, want8=want8, out_c0=c0)
                x = out
            return x

        outer = gn   # TransPoseNet applies ReLU after every residual add (networks.py:240-254); Network does not (:105-120)
        for bi, block in enumerate(blocks):
            kind = block['kind']
            readers = first_conv_of(bi + 1)
            target = out_pf if bi == len(blocks) - 1 else None
            if target is not None and (kind not in ('residual', 'residual_skip') or not gn):
                raise RuntimeError('crossloc_b200: a sliced plan output needs a final residual block')
            if kind == 'residual':
                res = chain(block['convs'], res, res, outer, readers, target=target)
            elif kind == 'mlr_merge':
                # networks.py:491-494: res = mlr_skip(cat); mlr = mlr_forward(mlr_norm(cat)); res = relu(res + mlr)
                sconv, snorm = convs[block['skip']]
                spack = packs[block['skip']]
                raw_s = self._raw(ws, 'rs', 3, spack.cout, block['skip'])
                st_s = next_stats()
                self._conv(stream, spack, res, g3, raw_s, st_s, groups_of(snorm, spack.cout), block['skip'])
                nin = block['norm_in']
                normed = self._act(ws, 'mlr_normed', 3, res.channels, 1)
                want_lo, want8 = planes_for([block['convs'][0]])
                assert not (want8 & 2), 'the MLR merge has no e2m1 planes (fp16 + fp4 is limited to plans without it)'
                st_n = ws.get('pfgn_stats')
                if st_n is None or st_n.shape != (batch, nin.num_groups, 2):
                    st_n = ws['pfgn_stats'] = torch.zeros(batch, nin.num_groups, 2, dtype=torch.float64, device=dev)
                else:
                    st_n.zero_()
                e0 = self._tick()
                _lib.check(self._lib.cl_pf_groupnorm(
                    res.h16.data_ptr(), g3.Mp, batch, g3.H, g3.W, res.channels, res.channels // nin.num_groups,
                    nin.weight.data_ptr(), nin.bias.data_ptr(), float(nin.eps), st_n.data_ptr(), normed.h16.data_ptr(),
                    2 if want_lo else 1, normed.f8.data_ptr() if want8 else 0, stream))
                self._tock(e0, 'pf_groupnorm', ('pf_groupnorm', res.channels), 0.0)
                self.launches += 2
                res = chain(block['convs'], normed, normed, outer, readers, merge_raw2=(raw_s, snorm, st_s))
            elif kind == 'residual_skip':
                # x = chain(res); res = skip_norm(skip(res)); res = [relu](res + x)   (networks.py:242-249)
                sconv, snorm = convs[block['skip']]
                spack = packs[block['skip']]
                raw_s = self._raw(ws, 'rs', 3, spack.cout, block['skip'])
                st_s = next_stats() if snorm is not None else None
                self._conv(stream, spack, res, g3, raw_s, st_s, groups_of(snorm, spack.cout), block['skip'])
                if gn:
                    res = chain(block['convs'], res, res, outer, readers, merge_raw2=(raw_s, snorm, st_s), target=target)
                else:
                    # vanilla Network: res = skip(res) + relu(conv(x)), no normalisation anywhere
                    skip_pf = scratch(spack.cout, (res,))
                    self._apply(stream, raw_s, g3, spack.cout, None, None, skip_pf, relu_inner=False)
                    res = chain(block['convs'], res, skip_pf, outer, readers)
            elif kind == 'plain':
                # conv -> [GN] -> relu without a residual (fc1, fc2); the last output feeds the head (hi + lo)
                names = block['convs']
                for i, name in enumerate(names):
                    conv, norm = convs[name]
                    pack = packs[name]
                    raw = self._raw(ws, 'r0', 3, pack.cout, name)
                    st = next_stats() if norm is not None else None
                    self._conv(stream, pack, res, g3, raw, st, groups_of(norm, pack.cout), name)
                    out = scratch(pack.cout, (res,))
                    nxt = [names[i + 1]] if i + 1 < len(names) else readers
                    want_lo, want8 = planes_for(nxt, also_lo=True)
                    self._apply(stream, raw, g3, pack.cout, norm, st, out, want_lo=want_lo, want8=want8)
                    res = out
            else:
                raise AssertionError(kind)

        if spec.get('output', 'head') == 'pf':
            return res
        if spec.get('output', 'head') == 'activation':
            from . import layout
            return layout.from_pf(res.h16, batch, g3.H, g3.W, self.terms)

        # ---- head
        head = spec['head']
        hconv = head['conv']
        co = hconv.out_channels
        self.head_in = res if self.tape is not None else None
        if duc:
            # full-size variant (networks.py:344-349): DUC 3x3 convolution on the tensor cores, then one kernel for
            # GroupNorm + ReLU + PixelShuffle + bilinear resize + fc3 + output maps
            dconv, dnorm = convs[duc['name']]
            dpack = packs[duc['name']]
            raw = self._raw(ws, 'duc', 3, dpack.cout)
            st = next_stats()
            self._conv(stream, dpack, res, g3, raw, st, groups_of(dnorm, dpack.cout), duc['name'])
            up_h, up_w = duc['size']
            out = torch.empty(batch, co, up_h, up_w, dtype=torch.float32, device=dev)
            mean = head['mean'].to(device=dev, dtype=torch.float32).contiguous()
            e0 = self._tick()
            _lib.check(self._lib.cl_duc_head_forward(
                raw.data_ptr(), batch, g3.H, g3.W, dpack.cout, co, duc['rate'], groups_of(dnorm, dpack.cout),
                st.data_ptr(), dnorm.weight.data_ptr(), dnorm.bias.data_ptr(), float(dnorm.eps),
                hconv.weight.detach().reshape(co, -1).contiguous().data_ptr(), hconv.bias.detach().contiguous().data_ptr(),
                mean.data_ptr(), head['num_task'], head['clamp'][0], head['clamp'][1], out.data_ptr(), up_h, up_w,
                stream))
            self._tock(e0, 'duc_head', ('duc_head',), 0.0)
            self.launches += 1
            self._keepalive = (mean,)
            return out
        out = torch.empty(batch, co, g3.H, g3.W, dtype=torch.float32, device=dev)
        mean = head['mean'].to(device=dev, dtype=torch.float32).contiguous()
        e0 = self._tick()
        _lib.check(self._lib.cl_head_forward(
            res.h16.data_ptr(), g3.Mp, self.terms, batch, g3.H, g3.W, hconv.in_channels, co,
            hconv.weight.detach().reshape(co, -1).contiguous().data_ptr(),
            hconv.bias.detach().contiguous().data_ptr(), mean.data_ptr(), head['num_task'],
            head['clamp'][0], head['clamp'][1], out.data_ptr(), stream))
        self._tock(e0, 'head', ('head',), 0.0)
        self.launches += 1
        # keep parameter temporaries alive until the stream has consumed them
        self._keepalive = (mean,)
        return out
