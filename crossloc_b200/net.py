"""Thin caller of the C++ network runtime (csrc/net.cu: cl_net_create / cl_net_forward).

The plan, the packed filters, the TMA tensor maps, every workspace and the CUDA graph live in the library; this
file only turns an nn.Module twin (networks.networks.TransPoseNet / Network) into the `cl_net_desc` layer table --
state-dict tensors as plain device pointers -- and keeps the handle in step with the module's parameters
(`load_state_dict`, optimizer steps, `.to()`).  It replaces `network(image)` of the reference's evaluation loop
(/root/reference/test_single_task.py:347, utils/evaluation.py:106-116).
"""
import ctypes
import os

import torch

from . import _lib

_c = ctypes
MAX_BLOCK_CONVS = 4
BLOCK_KINDS = {'residual': 0, 'residual_skip': 1, 'plain': 2}
PRECISIONS = {'fp16x1': 1, 'fp16+fp8': 2, 'fp16x3': 3, 'fp16+fp4': 4}
# fp16-MMA equivalents issued per product on the large layers
MMA_EQUIVALENTS = {'fp16x1': 1.0, 'fp16+fp8': 2.0, 'fp16x3': 3.0, 'fp16+fp4': 1.5}
OP_KINDS = {0: 'memset', 1: 'stem', 2: 'conv', 3: 'gn_apply', 4: 'head', 5: 'duc_head', 6: 'raw_stats', 7: 'frames', 8: 'fork', 9: 'conv_fused'}


class NetLayer(_c.Structure):
    _fields_ = [('cin', _c.c_int32), ('cout', _c.c_int32), ('ksize', _c.c_int32), ('stride', _c.c_int32),
                ('weight', _c.c_void_p), ('bias', _c.c_void_p), ('gn_groups', _c.c_int32),
                ('gn_weight', _c.c_void_p), ('gn_bias', _c.c_void_p), ('gn_eps', _c.c_float)]


class NetBlock(_c.Structure):
    _fields_ = [('kind', _c.c_int32), ('n_convs', _c.c_int32), ('convs', _c.c_int32 * MAX_BLOCK_CONVS),
                ('skip', _c.c_int32)]


class NetDesc(_c.Structure):
    _fields_ = [('abi_version', _c.c_int32), ('precision', _c.c_int32), ('relu_after_add', _c.c_int32),
                ('n_layers', _c.c_int32), ('layers', _c.POINTER(NetLayer)), ('stem', _c.c_int32 * 4),
                ('n_blocks', _c.c_int32), ('blocks', _c.POINTER(NetBlock)), ('head_layer', _c.c_int32),
                ('head_mean', _c.c_void_p), ('num_task', _c.c_int32), ('clamp_lo', _c.c_float),
                ('clamp_hi', _c.c_float), ('duc_layer', _c.c_int32), ('duc_rate', _c.c_int32)]


def supported(spec):
    """The C++ runtime covers the plans that start at the image and end in a head (everything the reference's
    evaluation builds for num_mlr == 0); MLR merges and training tapes stay on the Python plan (crossloc_b200.cnn)."""
    if spec.get('input', 'image') != 'image' or spec.get('output', 'head') != 'head':
        return False
    return all(b['kind'] in BLOCK_KINDS for b in spec['blocks'])


class NetRuntime:
    """One cl_net handle per (module, device); rebuilt when parameter storage moves, refreshed when values change."""

    def __init__(self, precision=None):
        from . import cnn
        self.precision = precision or cnn.PRECISION
        if self.precision not in PRECISIONS:
            raise ValueError('unknown conv precision %r (%s)' % (self.precision, ' | '.join(PRECISIONS)))
        self.nterms = MMA_EQUIVALENTS[self.precision]
        self._handle = None
        self._addresses = None
        self._versions = None
        self._keep = None
        self._lib = None
        self.launches = 0            # kernels launched through this runtime so far
        self.launches_per_forward = 0
        self.profiling = False

    def __deepcopy__(self, memo):
        return NetRuntime(self.precision)   # a copied module gets its own handle on first use

    def __getstate__(self):
        return {'precision': self.precision}

    def __setstate__(self, state):
        self.__init__(state['precision'])

    # ------------------------------------------------------------------ handle management
    def _tensors(self, spec):
        out = []
        for _, conv, norm in spec['layers']:
            out += [conv.weight, conv.bias] + ([norm.weight, norm.bias] if norm is not None else [])
        out += [spec['head']['conv'].weight, spec['head']['conv'].bias, spec['head']['mean']]
        return [t for t in out if t is not None]

    def _destroy(self):
        if self._handle is not None and self._lib is not None:
            self._lib.cl_net_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:   # interpreter shutdown
            pass

    def _create(self, spec, device):
        lib = self._lib
        layers = list(spec['layers'])
        names = [n for n, _, _ in layers]
        head = spec['head']
        layers.append(('__head__', head['conv'], None))
        index = {n: i for i, n in enumerate(names)}
        keep = []

        def ptr(t, dtype=torch.float32):
            if t is None:
                return None
            t = t.detach()
            if t.dtype != dtype or not t.is_contiguous() or t.device != device:
                t = t.to(device=device, dtype=dtype).contiguous()
                keep.append(t)   # a converted copy: cl_net_update re-reads it, so it is refreshed by _sync below
            return t.data_ptr()

        arr = (NetLayer * len(layers))()
        for i, (_, conv, norm) in enumerate(layers):
            L = arr[i]
            L.cin, L.cout = conv.in_channels, conv.out_channels
            L.ksize, L.stride = conv.kernel_size[0], conv.stride[0]
            if conv.kernel_size[0] != conv.kernel_size[1] or conv.padding[0] != conv.kernel_size[0] // 2 or conv.groups != 1:
                raise RuntimeError('crossloc_b200: unsupported convolution %r' % (conv,))
            L.weight, L.bias = ptr(conv.weight), ptr(conv.bias)
            if norm is not None:
                L.gn_groups, L.gn_eps = norm.num_groups, float(norm.eps)
                L.gn_weight, L.gn_bias = ptr(norm.weight), ptr(norm.bias)
        blocks = (NetBlock * max(1, len(spec['blocks'])))()
        for i, b in enumerate(spec['blocks']):
            blocks[i].kind = BLOCK_KINDS[b['kind']]
            blocks[i].n_convs = len(b['convs'])
            for k, name in enumerate(b['convs']):
                blocks[i].convs[k] = index[name]
            blocks[i].skip = index[b['skip']] if b['kind'] == 'residual_skip' else -1
        roles = spec.get('roles', {r: r for r in ('conv1', 'conv2', 'conv3', 'conv4')})
        desc = NetDesc()
        desc.abi_version = 1
        desc.precision = PRECISIONS[self.precision]
        desc.relu_after_add = 1 if spec['group_norm'] else 0
        desc.n_layers, desc.layers = len(layers), arr
        for k, r in enumerate(('conv1', 'conv2', 'conv3', 'conv4')):
            desc.stem[k] = index[roles[r]]
        desc.n_blocks, desc.blocks = len(spec['blocks']), blocks
        desc.head_layer = len(layers) - 1
        desc.head_mean = ptr(head['mean'])
        desc.num_task = int(head['num_task'])
        desc.clamp_lo, desc.clamp_hi = float(head['clamp'][0]), float(head['clamp'][1])
        duc = head.get('duc')
        desc.duc_layer = index[duc['name']] if duc else -1
        desc.duc_rate = int(duc['rate']) if duc else 0
        handle = _c.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.cl_net_create(_c.byref(desc), _c.byref(handle)))
        self._handle = handle
        self._keep = keep

    def _sync(self, spec, device):
        """Create / refresh the handle so that it reflects the module's current parameters."""
        if self._lib is None:
            self._lib = _lib.load()
        tensors = self._tensors(spec)
        addresses = (str(device), tuple((t.data_ptr(), tuple(t.shape), t.dtype) for t in tensors), tuple(spec['blocks'][i]['kind'] for i in range(len(spec['blocks']))))
        versions = tuple(t._version for t in tensors)
        if self._handle is None or addresses != self._addresses:
            self._destroy()
            self._create(spec, device)
        elif versions != self._versions:
            if self._keep:   # converted copies exist (dtype / device mismatch): simplest is a fresh handle
                self._destroy()
                self._create(spec, device)
            else:
                with torch.cuda.device(device):
                    _lib.check(self._lib.cl_net_update(self._handle))
        self._addresses, self._versions = addresses, versions

    # ------------------------------------------------------------------ forward
    def _out(self, b, h, w, device):
        c, ho, wo = _c.c_int(), _c.c_int(), _c.c_int()
        _lib.check(self._lib.cl_net_output_shape(self._handle, b, h, w, _c.byref(c), _c.byref(ho), _c.byref(wo)))
        return torch.empty(b, c.value, ho.value, wo.value, dtype=torch.float32, device=device)

    def _count(self, b, h, w):
        n = _c.c_int()
        _lib.check(self._lib.cl_net_buffers(self._handle, b, h, w, None, None, None, _c.byref(n)))
        self.launches_per_forward = n.value
        self.launches += n.value

    def forward(self, spec, image):
        """image: NCHW fp32 CUDA tensor -> the network output [B, Co, Ho, Wo] (fp32, same device)."""
        if not image.is_cuda:
            raise RuntimeError('crossloc_b200: the coordinate network runs on a CUDA device only (no CPU fallback)')
        if image.size(0) == 0:
            raise RuntimeError('crossloc_b200: empty batch')
        image = image.contiguous().to(torch.float32)
        dev = image.device
        self._sync(spec, dev)
        b, _, h, w = image.shape
        with torch.cuda.device(dev):
            out = self._out(b, h, w, dev)
            _lib.check(self._lib.cl_net_forward(self._handle, image.data_ptr(), b, h, w, out.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream))
        self._count(b, h, w)
        return out

    def forward_frames(self, spec, frames_u8, mean=None, std=None):
        """uint8 HWC frames [B, H, W, C] (CUDA, or pinned / pageable host memory) -> network output.  ToTensor
        [+ Normalize] happen on the device, bit-identical to torchvision (dataloader/dataloader.py:189-212)."""
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4:
            raise RuntimeError('frames must be a uint8 tensor [B, H, W, C]')
        frames_u8 = frames_u8.contiguous()
        dev = frames_u8.device if frames_u8.is_cuda else torch.device('cuda', torch.cuda.current_device())
        self._sync(spec, dev)
        b, h, w, _ = frames_u8.shape
        m = None if mean is None else torch.as_tensor(mean, dtype=torch.float32).contiguous()
        s = None if std is None else torch.as_tensor(std, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            out = self._out(b, h, w, dev)
            _lib.check(self._lib.cl_net_forward_frames(
                self._handle, frames_u8.data_ptr(), b, h, w, None if m is None else m.data_ptr(),
                None if s is None else s.data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        self._count(b, h, w)
        return out

    def wait_fork(self, stream):
        """Make `stream` (a torch.cuda.Stream) wait until the last enqueued forward enters its residual blocks."""
        if self._handle is None:
            return False
        _lib.check(self._lib.cl_net_wait_fork(self._handle, stream.cuda_stream))
        return True

    # ------------------------------------------------------------------ per-op device times
    def set_profiling(self, enable):
        """Eager launches with a CUDA event between ops instead of the graph (bench.py's per-kernel table)."""
        if self._handle is None:
            raise RuntimeError('run one forward first')
        rc = self._lib.cl_net_profile(self._handle, 1 if enable else 0, 0, 0, 0, 0, None, None, None, None, None)
        if rc < 0:
            _lib.check(rc)
        self.profiling = bool(enable)

    def read_profile(self, b, h, w, keep_enabled=True):
        """[(kind, label tuple, flops per launch, total ms, forwards)] for the plan of (b, h, w); clears the record."""
        cap = 256
        kinds = (_c.c_int32 * cap)()
        labels = (_c.c_int32 * (4 * cap))()
        flops = (_c.c_double * cap)()
        ms = (_c.c_float * cap)()
        fw = _c.c_int()
        n = self._lib.cl_net_profile(self._handle, 1 if keep_enabled else 0, b, h, w, cap, kinds, labels, flops, ms, _c.byref(fw))
        if n < 0:
            _lib.check(n)
        self.profiling = bool(keep_enabled)
        return [(OP_KINDS.get(kinds[i], '?'), tuple(labels[4 * i:4 * i + 4]), flops[i], ms[i], fw.value) for i in range(n)]


ENGINE = os.environ.get('CROSSLOC_B200_ENGINE', 'native')   # 'python': the plan of crossloc_b200.cnn (comparison, debugging)
