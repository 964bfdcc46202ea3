"""Debug build only (CL_DEBUG_TRAP=1 python -m crossloc_b200.build --force): where the CTA-pair convolution waits on the narrow
strided layers (conv2: 32 -> 64, conv3: 64 -> 128, conv4: 128 -> 256; all 3x3 stride 2)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.getcwd())
from crossloc_b200 import _lib
from tests import test_cnn_gpu as T
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
for shape in [(32, 64, 3, 2, 32, 480, 720), (64, 128, 3, 2, 32, 240, 360), (128, 256, 3, 2, 32, 120, 180)]:
    cin, cout, k, stride, b, h, w = shape
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).cuda()
    x = torch.randn(b, cin, h, w, device='cuda').relu()
    T.run_conv(x, conv, 3, 32)
    lib.cl_debug_counters(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T.run_conv(x, conv, 3, 32)
    lib.cl_debug_counters(buf, 1)
    v = list(buf)
    tiles = max(v[7], 1)
    print(shape, 'tiles(all leaders)=%d' % v[7], 'per tile [cycles]: mma loop=%.0f (tempty wait %.0f, full wait %.0f) | producer empty wait %.0f | '
          'epilogue(w4 leader): tfull wait %.0f, busy %.0f' % (v[6] / tiles, v[0] / tiles, v[1] / tiles, v[2] / tiles, v[4] / tiles, v[5] / tiles))
