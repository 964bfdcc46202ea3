"""Debug build only (CL_DEBUG_TRAP=1 python -m crossloc_b200.build --force): where the fp4 convolution kernel waits."""
import ctypes, os, sys, torch
sys.path.insert(0, os.getcwd())
from crossloc_b200 import _lib
from tests import test_cnn_gpu as T
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
for shape in [(512, 512, 1, 1, 32, 60, 90), (512, 512, 3, 1, 32, 60, 90)]:
    cin, cout, k, stride, b, h, w = shape
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).cuda()
    x = torch.randn(b, cin, h, w, device='cuda').relu()
    T.run_conv_fp4(x, conv, 32)
    lib.cl_debug_counters(buf, 1)
    T.run_conv_fp4(x, conv, 32)
    lib.cl_debug_counters(buf, 1)
    v = list(buf)
    tiles = max(v[7], 1)
    print(shape, 'tiles(all leaders)=%d' % v[7], 'per tile [cycles]: ovl=%.0f full_p0=%.0f sf=%.0f full_p1=%.0f | epi(w4) tfull=%.0f wait_read=%.0f | mma loop=%.0f | epi(w4): tmem_ld=%.0f store path=%.0f stats=%.0f | scale loader: slot wait=%.0f tile wait=%.0f' % (
        v[0] / tiles, v[1] / tiles, v[2] / tiles, v[3] / tiles, v[4] / (2 * tiles), v[5] / (2 * tiles), v[6] / tiles,
        v[8] / (2 * tiles), v[9] / (2 * tiles), v[10] / (2 * tiles), v[11] / tiles, v[12] / tiles))
