"""Solver kernel times (cl_dsac_timing) for the refinement cluster size chosen by CROSSLOC_B200_REFINE_CLUSTER."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crossloc_b200 import _lib, dsac, synth
lib = _lib.load()
coords, _, poses, focal = synth.make_batch(100, 32)
c = torch.from_numpy(coords).cuda()
f = torch.from_numpy(focal).cuda()
pose = torch.zeros(32, 4, 4, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    dsac.forward_rgb_batch(c, pose, 256, 10., f, 360., 240., 100., 100., 8, seed=1305, image_base=0)
torch.cuda.synchronize()
_lib.check(lib.cl_dsac_timing(1, stream, None, None))
for i in range(10):
    dsac.forward_rgb_batch(c, pose, 256, 10., f, 360., 240., 100., 100., 8, seed=1305, image_base=32 * i)
torch.cuda.synchronize()
t = (ctypes.c_float * 3)(); n = ctypes.c_int()
_lib.check(lib.cl_dsac_timing(0, stream, t, ctypes.byref(n)))
print('cluster', os.environ.get('CROSSLOC_B200_REFINE_CLUSTER', 'auto'), 'sample %.3f score %.3f refine %.3f ms per 32 frames' % tuple(x / n.value for x in t))
