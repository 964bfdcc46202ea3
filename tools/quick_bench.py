import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networks.networks as nets
from crossloc_b200.cnn import CoordNetEngine
torch.manual_seed(2021)
net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1).eval().cuda()
for prec in ['fp16x3', 'fp16x1']:
    eng = CoordNetEngine(precision=prec)
    for B in [1, 8, 32]:
        x = torch.rand(B, 3, 480, 720, device='cuda')
        with torch.no_grad():
            for _ in range(3): eng.forward(net._spec(), x)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.time(); e0.record()
            n = 5
            for _ in range(n): eng.forward(net._spec(), x)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
        print(prec, 'B', B, '%.2f ms/batch  %.1f img/s  %.1f TFLOP/s  (wall %.2f ms)' % (ms, B / ms * 1e3, B * 295.413 / ms, (time.time() - t0) / n * 1e3), flush=True)
# stock torch for context
torch.backends.cudnn.allow_tf32 = True
x = torch.rand(32, 3, 480, 720, device='cuda')
with torch.no_grad():
    for _ in range(2): net.forward_reference(x)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3): net.forward_reference(x)
    torch.cuda.synchronize(); print('torch(cudnn tf32) B32 %.2f ms' % ((time.time() - t0) / 3 * 1e3))
