import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crossloc_b200 import train
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for shape in [(512, 512, 1, 1, 2, 9, 14), (256, 256, 3, 1, 2, 9, 14), (64, 128, 3, 2, 2, 21, 27), (32, 64, 3, 2, 2, 16, 24)]:
    cin, cout, k, stride, b, h, w = shape
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).cuda()
    x = torch.randn(b, cin, h, w, device='cuda').relu().requires_grad_(True)
    y_ref = conv(x)
    gy = torch.randn_like(y_ref) * 1e-4
    gx_ref, gw_ref = torch.autograd.grad(y_ref, (x, conv.weight), gy)
    for name, fn in [('fwd', lambda: (train.conv_forward(x.detach(), conv.weight.detach(), stride) + conv.bias[None, :, None, None], y_ref)),
                     ('dgrad', lambda: (train.conv_dgrad(gy, conv.weight.detach(), (h, w), stride), gx_ref)),
                     ('wgrad', lambda: (train.conv_wgrad(gy, x.detach(), conv.weight.shape, stride), gw_ref))]:
        try:
            out, ref = fn()
            torch.cuda.synchronize()
            print(shape, name, 'rel', rel(out, ref), flush=True)
        except Exception as e:
            print(shape, name, 'FAILED', str(e)[:200], flush=True)
            os._exit(1)
