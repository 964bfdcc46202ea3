import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import networks.networks as nets
from loss.coord import scene_coords_regression_loss
from tests.test_loss_cpu import pixel_grid
from crossloc_b200 import train
DEV='cuda'
torch.manual_seed(3)
net = nets.TransPoseNet(torch.tensor([0., 0., 50.]), True, False, 1, 1, 3, 1).to(DEV).train()
x = torch.rand(2, 3, 64, 96, device=DEV)
gt = torch.randn(2, 3, 8, 12, device=DEV) * 5 + torch.tensor([0., 0., 50.], device=DEV)[None, :, None, None]
pose = torch.eye(4, device=DEV).repeat(2, 1, 1)
cam = torch.eye(3, device=DEV); cam[0, 0] = cam[1, 1] = 60.0; cam[0, 2], cam[1, 2] = 48.0, 32.0
probe = torch.randn(2, 4, 8, 12, device=DEV)
def step(forward):
    net.zero_grad()
    out = forward(x)
    loss = (out * probe).sum()
    loss.backward()
    return loss.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}
loss_ref, g_ref = step(net.forward_reference)
loss_ref2, g_ref2 = step(net.forward_reference)
loss_nat, g_nat = step(net.forward_train)
scale = max(float(g.double().norm()) for g in g_ref.values())
def err(a,b,n): return float((a[n].double() - b[n].double()).norm()) / max(float(b[n].double().norm()), 1e-4 * scale)
print('loss', float(loss_ref), float(loss_nat))
rows = sorted(((err(g_nat,g_ref,n), err(g_ref2,g_ref,n), n, float(g_ref[n].norm())) for n in g_ref), reverse=True)
for n in g_ref:
    if n.endswith('weight') and g_ref[n].dim()==4: print('%.3e  %s  |g|=%.3e' % (err(g_nat,g_ref,n), n, float(g_ref[n].norm())))
# single-layer check in isolation with the same shapes
conv = net.decoder.dec_add_res_block1[0]
xin = torch.randn(2, 128, 8, 12, device=DEV).relu().requires_grad_(True)
y = conv(xin); gy = torch.randn_like(y)
gx_ref, gw_ref = torch.autograd.grad(y, (xin, conv.weight), gy)
gw = train.conv_wgrad(gy, xin.detach(), conv.weight.shape, 1); gx = train.conv_dgrad(gy, conv.weight.detach(), (8,12), 1)
print('isolated wgrad rel', float((gw-gw_ref).norm()/gw_ref.norm()), 'dgrad', float((gx-gx_ref).norm()/gx_ref.norm()))
# capture the real upstream gradient at dec_add_res_block1[6] and compare dgrad on it
conv6 = net.decoder.dec_add_res_block1[6]
cap = {}
h = conv6.register_full_backward_hook(lambda m, gi, go: cap.update(go=go[0].detach().clone(), gi=gi[0].detach().clone()))
net.zero_grad(); out = net.forward_reference(x); ((out * probe).sum()).backward(); h.remove()
gy = cap['go']; gx_ref = cap['gi']
gx = train.conv_dgrad(gy.contiguous(), conv6.weight.detach(), (8, 12), 1)
print('real-gy dgrad rel', float((gx - gx_ref).norm() / gx_ref.norm()), 'gy amax', float(gy.abs().max()), 'gy rms', float(gy.pow(2).mean().sqrt()), 'contig', gy.is_contiguous())
import torch.nn.functional as F
gx_t = torch.nn.grad.conv2d_input(gx_ref.shape, conv6.weight, gy, stride=1, padding=1)
print('torch conv2d_input vs hook', float((gx_t - gx_ref).norm() / gx_ref.norm()))
s = train._amax_scale(gy); print('scale', float(s), 'max scaled', float((gy*s).abs().max()))
q = (gy * s); hi = q.half().float(); lo = (q - hi).half().float(); print('split recon rel err', float((hi + lo - q).norm() / q.norm()), 'frac hi==0', float((hi == 0).float().mean()))
