import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import networks.networks as nets
from crossloc_b200 import train
DEV='cuda'
torch.manual_seed(3)
net = nets.TransPoseNet(torch.tensor([0., 0., 50.]), True, False, 1, 1, 3, 1).to(DEV).train()
x = torch.rand(2, 3, 64, 96, device=DEV)
probe = torch.randn(2, 4, 8, 12, device=DEV)
blk = net.decoder.dec_add_res_block1
def run(forward):
    caps = {}
    hooks = [m.register_full_backward_hook(lambda m, gi, go, i=i: caps.__setitem__(i, (go[0].detach().clone(), None if gi[0] is None else gi[0].detach().clone()))) for i, m in enumerate(blk) if not isinstance(m, torch.nn.Conv2d)]
    fw = {}
    fhooks = [m.register_forward_hook(lambda m, inp, out, i=i: fw.__setitem__(i, out.detach().clone())) for i, m in enumerate(blk)]
    net.zero_grad(); out = forward(x); (out * probe).sum().backward()
    for h in hooks + fhooks: h.remove()
    return caps, fw
c_ref, f_ref = run(net.forward_reference)
c_nat, f_nat = run(net.forward_train)
def rel(a, b): return float((a - b).norm() / b.norm())
for i in sorted(f_ref): print('fwd out of block[%d] %s rel %.2e' % (i, type(blk[i]).__name__, rel(f_nat[i], f_ref[i])) if i in f_nat else 'fwd %d missing in native (conv via function)' % i)
for i in sorted(c_ref, reverse=True):
    go_r, gi_r = c_ref[i]; go_n, gi_n = c_nat[i]
    print('bwd block[%d] %s: grad_out rel %.2e  grad_in rel %.2e' % (i, type(blk[i]).__name__, rel(go_n, go_r), rel(gi_n, gi_r)))
go_r, gi_r = c_ref[5]; go_n, gi_n = c_nat[5]
m_r = f_ref[5] > 0; m_n = f_nat[5] > 0
print('mask mismatches', int((m_r != m_n).sum()), 'of', m_r.numel())
print('ref consistency', rel(go_r * m_r, gi_r), 'nat consistency', rel(go_n * m_n, gi_n), 'nat with ref mask', rel(go_n * m_r, gi_r))
pre_r, pre_n = f_ref[4], f_nat[4]
d = (pre_r - pre_n).abs()
print('pre-activation abs diff max %.3e mean %.3e; |pre| quantiles' % (float(d.max()), float(d.mean())), [float(q) for q in torch.quantile(pre_r.abs().flatten(), torch.tensor([0.001, 0.01, 0.1, 0.5], device=DEV))])
