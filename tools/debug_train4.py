import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
import networks.networks as nets
from crossloc_b200 import train
DEV='cuda'
torch.manual_seed(3)
net = nets.TransPoseNet(torch.tensor([0., 0., 50.]), True, False, 1, 1, 3, 1).to(DEV).train()
x = torch.rand(2, 3, 64, 96, device=DEV)
probe = torch.randn(2, 4, 8, 12, device=DEV)
blk = net.decoder.dec_add_res_block1
keep = {}
h = blk[5].register_forward_hook(lambda m, i, o: keep.update(t=o, c=o.detach().clone()))
net.zero_grad(); out = net.forward_train(x)
torch.cuda.synchronize()
print('after forward: relu5 output changed?', float((keep['t'] - keep['c']).abs().max()))
# step the backward manually through the last conv only
orig_wgrad, orig_dgrad = train.conv_wgrad, train.conv_dgrad
def chk(tag):
    torch.cuda.synchronize(); print(tag, 'relu5 out max abs change', float((keep['t'].detach() - keep['c']).abs().max()), flush=True)
def wg(*a, **k):
    r = orig_wgrad(*a, **k); chk('after wgrad %s' % (tuple(a[2]),)); return r
def dg(*a, **k):
    r = orig_dgrad(*a, **k); chk('after dgrad cout=%d' % a[1].shape[0]); return r
train.conv_wgrad, train.conv_dgrad = wg, dg
(out * probe).sum().backward()
chk('end')
