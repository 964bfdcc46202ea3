"""Gradient accuracy of the fused training plan against stock autograd in fp32 (TF32 off) on full-size frames, next to the
error stock autograd itself makes with TF32 on (torch's default for convolutions): one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import networks.networks as nets  # noqa: E402
from crossloc_b200 import synth, train_plan  # noqa: E402
from loss.coord import scene_coords_regression_loss  # noqa: E402
from tests.test_loss_cpu import pixel_grid  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda', 0)
torch.manual_seed(2021)
net = nets.TransPoseNet(torch.tensor(synth.NATURESCAPE_MEAN, dtype=torch.float32), False, False, 2, 2, 3, 1).to(dev).train()
coords, gt, poses, focal = synth.make_batch(0, batch)
images = torch.rand(batch, 3, 480, 720, device=dev)
gt = torch.from_numpy(gt).to(dev)
poses = torch.from_numpy(poses).float().to(dev)
cam = torch.eye(3, device=dev)
cam[0, 0] = cam[1, 1] = 480.0
cam[0, 2], cam[1, 2] = 360.0, 240.0
grid = pixel_grid().to(dev)


def grads(forward):
    net.zero_grad()
    pred = forward(images)
    c, u = torch.split(pred, [3, 1], dim=1)
    loss, _ = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
    loss.backward()
    return float(loss), {n: p.grad.detach().double().clone() for n, p in net.named_parameters()}


def compare(g, ref):
    tot = (sum(float((g[n] - ref[n]).norm()) ** 2 for n in ref) ** 0.5) / (sum(float(ref[n].norm()) ** 2 for n in ref) ** 0.5)
    scale = max(float(v.norm()) for v in ref.values())
    per = sorted(float((g[n] - ref[n]).norm()) / max(float(ref[n].norm()), 1e-4 * scale) for n in ref if ref[n].dim() == 4)
    return {'all_parameters_rel_l2': tot, 'conv_weight_median': per[len(per) // 2], 'conv_weight_worst': per[-1]}


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
loss_ref, ref = grads(net.forward_reference)
torch.backends.cudnn.allow_tf32 = True
loss_tf32, g_tf32 = grads(net.forward_reference)
out = {'batch': batch, 'loss_fp32_autograd': loss_ref,
       'stock_autograd_tf32': dict(compare(g_tf32, ref), loss_rel=abs(loss_tf32 - loss_ref) / abs(loss_ref))}
for name, kw in (('native_default', {}), ('native_fp16x3_everywhere', dict(backward='fp16x3', forward='fp16x3', wgrad='fp16x3')),
                 ('native_tf32_grade', dict(backward='fp16x1'))):
    object.__setattr__(net, '_train_plan', None)
    loss, g = grads(lambda t: train_plan.forward_train(net, t, **kw))
    out[name] = dict(compare(g, ref), loss_rel=abs(loss - loss_ref) / abs(loss_ref))
print(json.dumps(out))
