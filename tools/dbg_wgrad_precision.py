"""Experiment (BASELINE config 4, batch 12 at 480x720): accuracy and time of the weight gradient in one fp16 pass.

Same forward, same data gradients (fp16x3), only the weight-gradient GEMMs change, so the ReLU pattern is identical and the
difference between the two gradient sets is the operand rounding of the weight gradient alone."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import networks.networks as nets  # noqa: E402
from crossloc_b200 import synth, train_plan  # noqa: E402
from loss.coord import scene_coords_regression_loss  # noqa: E402
from tests.test_loss_cpu import pixel_grid  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
FWD = sys.argv[2] if len(sys.argv) > 2 else 'fp16+fp4'
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


def step(backward, wgrad):
    net.zero_grad()
    pred = train_plan.forward_train(net, images, backward=backward, forward=FWD, wgrad=wgrad)
    c, u = torch.split(pred, [3, 1], dim=1)
    loss, _ = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in net.named_parameters()}


def timed(backward, wgrad, n=4):
    step(backward, wgrad)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        step(backward, wgrad)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


ref = step('fp16x3', 'fp16x3')
again = step('fp16x3', 'fp16x3')
scale = max(float(g.double().norm()) for g in ref.values())


def report(tag, g):
    errs = {n: float((g[n].double() - ref[n].double()).norm()) / max(float(ref[n].double().norm()), 1e-4 * scale) for n in ref}
    conv = {n: e for n, e in errs.items() if ref[n].dim() == 4}
    worst = max(conv, key=conv.get)
    tot = (sum(float((g[n].double() - ref[n].double()).norm()) ** 2 for n in ref) ** 0.5) / (sum(float(ref[n].double().norm()) ** 2 for n in ref) ** 0.5)
    print('%-28s all-parameter relative L2 %.3g   worst conv weight %s %.3g   median conv weight %.3g' % (
        tag, tot, worst, conv[worst], sorted(conv.values())[len(conv) // 2]), flush=True)


report('repeat (atomics order)', again)
report('wgrad fp16x1', step('fp16x3', 'fp16x1'))
report('dgrad fp16+fp4, wgrad fp16x3', step('fp16+fp4', 'fp16x3'))
report('dgrad fp16+fp4, wgrad fp16x1', step('fp16+fp4', 'fp16x1'))
report('dgrad + wgrad fp16x1', step('fp16x1', 'fp16x1'))
for bw, wg in (('fp16x3', 'fp16x3'), ('fp16x3', 'fp16x1'), ('fp16+fp4', 'fp16x1'), ('fp16x1', 'fp16x1')):
    print('fwd+loss+bwd  dgrad %s wgrad %s: %.2f ms' % (bw, wg, timed(bw, wg)), flush=True)
