"""Kernel-time table of one fused training step per data-gradient mode (torch.profiler / CUPTI)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import networks.networks as nets  # noqa: E402
from crossloc_b200 import synth, train_plan  # noqa: E402
from loss.coord import scene_coords_regression_loss  # noqa: E402
from tests.test_loss_cpu import pixel_grid  # noqa: E402

batch = 12
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


def step(backward):
    net.zero_grad()
    pred = train_plan.forward_train(net, images, backward=backward, forward='fp16+fp8', wgrad='fp16x1')
    c, u = torch.split(pred, [3, 1], dim=1)
    loss, _ = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
    loss.backward()


for mode in sys.argv[1:] or ['fp16x3', 'fp16+fp4']:
    step(mode)
    step(mode)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(mode)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(ev.name[:70], [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    total = sum(v[1] for v in rows.values())
    print('==== data gradient %s: %.2f ms of kernels' % (mode, total / 1e3))
    for k, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:16]:
        print('  %-72s %4d  %8.3f ms' % (k, n, us / 1e3))

import time  # noqa: E402
for mode in sys.argv[1:] or ['fp16x3', 'fp16+fp4']:
    for _ in range(2):
        step(mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net.zero_grad()
    pred = train_plan.forward_train(net, images, backward=mode, forward='fp16+fp8', wgrad='fp16x1')
    t1 = time.perf_counter()
    c, u = torch.split(pred, [3, 1], dim=1)
    loss, _ = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print('host enqueue [%s]: forward %.2f ms, loss %.2f ms, backward %.2f ms, drain %.2f ms, total %.2f ms' % (
        mode, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
