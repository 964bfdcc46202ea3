#!/usr/bin/env python
"""BASELINE config 4: one train_single_task.py-shaped step (coord MLE loss, forward + backward, Adam) at 480x720.

    python tools/train_step_bench.py [batch] [steps]
Times the native fused plan (crossloc_b200.train_plan), the per-layer native path (crossloc_b200.train) and, for
context, the same step through stock autograd (cuDNN, TF32 allowed as torch's default) on the same GPU.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import networks.networks as nets  # noqa: E402
from crossloc_b200 import synth  # noqa: E402
from loss.coord import scene_coords_regression_loss  # noqa: E402
from tests.test_loss_cpu import pixel_grid  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device('cuda', 0)
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.tensor(synth.NATURESCAPE_MEAN, dtype=torch.float32), False, False, 2, 2, 3, 1).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    coords, gt, poses, focal = synth.make_batch(0, batch)
    images = torch.rand(batch, 3, 480, 720, device=dev)
    gt = torch.from_numpy(gt).to(dev)
    poses = torch.from_numpy(poses).float().to(dev)
    cam = torch.eye(3, device=dev)
    cam[0, 0] = cam[1, 1] = 480.0
    cam[0, 2], cam[1, 2] = 360.0, 240.0
    grid = pixel_grid().to(dev)
    out = {}
    from crossloc_b200 import train_plan
    variants = (('native_fused', lambda t: train_plan.forward_train(net, t, backward='fp16x3', forward='fp16+fp8')),
                ('native_fused_fwd_fp16x3', lambda t: train_plan.forward_train(net, t, backward='fp16x3', forward='fp16x3')),
                ('native_fused_bwd_fp16x1', lambda t: train_plan.forward_train(net, t, backward='fp16x1', forward='fp16+fp8')),
                ('native_layerwise', lambda t: net.forward_train(t, fused=False)),
                ('torch_autograd_cudnn', net.forward_reference))
    for name, fwd in variants:
        times = []
        phases = [0.0, 0.0, 0.0, 0.0]
        for i in range(steps + 1):
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            t0 = time.perf_counter()
            ev[0].record()
            opt.zero_grad()
            pred = fwd(images)
            ev[1].record()
            c, u = torch.split(pred, [3, 1], dim=1)
            loss, rate = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
            ev[2].record()
            loss.backward()
            ev[3].record()
            opt.step()
            ev[4].record()
            torch.cuda.synchronize()
            if i > 0:
                times.append(time.perf_counter() - t0)
                for j in range(4):
                    phases[j] += ev[j].elapsed_time(ev[j + 1])
        out[name] = {'ms_per_step': 1e3 * sum(times) / len(times), 'images_per_s': batch * len(times) / sum(times),
                     'loss': float(loss),
                     'device_ms': dict(zip(('forward', 'loss', 'backward', 'adam'), [p / len(times) for p in phases]))}
    out['batch'] = batch
    out['peak_mem_gb'] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(out))


if __name__ == '__main__':
    main()
