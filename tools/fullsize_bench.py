#!/usr/bin/env python
"""Full-size variant end to end: TransPoseNet(full_size_output=True) on 480x720 frames
-> [B,4,480,720] map (OUTPUT_SUBSAMPLE = 1) -> DSAC* over 345,600 cells.  Prints one JSON line with device times of
the network (fused DUC head included) and of the whole localization, plus the stock-torch forward for context.

    python tools/fullsize_bench.py [batch] [hyps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import networks.networks as nets  # noqa: E402
from crossloc_b200 import synth  # noqa: E402
from crossloc_b200.pipeline import Localizer  # noqa: E402


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    hyps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device('cuda', 0)
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1, full_size_output=True).eval().to(dev)
    images = torch.rand(batch, 3, 480, 720, generator=torch.Generator().manual_seed(0)).to(dev)
    scenes = [synth.make_scene(i, subsample=1) for i in range(batch)]
    offsets = torch.from_numpy(np.stack([s['coords'] for s in scenes])).to(dev)
    focal = torch.tensor([s['focal'] for s in scenes], dtype=torch.float32, device=dev)
    loc = Localizer(net, hyps=hyps, device=dev)
    with torch.no_grad():
        ms_net = timed(lambda: net(images), 5)
        ms_all = timed(lambda: loc.localize_device(images, focal, offsets, image_base=0), 3)
        torch.backends.cudnn.allow_tf32 = True
        ms_torch = timed(lambda: net.forward_reference(images), 3)
        torch.backends.cudnn.allow_tf32 = False
        ref = net.forward_reference(images[:2])
        out = net(images[:2])
    pose = loc.localize_device(images, focal, offsets, image_base=0)
    torch.cuda.synchronize()
    errs = np.array([synth.pose_errors(scenes[b]['pose'], pose[b].cpu().numpy()) for b in range(batch)])
    print(json.dumps({
        'workload': 'full-size 480x720 map, batch %d, %d hypotheses, 345600 cells' % (batch, hyps),
        'network_ms': ms_net, 'localize_ms': ms_all, 'images_per_s': batch / ms_all * 1e3,
        'stock_torch_cudnn_tf32_network_ms': ms_torch,
        'coord_rel_err_vs_fp32_torch': float((out[:, :3] - ref[:, :3]).norm() / ref[:, :3].norm()),
        'median_t_err_m': float(np.median(errs[:, 0])), 'median_r_err_deg': float(np.median(errs[:, 1]))}))


if __name__ == '__main__':
    main()
