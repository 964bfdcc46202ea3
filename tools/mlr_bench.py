#!/usr/bin/env python
"""MLR CrossLoc model end to end: TransPoseNet(num_mlr=3) -- three encoders, the merge block,
one decoder (68.4 M parameters) -- on 480x720 frames + 256-hypothesis DSAC*.  One JSON line: device times of the network
and of the whole localization, the stock-torch forward (cuDNN, TF32 allowed) for context, parity vs fp32 torch.

    python tools/mlr_bench.py [batch] [num_mlr]
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
from tools.fullsize_bench import timed  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    num_mlr = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device('cuda', 0)
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1, num_mlr=num_mlr).eval().to(dev)
    images = torch.rand(batch, 3, 480, 720, generator=torch.Generator().manual_seed(0)).to(dev)
    coords, _, poses, focal = synth.make_batch(0, batch)
    offsets, focal_d = torch.from_numpy(coords).to(dev), torch.from_numpy(focal).to(dev)
    loc = Localizer(net, hyps=256, device=dev)
    with torch.no_grad():
        ms_net = timed(lambda: net(images), 5)
        ms_all = timed(lambda: loc.localize_device(images, focal_d, offsets, image_base=0), 5)
        torch.backends.cudnn.allow_tf32 = True
        ms_torch = timed(lambda: net.forward_reference(images), 3)
        torch.backends.cudnn.allow_tf32 = False
        ref = net.forward_reference(images[:2])
        out = net(images[:2])
    pose = loc.localize_device(images, focal_d, offsets, image_base=0)
    torch.cuda.synchronize()
    errs = np.array([synth.pose_errors(poses[b], pose[b].cpu().numpy()) for b in range(batch)])
    print(json.dumps({
        'workload': 'MLR model, %d encoders, batch %d, 480x720, 256 hypotheses' % (num_mlr, batch),
        'parameters': sum(p.numel() for p in net.parameters()),
        'network_ms': ms_net, 'localize_ms': ms_all, 'images_per_s': batch / ms_all * 1e3,
        'stock_torch_cudnn_tf32_network_ms': ms_torch,
        'coord_rel_err_vs_fp32_torch': float((out[:, :3] - ref[:, :3]).norm() / ref[:, :3].norm()),
        'median_t_err_m': float(np.median(errs[:, 0])), 'median_r_err_deg': float(np.median(errs[:, 1]))}))


if __name__ == '__main__':
    main()
