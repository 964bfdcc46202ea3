#!/usr/bin/env python
"""Bandwidth showcase of the DSAC* score pass: a full-size 480x720 coordinate map
(345,600 cells, sub-sampling 1) makes the block-per-hypothesis streaming pass genuinely memory-bound.

Prints one JSON line: streamed bytes / time for the whole solve and, from CUDA events around a scoring-only
configuration (refine off), an estimate for the sample + score kernels.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from crossloc_b200 import dsac, synth  # noqa: E402


def main():
    batch, hyps = 4, 256
    dev = torch.device('cuda', 0)
    scenes = [synth.make_scene(i, subsample=1) for i in range(batch)]
    coords = torch.from_numpy(np.stack([s['coords'] for s in scenes])).to(dev)       # [B,3,480,720]
    focal = torch.tensor([s['focal'] for s in scenes], dtype=torch.float32, device=dev)
    pose = torch.zeros(batch, 4, 4, device=dev)
    out = {}
    for refine in (False, True):
        for _ in range(2):
            dsac.forward_rgb_batch(coords, pose, hyps, 10.0, focal, 360.0, 240.0, 100.0, 100.0, 1, seed=1305,
                                   image_base=0, refine=refine)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            dsac.forward_rgb_batch(coords, pose, hyps, 10.0, focal, 360.0, 240.0, 100.0, 100.0, 1, seed=1305,
                                   image_base=0, refine=refine)
        e1.record()
        torch.cuda.synchronize()
        out['ms_refine_%s' % ('on' if refine else 'off')] = e0.elapsed_time(e1) / n
    cells = 480 * 720
    streamed = batch * hyps * cells * 12
    errs = [synth.pose_errors(scenes[b]['pose'], pose[b].cpu().numpy()) for b in range(batch)]
    out.update({'batch': batch, 'hyps': hyps, 'cells': cells, 'streamed_bytes': streamed,
                'sample_plus_score_gbs': streamed / (out['ms_refine_off'] * 1e-3) / 1e9,
                'cell_evals_per_s': batch * hyps * cells / (out['ms_refine_off'] * 1e-3),
                'median_t_err_m': float(np.median([e[0] for e in errs])),
                'note': 'map of 4.15 MB per image is L2 resident: streamed bytes are L2 -> SM traffic, compulsory HBM bytes are '
                        '%d' % (batch * cells * 12)})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
