#!/usr/bin/env python
"""One bench-shaped step (32 frames, 256 hypotheses) for ncu captures: `ncu ... python tools/profile_step.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import build_network, synthetic_batch  # noqa: E402
from crossloc_b200.pipeline import Localizer  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device('cuda', 0)
    net = build_network(dev)
    loc = Localizer(net, hyps=256, device=dev)
    images, offsets, _, focal = synthetic_batch(0, 32)
    images, offsets, focal = images.to(dev), offsets.to(dev), focal.to(dev)
    for s in range(steps):
        loc.localize_device(images, focal, offsets, image_base=32 * s)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
