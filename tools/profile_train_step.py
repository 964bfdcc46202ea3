#!/usr/bin/env python
"""A few fused training steps (batch 12, 480x720) for ncu launch lists: `ncu ... python tools/profile_train_step.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import networks.networks as nets  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device('cuda', 0)
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    images = torch.rand(batch, 3, 480, 720, device=dev)
    probe = torch.randn(batch, 4, 60, 90, device=dev)
    for _ in range(steps):
        opt.zero_grad()
        (net.forward_train(images) * probe).sum().backward()
        opt.step()
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
