#!/usr/bin/env python
"""BASELINE config 5: a naturescape-shaped synthetic set localized across the GPUs of one box.

    python tools/eval_synthetic.py --images 2048 --hyps 256
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/eval_synthetic.py

Rank r localizes images {i : i mod world == r} in batches of 32 (CNN forward + DSAC*), one NCCL all-gather
collects [pose(16), t_err, r_err] per image and rank 0 prints the reference's accuracy summary
(/root/reference/utils/evaluation.py:212-230) as one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import HEIGHT, WIDTH, build_network  # noqa: E402
from crossloc_b200 import parallel, synth  # noqa: E402
from crossloc_b200.pipeline import Localizer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=2048)
    ap.add_argument('--hyps', type=int, default=256)
    ap.add_argument('--batch', type=int, default=32)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    net = build_network(dev)
    loc = Localizer(net, hyps=args.hyps, device=dev)
    scenes = {}

    def localize(indices):
        batch = [synth.make_scene(i) for i in indices]
        for i, s in zip(indices, batch):
            scenes[i] = s['pose']
        g = torch.Generator().manual_seed(indices[0])
        images = torch.rand(len(indices), 3, HEIGHT, WIDTH, generator=g).pin_memory()
        offsets = torch.stack([torch.from_numpy(s['coords']) for s in batch]).to(dev)
        focal = torch.tensor([s['focal'] for s in batch], dtype=torch.float32, device=dev)
        return loc.localize(images, focal, offsets, image_base=indices[0]).clone()

    def gt(i):
        return scenes[i] if i in scenes else synth.make_scene(i)['pose']

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows, summary = parallel.evaluate_sharded(localize, gt, args.images, args.batch, rank, world, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        summary.update({'images': args.images, 'world': world, 'hyps': args.hyps, 'wall_s': dt,
                        'note': 'wall time includes synthetic scene generation on the host'})
        print(json.dumps(summary))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
