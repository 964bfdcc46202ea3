"""world_size-2 gloo test of the N>1 path: image sharding + pose gather + the reference's summary statistics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crossloc_b200 import parallel, synth

N_TOTAL, HYPS = 7, 16   # odd count: ranks get 4 and 3 images (ragged shards)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _localize(indices):
    """Stand-in for the GPU localizer on a CPU-only box: the tier-2 oracle acts as the solver (tests may use it)."""
    from oracle import dsac_oracle_c as tier2
    poses = []
    for i in indices:
        s = synth.make_scene(i, height=96, width=144, focal=120.0)
        o = tier2.forward_rgb(s['coords'], HYPS, 10.0, 120.0, 72.0, 48.0, 100.0, 100.0, 8, seed=1305, image=i)
        poses.append(torch.from_numpy(o['pose']))
    return torch.stack(poses)


def _gt(i):
    return synth.make_scene(i, height=96, width=144, focal=120.0)['pose']


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rows, summary = parallel.evaluate_sharded(_localize, _gt, N_TOTAL, batch=2, rank=rank, world=world)
    np.save(os.path.join(out_dir, 'rows%d.npy' % rank), rows.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(tmp_path):
    assert parallel.shard_indices(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_indices(7, 1, 2) == [1, 3, 5]
    single, summary = parallel.evaluate_sharded(_localize, _gt, N_TOTAL, batch=3)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / 'rows0.npy')
    r1 = np.load(tmp_path / 'rows1.npy')
    assert np.array_equal(r0, r1)                    # every rank holds the full table
    assert np.array_equal(r0, single.numpy())        # sharded == unsharded, in global image order
    assert summary['count'] == N_TOTAL and summary['median_t_m'] < 5.0


def test_summary_matches_reference_buckets():
    t = np.array([1.0, 4.0, 12.0, 25.0])
    r = np.array([1.0, 6.0, 8.0, 9.0])
    s = parallel.summarize(t, r)
    assert s['30m10deg'] == 100.0 and s['20m10deg'] == 75.0 and s['10m10deg'] == 50.0
    assert s['10m7deg'] == 50.0 and s['5m5deg'] == 25.0 and s['3m3deg'] == 25.0
    assert s['median_t_m'] == 8.0 and s['median_r_deg'] == 7.0
