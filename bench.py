#!/usr/bin/env python
"""Headline benchmark: images/sec localized (480x720 RGB -> 6-DoF pose, 256 DSAC* hypotheses).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of 32 synthetic frames per GPU (BASELINE.json configs[2]:
full CNN forward + 256-hypothesis pose solve): stem -> 28 tcgen05 convolutions + GroupNorm -> head -> DSAC*
sample / score / refine.  Weights are random-init TransPoseNet(2+2 extra blocks) under seed 2021; the solver
input is the regressed map plus a synthetic consistent scene.

  value  whole-job images/s with the frames already resident in HBM (device-timed, max over ranks)
  e2e    the same metric through crossloc_b200.pipeline.Localizer with pinned HOST frames: the host-to-device
         copy of every step's frames and the device-to-host read of its poses are inside the timed region
  roofline      the dominant kernel: conv_igemm on the nine 3x3 512->512 layers (tensor-core bound)
  cpu_baseline  the reference's CPU path on this box's host cores (bounded sample)

`--impl reference` times the reference's own CPU implementation of the path: the network as stock PyTorch
ops on the host cores (identical op sequence to /root/reference/networks/networks.py) followed by the C/OpenMP
restatement of dsacstar_rgb_forward (oracle/dsac_oracle.c; the reference extension itself cannot be built
here -- OpenCV C++ is absent).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'images/sec localized (480x720, 256 hyps)'
UNIT = 'images/s'
BATCH, HEIGHT, WIDTH, HYPS = 32, 480, 720, 256
CONV_GFLOP_PER_IMAGE = 295.413


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--hyps', type=int, default=HYPS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def build_network(device):
    import networks.networks as nets
    torch.manual_seed(2021)   # the reference scripts' seed (test_single_task.py:265)
    net = nets.TransPoseNet(torch.zeros(3), False, False, enc_add_res_block=2, dec_add_res_block=2,
                            num_task_channel=3, num_pos_channel=1)
    return net.eval().to(device)


def synthetic_batch(first_index, batch):
    from crossloc_b200 import synth
    coords, _, poses, focal = synth.make_batch(first_index, batch)
    g = torch.Generator().manual_seed(first_index)
    images = torch.rand(batch, 3, HEIGHT, WIDTH, generator=g)
    return images, torch.from_numpy(coords), poses, torch.from_numpy(focal)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU reference arm
class CpuPath:
    """The all-CPU path: stock torch network on the host cores + C/OpenMP DSAC* (batch size 1, as the reference)."""

    def __init__(self, hyps, threads, frames):
        from oracle import dsac_oracle_c as tier2
        self.tier2 = tier2
        self.hyps = hyps
        torch.set_num_threads(threads)
        self.net = build_network('cpu')
        self.images, self.offsets, _, self.focal = synthetic_batch(5000, frames)
        with torch.no_grad():   # warm-up: thread pools, oneDNN primitive caches
            self.net.forward_reference(self.images[:1])

    def rate(self, frames):
        """images/s over `frames` frames, plus the split between network and solver."""
        t0 = time.perf_counter()
        t_net = t_solve = 0.0
        for i in range(frames):   # utils/evaluation.py:69 evaluates with batch size 1
            j = i % self.images.size(0)
            t1 = time.perf_counter()
            with torch.no_grad():
                pred = self.net.forward_reference(self.images[j:j + 1])
            coords = (pred[:, :3] + self.offsets[j:j + 1]).contiguous().numpy()
            t2 = time.perf_counter()
            self.tier2.forward_rgb(coords[0], self.hyps, 10.0, float(self.focal[j]), WIDTH / 2, HEIGHT / 2, 100.0,
                                   100.0, 8, seed=1305, image=5000 + j)
            t3 = time.perf_counter()
            t_net += t2 - t1
            t_solve += t3 - t2
        total = time.perf_counter() - t0
        return frames / total, {'network_s_per_image': t_net / frames, 'solver_s_per_image': t_solve / frames}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return   # one CPU path per box: the other ranks exit without work
    threads = len(os.sched_getaffinity(0))
    os.environ['OMP_NUM_THREADS'] = str(threads)   # torchrun pins it to 1; the CPU arm may use every host core
    per_step = 2   # bounded sample: two frames per step keeps --steps 10 --warmup 3 within a few minutes
    path = CpuPath(args.hyps, threads, per_step)
    rates = []
    detail = {}
    for step in range(args.warmup + args.steps):
        rate, detail = path.rate(per_step)
        if step >= args.warmup:
            rates.append(rate)
    value = per_step * len(rates) / sum(per_step / r for r in rates)   # frames / total time of the timed steps
    sample = '%d frames per step, batch size 1, %d hypotheses' % (per_step, args.hyps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * per_step / value, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'batch32_480x720_forward+dsac256', 'network': 'TransPoseNet enc+2/dec+2, random init seed 2021',
                   'reference_path': 'stock torch ops on CPU + C/OpenMP restatement of dsacstar_rgb_forward'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample, **detail},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch.distributed as dist
    from crossloc_b200.pipeline import Localizer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: the native arm needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    B = args.batch

    net = build_network(dev)
    loc = Localizer(net, hyps=args.hyps, device=dev)
    # distinct synthetic frames per rank: rank r localizes images r*B .. r*B+B-1 of every step (weak scaling)
    images_h, offsets_h, gt_poses, focal_h = synthetic_batch(rank * B, B)
    images_h = images_h.pin_memory()
    images_d = images_h.to(dev)
    offsets_d = offsets_h.to(dev)
    focal_d = focal_h.to(dev)
    gathered = torch.empty(world * B, 16, dtype=torch.float32, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(image_base):
        # the solve (and the pose gather behind it) run on the localizer's solver stream and overlap the next step's CNN
        pose = loc.localize_device(images_d, focal_d, offsets_d, image_base=image_base, overlap=True)
        if world > 1:   # trivial pose gather: 2 KB per rank
            with torch.cuda.stream(loc.solver_stream):
                dist.all_gather_into_tensor(gathered, pose.reshape(B, 16))
                loc.solver_done.record(loc.solver_stream)
        return pose

    # ---- device-resident throughput ("value")
    engine_launches0 = 0
    for w in range(args.warmup):
        step_device(w * world * B + rank * B)
    barrier()
    engine = net._engine
    engine.events = []
    engine_launches0 = engine.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for k in range(args.steps):
        # no explicit L2 flush: one step streams >1 GB of activations through the 126 MB L2 (config.l2)
        pose = step_device((args.warmup + k) * world * B + rank * B)
    torch.cuda.current_stream().wait_event(loc.solver_done)   # the last solve is inside the timed region
    stop.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    launches_per_step = (engine.launches - engine_launches0) / args.steps + 3   # + DSAC sample / score / refine
    conv_events = engine.events
    engine.events = None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- accuracy guard on this rank's last batch
    from crossloc_b200 import synth
    pose_np = pose.cpu().numpy()
    errs = np.array([synth.pose_errors(gt_poses[b], pose_np[b]) for b in range(B)])

    # ---- roofline of the dominant kernel: 3x3 512->512 convolutions
    peaks = {}
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
    peak_tf, peak_src = (peaks.get('bf16_tflops_sustained'), 'MEASURED_PEAKS.json bf16_tflops_sustained') \
        if peaks.get('bf16_tflops_sustained') else (1400.0, 'B200_PROFILING.md fallback (sustained)')
    dom = [(fl, e0.elapsed_time(e1)) for (name, shape, fl, e0, e1) in conv_events if shape == (512, 512, 3, 1)]
    by_shape = {}
    for (name, shape, fl, e0, e1) in conv_events:
        key = '/'.join(str(x) for x in shape)
        ent = by_shape.setdefault(key, [0, 0.0, fl])
        ent[0] += 1
        ent[1] += e0.elapsed_time(e1)
    kernel_ms = {k: {'launches_per_step': v[0] / args.steps, 'ms_per_step': v[1] / args.steps,
                     'avg_ms': v[1] / v[0], 'tflops_useful': (v[2] / (v[1] / v[0] * 1e-3) / 1e12) if v[2] else None}
                 for k, v in by_shape.items()}
    conv_total_ms = sum(e0.elapsed_time(e1) for (_, shape, _, e0, e1) in conv_events if len(shape) == 4 and shape[0] != 'gn_apply') / args.steps
    roofline = None
    if dom:
        avg_ms = sum(ms for _, ms in dom) / len(dom)
        achieved = dom[0][0] / (avg_ms * 1e-3) / 1e12
        nterms = engine.nterms
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (same shape and mode)
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'r1s3_conv3x3_pair_ncu_full.json')
        if os.path.exists(prof) and B == BATCH and engine.precision == 'fp16+fp8':
            traffic = json.load(open(prof)).get('dram_bytes_per_launch')
        roofline = {'bound': 'tensor', 'kernel': 'conv_igemm_pair_kernel<64> 3x3 512->512 @60x90 x%d images' % B,
                    'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                    'traffic': traffic, 'traffic_unit': 'bytes per launch (dram read + write, ncu)',
                    'peak_source': peak_src, 'avg_launch_ms': avg_ms, 'launches_timed': len(dom),
                    'issued_tflops': achieved * nterms, 'issued_frac': achieved * nterms / peak_tf,
                    'note': 'achieved counts algorithmic FLOPs (2*pixels*Cout*Cin*9); %s issues %d fp16-MMA equivalents per '
                            'product (fp16+fp8: one fp16 MMA + two e4m3 MMAs at twice the rate)' % (engine.precision, nterms),
                    'all_conv_ms_per_step': conv_total_ms}

    # ---- end to end through the public API with host buffers ("e2e")
    barrier()
    for w in range(2):
        loc.submit(images_h, focal_d, offsets_d, image_base=w * B)
    loc.result()
    loc.result()
    barrier()
    t0 = time.perf_counter()
    loc.submit(images_h, focal_d, offsets_d, image_base=0)
    for k in range(1, args.steps):
        loc.submit(images_h, focal_d, offsets_d, image_base=k * world * B + rank * B)
        loc.result()
    last = loc.result()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    assert np.isfinite(last.numpy()).all()

    line = None
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            os.environ['OMP_NUM_THREADS'] = str(threads)
            sample = 6
            rate, detail = CpuPath(args.hyps, threads, sample).rate(sample)
            cpu_baseline = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                            'sample': '%d frames, batch size 1, %d hypotheses: stock torch network on the host cores + '
                                      'oracle/dsac_oracle.c (OpenMP)' % (sample, args.hyps), **detail}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16+f8 split products -> f32 accumulate (conv) + f64 (pose solve)', 'data': 'synthetic',
            'config': {'workload': 'batch32_480x720_forward+dsac256', 'batch_per_gpu': B, 'hypotheses': args.hyps,
                       'network': 'TransPoseNet enc+2/dec+2, random init seed 2021', 'conv_precision': engine.precision,
                       'parallelism': 'dp%d (images sharded, NCCL all-gather of poses)' % world,
                       'l2': 'one step streams >1 GB of activations (L2 is 126 MB); inputs not re-used between steps'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(images_h.numel() * 4),
                    'd2h_bytes_per_step': int(B * 16 * 4)},
            'gpu_launches': int(round(launches_per_step * args.steps)),
            'gpu_launches_per_step': launches_per_step,
            'roofline': roofline,
            'cpu_baseline': cpu_baseline,
            'conv_tflops_useful': B * CONV_GFLOP_PER_IMAGE / conv_total_ms if conv_total_ms else None,
            'pose_error_median': {'t_m': float(np.median(errs[:, 0])), 'r_deg': float(np.median(errs[:, 1]))},
            'kernel_ms': kernel_ms,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there) get stderr instead."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    args = parse_args()
    _protect_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
