#!/usr/bin/env python
"""Headline benchmark: images/sec localized (480x720 RGB -> 6-DoF pose, 256 DSAC* hypotheses).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload localize|train|eval5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of 32 synthetic frames per GPU (BASELINE.json configs[2]:
full CNN forward + 256-hypothesis pose solve): stem -> 28 tcgen05 convolutions + GroupNorm -> head -> DSAC*
sample / score / refine.  Weights are random-init TransPoseNet(2+2 extra blocks) under seed 2021; the solver
input is the regressed map plus a synthetic consistent scene.

  value          whole-job images/s with the frames already resident in HBM (device-timed, max over ranks); the CNN runs
                 through the C++ runtime (cl_net_forward) in its per-op event mode so that every kernel of the timed
                 region is measured, the solve of batch k overlaps the residual blocks of batch k+1
  e2e            the same metric through crossloc_b200.pipeline.Localizer with pinned HOST frames (CUDA-graph replay):
                 the host-to-device copy of every step's frames and the device-to-host read of its poses are timed
  roofline       the dominant kernel: conv_igemm on the nine 3x3 512->512 layers (tensor-core bound)
  roofline_score the DSAC* scoring pass (streamed map bytes vs the HBM peak, cells/s)
  cpu_baseline   the reference's CPU path on this box's host cores (bounded sample)
  gpu_stock_baseline  SURVEY 8d "baseline B": stock torch / cuDNN forward on this GPU + .cpu() + the CPU solver
  parity         coordinate map vs fp32 torch on the timed weights; poses vs the CPU oracle on the timed frames
  latency_batch1 the reference's own call pattern (batch 1, CPU tensors in and out), per-frame latency
  config4_train / config5_eval   BASELINE configs 4 and 5 measured in the same run (also --workload train / eval5)

`--impl reference` times the reference's own CPU implementation of the path: the network as stock PyTorch
ops on the host cores (identical op sequence to /root/reference/networks/networks.py) followed by the C/OpenMP
restatement of dsacstar_rgb_forward (oracle/dsac_oracle.c; the reference extension itself cannot be built
here -- OpenCV C++ is absent).
"""
import argparse
import ctypes
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
TRAIN_BATCH = 12                 # script_clean_training/encoder_pretrain.sh:6
EVAL5_IMAGES_PER_GPU = 256       # x 8 GPUs = the 2048-image set of BASELINE config 5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='localize', choices=['localize', 'train', 'eval5'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--hyps', type=int, default=HYPS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='headline numbers only (no baselines, parity, configs 4/5)')
    return ap.parse_args()


def config_of(args, world):
    """Workload description shared verbatim by both arms (the driver compares the two lines)."""
    return {'workload': 'batch%d_480x720_forward+dsac%d' % (args.batch, args.hyps), 'batch_per_gpu': args.batch,
            'hypotheses': args.hyps, 'height': HEIGHT, 'width': WIDTH,
            'network': 'TransPoseNet enc+2/dec+2, random init seed 2021',
            'parallelism': 'dp%d (images sharded, pose gather)' % world,
            'l2': 'inputs larger than L2: one step streams >1 GB of activations (L2 is 126 MB), no re-use between steps'}


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


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    p = json.load(open(path)) if os.path.exists(path) else {}
    tf = (p.get('bf16_tflops_sustained'), 'MEASURED_PEAKS.json bf16_tflops_sustained (measured)') if p.get('bf16_tflops_sustained') \
        else (1400.0, 'B200_PROFILING.md fallback (sustained)')
    bw = (p.get('hbm_gbs'), 'MEASURED_PEAKS.json hbm_gbs (measured)') if p.get('hbm_gbs') else (6650.0, 'B200_PROFILING.md fallback')
    return tf, bw


# arithmetic of the path per convolution scheme (crossloc_b200.cnn.PRECISION)
DTYPES = {
    'fp16+fp4': 'f16 products + block-scaled e2m1 (fp4) correction products -> f32 accumulate (conv) + f64 (pose solve)',
    'fp16+fp8': 'f16 products + e4m3 (fp8) correction products -> f32 accumulate (conv) + f64 (pose solve)',
    'fp16x3': 'f16 split products (three per product) -> f32 accumulate (conv) + f64 (pose solve)',
    'fp16x1': 'f16 products -> f32 accumulate (conv) + f64 (pose solve)',
}


def ncu_traffic(kernel):
    """DRAM bytes per launch from the committed `ncu --set full` capture of this kernel, with its provenance."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None, None
    ent = json.load(open(path)).get(kernel)
    if not ent:
        return None, None
    return ent.get('dram_bytes_per_launch'), '%s@%s' % (ent.get('source'), ent.get('commit'))


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
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


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
    sample = '%d frames per step at batch size 1 (the reference evaluates one frame at a time, utils/evaluation.py:69), ' \
             '%d hypotheses' % (per_step, args.hyps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * per_step / value, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_of(args, int(os.environ.get('WORLD_SIZE', str(args.gpus)))),
        'impl_detail': {'reference_path': 'stock torch ops on the host cores + C/OpenMP restatement of dsacstar_rgb_forward',
                        'frames_per_step': per_step},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample, **detail},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------- extras of the native arm
def stock_gpu_baseline(dev, hyps, frames=12):
    """SURVEY 8d baseline B, the reference's real deployment: stock torch / cuDNN forward at batch size 1 on this GPU,
    `.cpu()`, then the CPU solver (C/OpenMP restatement on all host cores); TF32 as torch ships it (convolutions: on)
    and switched off.  Also the stock forward alone at the bench batch size."""
    from oracle import dsac_oracle_c as tier2
    net = build_network(dev)
    images, offsets, _, focal = synthetic_batch(7000, 4)
    out = {'frames_timed': frames, 'batch': 1, 'solver': 'oracle/dsac_oracle.c, OpenMP, %d threads' % len(os.sched_getaffinity(0))}
    saved = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            t_net = t_solve = 0.0
            for i in range(frames + 3):
                j = i % 4
                t0 = time.perf_counter()
                with torch.no_grad():
                    pred = net.forward_reference(images[j:j + 1].to(dev))      # network(image.cuda())
                    coords = (pred[:, :3] + offsets[j:j + 1].to(dev)).cpu()    # .cpu(), utils/evaluation.py:161
                t1 = time.perf_counter()
                tier2.forward_rgb(np.ascontiguousarray(coords[0].numpy()), hyps, 10.0, float(focal[j]), WIDTH / 2, HEIGHT / 2,
                                  100.0, 100.0, 8, seed=1305, image=7000 + j)
                t2 = time.perf_counter()
                if i >= 3:
                    t_net += t1 - t0
                    t_solve += t2 - t1
            key = 'tf32_on' if tf32 else 'tf32_off'
            out[key] = {'images_per_s': frames / (t_net + t_solve), 'network_ms_per_image': 1e3 * t_net / frames,
                        'solver_ms_per_image': 1e3 * t_solve / frames}
            big = torch.rand(BATCH, 3, HEIGHT, WIDTH, device=dev)
            with torch.no_grad():
                for _ in range(2):
                    net.forward_reference(big)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    net.forward_reference(big)
                e1.record()
                torch.cuda.synchronize()
            out[key]['forward_ms_per_%d_frames' % BATCH] = e0.elapsed_time(e1) / 3
            del big
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def parity_block(net, dev, images_d, offsets_d, focal_h, poses_native, image_base, hyps, gt_poses):
    """Parity guard reported with the throughput: the regressed map vs fp32 torch on the timed weights, and
    the poses of the timed frames vs the CPU oracle run on the very same maps with the same (seed, image index)."""
    from crossloc_b200 import synth
    from oracle import dsac_oracle_c as tier2
    saved = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            native = net(images_d[:4])
            ref = net.forward_reference(images_d[:4])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    rel = float((native[:, :3].double() - ref[:, :3].double()).norm() / ref[:, :3].double().norm())
    mx = float((native[:, :3] - ref[:, :3]).abs().max() / ref[:, :3].abs().max())
    solver_in = (native[:, :3] + offsets_d[:4]).cpu().numpy()
    dpose, same = [], 0
    for b in range(4):
        o = tier2.forward_rgb(np.ascontiguousarray(solver_in[b]), hyps, 10.0, float(focal_h[b]), WIDTH / 2, HEIGHT / 2, 100.0,
                              100.0, 8, seed=1305, image=image_base + b)
        d = float(np.abs(o['pose'] - poses_native[b]).max() / max(1.0, np.abs(o['pose']).max()))
        dpose.append(d)
        same += int(d < 1e-3)
    errs = np.array([synth.pose_errors(gt_poses[b], poses_native[b]) for b in range(poses_native.shape[0])])
    return {'coord_rel_l2_vs_fp32_torch': rel, 'coord_max_abs_rel': mx, 'coord_tolerance': 1e-3, 'frames_checked': 4,
            'oracle_pose_max_rel_diff': max(dpose), 'oracle_poses_matching_1e-3': same,
            'median_t_err_m': float(np.median(errs[:, 0])), 'median_r_err_deg': float(np.median(errs[:, 1])),
            'note': 'oracle = oracle/dsac_oracle.c on the GPU-regressed map of the same frame, same seed and image index'}


def latency_batch1(net, dev, hyps, frames=200):
    """The reference's call pattern per frame (test_single_task.py:347-356, utils/evaluation.py:156-172): a CPU image,
    `network(image.cuda())`, torch.split, `.cpu()`, dsacstar.forward_rgb on CPU tensors with a CPU [4, 4] output."""
    import dsacstar
    images, offsets, _, focal = synthetic_batch(9000, 4)
    offs_d = offsets.to(dev)
    lat, t_net, t_solve = [], [], []
    dsacstar.set_seed(1305, 9000)
    for i in range(frames + 10):
        j = i % 4
        image = images[j:j + 1]
        t0 = time.perf_counter()
        with torch.no_grad():
            predictions = net(image.cuda())
            predictions, _unc = torch.split(predictions, [net.num_task_channel, net.num_pos_channel], dim=1)
            predictions = predictions + offs_d[j:j + 1]       # consistent synthetic scene
            out_pose = torch.zeros((4, 4))
            scene_coords = predictions.cpu()
        t1 = time.perf_counter()
        dsacstar.forward_rgb(scene_coords, out_pose, hyps, 10, float(focal[j]), float(WIDTH / 2), float(HEIGHT / 2), 100, 100,
                             net.OUTPUT_SUBSAMPLE)
        t2 = time.perf_counter()
        if i >= 10:
            lat.append(1e3 * (t2 - t0))
            t_net.append(1e3 * (t1 - t0))
            t_solve.append(1e3 * (t2 - t1))
    lat.sort()
    return {'p50_ms': lat[len(lat) // 2], 'p90_ms': lat[int(len(lat) * 0.9)], 'mean_ms': sum(lat) / len(lat),
            'network_incl_h2d_d2h_p50_ms': sorted(t_net)[len(t_net) // 2], 'solver_p50_ms': sorted(t_solve)[len(t_solve) // 2],
            'frames': frames, 'hypotheses': hyps,
            'pattern': 'network(image.cuda()) -> split -> .cpu() -> dsacstar.forward_rgb(cpu map, cpu [4,4]); pageable host image'}


def train_workload(dev, steps=5, warmup=2):
    """BASELINE config 4: train_single_task.py-shaped step (coord MLE loss, forward + backward + Adam), batch 12 at
    480x720: the native fused plan vs stock autograd (cuDNN, TF32 as torch ships it) on the same GPU."""
    import networks.networks as nets
    from crossloc_b200 import synth, train_plan
    from loss.coord import scene_coords_regression_loss
    batch = TRAIN_BATCH
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.tensor(synth.NATURESCAPE_MEAN, dtype=torch.float32), False, False, 2, 2, 3, 1).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    _, gt, poses, _ = synth.make_batch(0, batch)
    images = torch.rand(batch, 3, HEIGHT, WIDTH, device=dev)
    gt = torch.from_numpy(gt).to(dev)
    poses = torch.from_numpy(poses).float().to(dev)
    cam = torch.eye(3, device=dev)
    cam[0, 0] = cam[1, 1] = 480.0
    cam[0, 2], cam[1, 2] = WIDTH / 2, HEIGHT / 2
    xs = torch.arange(0, 135 * 8, 8, dtype=torch.float32) + 4        # utils/learning.py:20-35 pixel grid
    grid = torch.stack([xs[None, :].expand(135, 135), xs[:, None].expand(135, 135)]).to(dev)
    out = {'batch': batch, 'steps': steps}
    variants = (('native_fused', lambda t: train_plan.forward_train(net, t)),
                ('native_fused_tf32_grade', lambda t: train_plan.forward_train(net, t, backward='fp16x1')),
                ('stock_autograd_cudnn_tf32', net.forward_reference))
    for name, fwd in variants:
        ms = []
        for i in range(steps + warmup):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt.zero_grad()
            pred = fwd(images)
            c, u = torch.split(pred, [3, 1], dim=1)
            loss, _rate = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
            loss.backward()
            opt.step()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ms.append(e0.elapsed_time(e1))
        out[name] = {'ms_per_step': sum(ms) / len(ms), 'images_per_s': batch * 1e3 * len(ms) / sum(ms), 'loss': float(loss)}
    out['speedup_vs_stock'] = out['stock_autograd_cudnn_tf32']['ms_per_step'] / out['native_fused']['ms_per_step']
    out['speedup_vs_stock_tf32_grade'] = out['stock_autograd_cudnn_tf32']['ms_per_step'] / out['native_fused_tf32_grade']['ms_per_step']
    out['arithmetic'] = {
        'native_fused': 'forward %s, data gradients %s (fp16 + block-scaled e2m1 correction products on the 256/512-channel '
                        'layers, fp16x3 elsewhere), weight gradients %s (one fp16 pass); all gradients within 5.3e-5 relative L2 '
                        'of the all-fp16x3 backward on the same forward (worst convolution weight 1.4e-4; tools/dbg_wgrad_precision.py) '
                        'and within 3.7e-4 of fp32 autograd, where stock TF32 autograd is at 1.2e-3 (profiles/r2s4_grad_vs_autograd.json)'
                        % (train_plan.FORWARD, train_plan.BACKWARD, train_plan.WGRAD),
        'native_fused_tf32_grade': 'both gradient GEMMs in one fp16 pass (10-bit mantissa operands like the TF32 kernels stock '
                                   'PyTorch trains with; gradients within 3.0e-4 of the three-term result)',
        'stock_autograd_cudnn_tf32': 'torch defaults: cuDNN convolutions with TF32 allowed'}
    out['peak_mem_gb'] = torch.cuda.max_memory_allocated() / 1e9
    del net, opt
    torch.cuda.empty_cache()
    return out


def eval5_workload(net, dev, hyps, rank, world, per_gpu=EVAL5_IMAGES_PER_GPU, batch=BATCH):
    """BASELINE config 5: a naturescape-shaped synthetic set (pre-generated, 256 frames per GPU: 2048 on 8) sharded over
    the ranks, 256 hypotheses, one pose gather, the reference's accuracy summary on rank 0.  Wall-clock images/s over
    the localization of the pre-generated frames (uint8 pinned host frames -> poses on the host)."""
    import torch.distributed as dist
    from crossloc_b200 import parallel, synth
    from crossloc_b200.pipeline import Localizer
    n_total = per_gpu * world
    mine = parallel.shard_indices(n_total, rank, world)
    loc = Localizer(net, hyps=hyps, device=dev)
    batches = []
    for s in range(0, len(mine), batch):   # pre-generation: scenes + frames, outside the timed region
        idx = mine[s:s + batch]
        scenes = [synth.make_scene(i) for i in idx]
        frames = torch.randint(0, 256, (len(idx), HEIGHT, WIDTH, 3), dtype=torch.uint8,
                               generator=torch.Generator().manual_seed(idx[0])).pin_memory()
        offs = torch.stack([torch.from_numpy(sc['coords']) for sc in scenes]).to(dev)
        focal = torch.tensor([sc['focal'] for sc in scenes], dtype=torch.float32, device=dev)
        batches.append((idx, frames, offs, focal, [sc['pose'] for sc in scenes]))
    loc.localize(batches[0][1], batches[0][3], batches[0][2], image_base=batches[0][0][0])   # warm-up: plan, graph
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    poses = []
    loc.submit(batches[0][1], batches[0][3], batches[0][2], image_base=batches[0][0][0])
    for k in range(1, len(batches)):
        loc.submit(batches[k][1], batches[k][3], batches[k][2], image_base=batches[k][0][0])
        poses.append(loc.result().clone())
    poses.append(loc.result().clone())
    torch.cuda.synchronize()
    rows = torch.empty(len(mine), 18, dtype=torch.float32)
    r = 0
    for (idx, _, _, _, gts), p in zip(batches, poses):
        for j in range(len(idx)):
            t, rot = synth.pose_errors(gts[j], p[j].numpy())
            rows[r, :16] = p[j].reshape(16)
            rows[r, 16], rows[r, 17] = t, rot
            r += 1
    allrows = parallel.gather_rows(mine, rows.to(dev), n_total)   # the one collective of the path: [n, 18] fp32
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    cpu = allrows.cpu().numpy()
    summary = parallel.summarize(cpu[:, 16], cpu[:, 17])
    summary.update({'images': n_total, 'world': world, 'hypotheses': hyps, 'wall_s': float(dt.item()),
                    'images_per_s': n_total / float(dt.item()),
                    'timed': 'uint8 pinned frames -> H2D -> CNN + solve -> poses on the host -> pose errors -> gather',
                    'note': 'median translation error is the constant offset of the random-init network output'})
    return summary


# ---------------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch.distributed as dist
    from crossloc_b200 import _lib
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
    lib = _lib.load()

    if args.workload == 'train':
        res = train_workload(dev, steps=args.steps, warmup=max(2, args.warmup))
        if rank == 0:
            emit({'metric': 'images/sec trained (480x720, coord MLE loss, forward + backward + Adam)', 'unit': UNIT,
                  'value': res['native_fused']['images_per_s'], 'n_gpus': 1, 'steps': args.steps, 'warmup': max(2, args.warmup),
                  'ms_per_step': res['native_fused']['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                  'vs_baseline': None, 'dtype': 'f16 split products -> f32', 'data': 'synthetic',
                  'config': {'workload': 'train_batch%d_480x720_coord_mle' % TRAIN_BATCH}, 'detail': res})
        return

    net = build_network(dev)
    if args.workload == 'eval5':
        res = eval5_workload(net, dev, args.hyps, rank, world)
        if rank == 0:
            emit({'metric': METRIC, 'unit': UNIT, 'value': res['images_per_s'], 'n_gpus': world, 'steps': 1, 'warmup': 1,
                  'ms_per_step': 1e3 * res['wall_s'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                  'dtype': DTYPES.get(net._runtime.precision if getattr(net, '_runtime', None) else '', 'f16 split products'),
                  'data': 'synthetic', 'config': {'workload': 'eval5_%d_frames_sharded' % res['images']}, 'detail': res})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    loc = Localizer(net, hyps=args.hyps, device=dev)
    # distinct synthetic frames per rank: rank r localizes images r*B .. r*B+B-1 of every step (weak scaling)
    images_h, offsets_h, gt_poses, focal_h = synthetic_batch(rank * B, B)
    images_h = images_h.pin_memory()
    images_d = images_h.to(dev)
    offsets_d = offsets_h.to(dev)
    focal_d = focal_h.to(dev)
    gathered = torch.empty(world * B, 16, dtype=torch.float32, device=dev) if world > 1 else None
    if world > 1:   # trivial pose gather: 2 KB per rank, on the solver stream behind every solve
        loc.after_solve = lambda pose: dist.all_gather_into_tensor(gathered, pose.reshape(B, 16))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(image_base):
        # the solve of this batch is deferred behind the fork point of the NEXT batch's CNN (pipeline.Localizer)
        return loc.localize_device(images_d, focal_d, offsets_d, image_base=image_base, overlap=True)

    # ---- device-resident throughput ("value")
    for w in range(args.warmup):
        step_device(w * world * B + rank * B)
    loc.flush()
    barrier()
    rt = net._runtime
    rt.set_profiling(True)                       # eager launches with an event between ops: every kernel is timed
    step_device(args.warmup * world * B + rank * B)   # one untimed step in that mode
    loc.flush()
    barrier()
    rt.read_profile(B, HEIGHT, WIDTH)            # discard
    timing = (ctypes.c_float * 3)()
    nsolves = ctypes.c_int()
    _lib.check(lib.cl_dsac_timing(1, loc.solver_stream.cuda_stream, None, None))
    launches0, solver0 = rt.launches, loc.solver_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for k in range(args.steps):
        pose = step_device((args.warmup + 1 + k) * world * B + rank * B)
    last_base = (args.warmup + args.steps) * world * B + rank * B
    torch.cuda.current_stream().wait_event(loc.flush())   # the last solve is inside the timed region
    stop.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    _lib.check(lib.cl_dsac_timing(0, loc.solver_stream.cuda_stream, timing, ctypes.byref(nsolves)))
    prof = rt.read_profile(B, HEIGHT, WIDTH, keep_enabled=False)
    launches_per_step = (rt.launches - launches0 + loc.solver_launches - solver0) / args.steps
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)
    pose_np = pose.cpu().numpy()

    # ---- per-kernel table and the two rooflines
    (peak_tf, peak_tf_src), (peak_bw, peak_bw_src) = peaks()
    by_shape = {}
    for kind, label, flops, ms, forwards in prof:
        if kind in ('memset', 'fork'):
            continue
        key = {'conv': '/'.join(str(x) for x in label), 'conv_fused': '/'.join(str(x) for x in label) + '+gn',
               'gn_apply': 'gn_apply/%d/%d/%d' % label[:3], 'stem': 'stem/pass%d' % label[0]}.get(kind, kind)
        ent = by_shape.setdefault(key, [0, 0.0, flops, max(forwards, 1)])
        ent[0] += 1
        ent[1] += ms
    kernel_ms = {k: {'launches_per_step': v[0], 'ms_per_step': v[1] / v[3], 'avg_ms': v[1] / v[3] / v[0],
                     'tflops_useful': (v[2] / (v[1] / v[3] / v[0] * 1e-3) / 1e12) if v[2] and v[1] else None}
                 for k, v in by_shape.items()}
    nsol = max(1, nsolves.value)
    for i, name in enumerate(('dsac_sample', 'dsac_score', 'dsac_refine')):
        kernel_ms[name] = {'launches_per_step': 1, 'ms_per_step': timing[i] / nsol, 'avg_ms': timing[i] / nsol,
                           'tflops_useful': None}
    conv_total_ms = sum(v['ms_per_step'] for k, v in kernel_ms.items() if k.count('/') == 3 and not k.startswith('gn'))
    roofline = None
    fused = '512/512/3/1+gn' in kernel_ms
    dom = kernel_ms.get('512/512/3/1+gn') or kernel_ms.get('512/512/3/1')
    if dom:
        achieved = dom['tflops_useful']
        nterms = rt.nterms
        kname = 'conv_igemm_pair_fp4_kernel' if rt.precision == 'fp16+fp4' else 'conv_igemm_pair_kernel<64%s>' % (
            ', fused GroupNorm epilogue' if fused else '')
        traffic, traffic_src = ncu_traffic('%s 3x3 512->512' % kname)
        roofline = {'bound': 'tensor', 'kernel': '%s 3x3 512->512 @60x90 x%d images' % (kname, B),
                    'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                    'traffic': traffic, 'traffic_source': traffic_src,
                    'traffic_unit': 'bytes per launch (dram read + write, ncu --set full)',
                    'peak_source': peak_tf_src, 'avg_launch_ms': dom['avg_ms'],
                    'launches_timed': dom['launches_per_step'] * args.steps,
                    'issued_tflops': achieved * nterms, 'issued_frac': achieved * nterms / peak_tf,
                    'note': 'achieved counts algorithmic FLOPs (2*pixels*Cout*Cin*9); %s issues %.1f fp16-MMA equivalents per '
                            'product (one fp16 MMA + the two 2^-11 correction products as block-scaled e2m1 MMAs at four times '
                            'the fp16 rate [fp16+fp4] or as e4m3 MMAs at twice the rate [fp16+fp8])' % (rt.precision, nterms),
                    'all_conv_ms_per_step': conv_total_ms}
    score_ms = kernel_ms['dsac_score']['avg_ms']
    cells = 60 * 90
    streamed = B * args.hyps * cells * 12 + B * args.hyps * 8           # every hypothesis streams the planar map once
    traffic, traffic_src = ncu_traffic('dsac_score_kernel')
    roofline_score = {'bound': 'hbm', 'kernel': 'dsac_score_kernel (%d x %d hypotheses x %d cells)' % (B, args.hyps, cells),
                      'achieved': streamed / (score_ms * 1e-3) / 1e9 if score_ms else None, 'peak': peak_bw, 'unit': 'GB/s',
                      'frac': (streamed / (score_ms * 1e-3) / 1e9 / peak_bw) if score_ms else None,
                      'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_bw_src, 'avg_launch_ms': score_ms,
                      'cell_evaluations_per_s': B * args.hyps * cells / (score_ms * 1e-3) if score_ms else None,
                      'note': 'streamed bytes = B*hyps*cells*12 (SURVEY 8d); the 64.8 KB maps are L2 resident, the pass is bound by '
                              'the fp64 projection the reference prescribes (dsacstar_util.h:395-443), not by HBM'}

    # ---- end to end through the public API with host buffers ("e2e"): graph replay, deferred solves
    barrier()
    for w in range(3):
        loc.submit(images_h, focal_d, offsets_d, image_base=w * B)
    for w in range(3):
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

    # ---- the same device-resident loop as CUDA-graph replays (the production mode), for comparison with `value`
    barrier()
    for w in range(2):
        step_device(w * B)
    loc.flush()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for k in range(args.steps):
        step_device(k * world * B + rank * B)
    torch.cuda.current_stream().wait_event(loc.flush())
    g1.record()
    barrier()
    graph_ms = g0.elapsed_time(g1) / args.steps

    extras = {}
    if not args.no_extras:
        extras['config5_eval'] = eval5_workload(net, dev, args.hyps, rank, world)
    line = None
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline and not args.no_extras:
            threads = len(os.sched_getaffinity(0))
            os.environ['OMP_NUM_THREADS'] = str(threads)
            sample = 6
            rate, detail = CpuPath(args.hyps, threads, sample).rate(sample)
            cpu_baseline = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                            'sample': '%d frames, batch size 1, %d hypotheses: stock torch network on the host cores + '
                                      'oracle/dsac_oracle.c (OpenMP)' % (sample, args.hyps), **detail}
        if world == 1 and not args.no_extras:
            extras['parity'] = parity_block(net, dev, images_d, offsets_d, focal_h, pose_np, last_base - 0, args.hyps, gt_poses)
            extras['latency_batch1'] = latency_batch1(net, dev, args.hyps)
            extras['latency_batch1']['batch%d_ms_per_frame' % B] = elapsed_ms / args.steps / B
            extras['gpu_stock_baseline'] = stock_gpu_baseline(dev, args.hyps)
            del loc
            torch.cuda.empty_cache()
            extras['config4_train'] = train_workload(dev)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': DTYPES.get(rt.precision, rt.precision), 'data': 'synthetic',
            'config': config_of(args, world),
            'impl_detail': {'conv_precision': rt.precision, 'timed_mode': 'cl_net_forward, per-op event mode (eager launches)',
                            'graph_mode_ms_per_step': graph_ms,
                            'solver': 'deferred: batch k solves while batch k+1 runs its residual blocks'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(images_h.numel() * 4),
                    'd2h_bytes_per_step': int(B * 16 * 4)},
            'gpu_launches': int(round(launches_per_step * args.steps)),
            'gpu_launches_per_step': launches_per_step,
            'roofline': roofline,
            'roofline_score': roofline_score,
            'cpu_baseline': cpu_baseline,
            'conv_tflops_useful': B * CONV_GFLOP_PER_IMAGE / conv_total_ms if conv_total_ms else None,
            'pose_error_median': {'t_m': extras.get('parity', {}).get('median_t_err_m'),
                                  'r_deg': extras.get('parity', {}).get('median_r_err_deg')},
            'kernel_ms': kernel_ms,
        }
        line.update(extras)
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
