"""GPU parity of the DSAC* solver: CUDA path (through the C ABI) vs the oracles on the same seeded inputs."""
import numpy as np
import pytest
import torch

from crossloc_b200 import dsac, rng, synth
from oracle import dsac_oracle_c as tier2
from tests.util import GOLDEN_CASES, PARAMS, golden_case, score_mismatch

pytestmark = pytest.mark.gpu


def _run(coords, hyps, focal, cx, cy, image_base, device='cuda', **kw):
    c = torch.from_numpy(np.ascontiguousarray(coords))
    if c.dim() == 3:
        c = c.unsqueeze(0)
    c = c.to(device)
    pose = torch.zeros(c.size(0), 4, 4, dtype=torch.float32, device=device)
    dbg = dsac.forward_rgb_batch(c, pose, hyps, PARAMS['thr'], focal, cx, cy, PARAMS['alpha'], PARAMS['max_reproj'],
                                 PARAMS['sub_sampling'], seed=PARAMS['seed'], image_base=image_base, debug=True, **kw)
    if pose.is_cuda:
        torch.cuda.synchronize()
    return pose.cpu().numpy(), {k: v.numpy() for k, v in dbg.items()}


def _counts(row):
    return [int(c) for c in row if c >= 0]


def _compare(ref, pose, dbg, b=0):
    assert int(ref['best']) == int(dbg['best'][b])
    assert (np.asarray(ref['tries']) == dbg['tries'][b]).all()
    bad = score_mismatch(ref['scores'], dbg['scores'][b])
    assert bad.sum() <= 2 and not bad[int(ref['best'])]   # duplicate-cell minimal sets tie in P3P
    ref_counts = list(ref['counts']) if 'counts' in ref else list(ref['refine_counts'])
    assert ref_counts == _counts(dbg['refine_counts'][b])
    assert np.abs(ref['pose'] - pose[b]).max() < 1e-4 * max(1.0, np.abs(ref['pose']).max())


@pytest.mark.parametrize('ci', range(len(GOLDEN_CASES)))
def test_matches_golden_vectors(ci):
    """Bit-for-bit decisions (tries, winner, refinement inlier counts) and pose vs the cv2-based fixtures."""
    idx, hyps, scene, (cx, cy), ref = golden_case(ci)
    pose, dbg = _run(scene['coords'], hyps, scene['focal'], cx, cy, idx)
    _compare(ref, pose, dbg)


@pytest.mark.parametrize('hyps', [64, 256])
def test_matches_tier2_oracle(hyps):
    for idx in (10, 11):
        s = synth.make_scene(idx)
        ref = tier2.forward_rgb(s['coords'], hyps, PARAMS['thr'], s['focal'], 360., 240., PARAMS['alpha'],
                                PARAMS['max_reproj'], 8, seed=PARAMS['seed'], image=idx)
        pose, dbg = _run(s['coords'], hyps, s['focal'], 360., 240., idx)
        _compare(ref, pose, dbg)
        # scores agree to round-off where the minimal sets are not degenerate
        rel = np.abs(ref['scores'] - dbg['scores'][0]) / np.max(ref['scores'])
        assert np.median(rel) < 1e-9


def test_batch_equals_single_calls_and_host_equals_device():
    coords, _, poses, focal = synth.make_batch(20, 6)
    pose_b, dbg_b = _run(coords, 64, torch.from_numpy(focal), 360., 240., 20)
    pose_h, dbg_h = _run(coords, 64, torch.from_numpy(focal), 360., 240., 20, device='cpu')
    assert np.array_equal(pose_b, pose_h) and np.array_equal(dbg_b['scores'], dbg_h['scores'])
    for b in range(6):
        pose_1, dbg_1 = _run(coords[b], 64, float(focal[b]), 360., 240., 20 + b)
        assert np.array_equal(pose_1[0], pose_b[b])
        assert np.array_equal(dbg_1['scores'][0], dbg_b['scores'][b])


def test_forced_samples_replay():
    s = synth.make_scene(1)
    _, free = _run(s['coords'], 16, 480., 360., 240., 1)
    cells = np.stack([rng.sample_cells_array(1305, 1, 16, int(t) - 1, 90, 60)[h] for h, t in enumerate(free['tries'][0])])
    pose, replay = _run(s['coords'], 16, 480., 360., 240., 1, forced_samples=cells)
    assert np.array_equal(free['hyps_rt'], replay['hyps_rt'])
    ref = tier2.forward_rgb(s['coords'], 16, 10., 480., 360., 240., 100., 100., 8, image=1, forced_samples=cells)
    assert ref['best'] == int(replay['best'][0])
    assert np.abs(ref['pose'] - pose[0]).max() < 1e-4 * np.abs(ref['pose']).max()


def test_ground_truth_map_gives_zero_error():
    s = synth.make_scene(7, noise_sigma=0.0, outlier_ratio=0.0)
    pose, dbg = _run(s['coords'], 64, s['focal'], 360., 240., 7)
    t_err, r_err = synth.pose_errors(s['pose'], pose[0])
    assert t_err < 1e-3 and r_err < 1e-3
    assert dbg['refine_counts'][0][0] == 5400


def test_degenerate_maps():
    const = np.full((3, 60, 90), 2.0, dtype=np.float32)
    pose, dbg = _run(const, 8, 480., 360., 240., 0, max_tries=50)
    assert (dbg['tries'] == 50).all() and np.isfinite(pose).all()
    ref = tier2.forward_rgb(const, 8, 10., 480., 360., 240., 100., 100., 8, image=0, max_tries=50)
    assert np.abs(ref['pose'] - pose[0]).max() < 1e-5
    # a plane through the camera centre / points behind the camera: must terminate and stay finite
    s = synth.make_scene(3)
    behind = s['coords'].copy()
    behind[2] = 2 * s['pose'][2, 3] - behind[2]
    pose, dbg = _run(behind, 16, 480., 360., 240., 3, max_tries=200)
    ref = tier2.forward_rgb(behind, 16, 10., 480., 360., 240., 100., 100., 8, image=3, max_tries=200)
    assert (np.asarray(ref['tries']) == dbg['tries'][0]).all()
    assert np.isfinite(pose).all()


def test_dsacstar_dropin_signature():
    """The reference call (utils/evaluation.py:160-172): CPU [1,3,H,W] map, CPU [4,4] pose written in place."""
    import dsacstar
    s = synth.make_scene(5)
    dsacstar.set_seed(1305, 5)
    out_pose = torch.zeros((4, 4))
    ret = dsacstar.forward_rgb(torch.from_numpy(s['coords']).unsqueeze(0), out_pose, 64, 10., 480., float(720 / 2),
                               float(480 / 2), 100., 100., 8)
    assert ret is None
    t_err, r_err = synth.pose_errors(s['pose'], out_pose.numpy())
    assert t_err < 1.0 and r_err < 1.0
    with pytest.raises(NotImplementedError):
        dsacstar.forward_rgbd()


def test_full_size_batch_accuracy():
    """BASELINE config 3 shape: 32 images x 256 hypotheses; every pose lands near the ground truth."""
    coords, _, poses, focal = synth.make_batch(100, 32)
    pose, dbg = _run(coords, 256, torch.from_numpy(focal), 360., 240., 100)
    errs = np.array([synth.pose_errors(poses[b], pose[b]) for b in range(32)])
    assert np.median(errs[:, 0]) < 0.3 and np.median(errs[:, 1]) < 0.2
    assert errs[:, 0].max() < 2.0


def test_full_size_map_sub_sampling_one():
    """SURVEY.md section 8f row 2 shape: a 120x180-cell map with sub-sampling 4 and a 480x720 one with sub-sampling 1
    (345,600 cells) go through the same kernels; decisions match the CPU oracle on the smaller one."""
    s = synth.make_scene(31, subsample=4)
    c = torch.from_numpy(s['coords']).unsqueeze(0).cuda()
    pose = torch.zeros(1, 4, 4, device='cuda')
    dbg = dsac.forward_rgb_batch(c, pose, 32, 10.0, 480.0, 360.0, 240.0, 100.0, 100.0, 4, seed=1305, image_base=31, debug=True)
    ref = tier2.forward_rgb(s['coords'], 32, 10.0, 480.0, 360.0, 240.0, 100.0, 100.0, 4, seed=1305, image=31)
    assert ref['best'] == int(dbg['best'][0]) and (np.asarray(ref['tries']) == dbg['tries'][0].numpy()).all()
    assert np.abs(ref['pose'] - pose[0].cpu().numpy()).max() < 1e-4 * np.abs(ref['pose']).max()
    big = synth.make_scene(32, subsample=1)
    c = torch.from_numpy(big['coords']).unsqueeze(0).cuda()
    dsac.forward_rgb_batch(c, pose, 64, 10.0, 480.0, 360.0, 240.0, 100.0, 100.0, 1, seed=1305, image_base=32)
    t_err, r_err = synth.pose_errors(big['pose'], pose[0].cpu().numpy())
    assert t_err < 0.05 and r_err < 0.02


def test_pose_medians_inside_the_oracle_seed_spread():
    """North-star parity criterion: the CPU oracle over 10 RNG seeds gives a [min, max] of the
    median translation / rotation error over 256 synthetic frames at 256 hypotheses; the CUDA path at the reference's
    default seed (thread_rand.h: 1305) must land inside it -- and, frame by frame, on the oracle's pose of that seed."""
    frames, hyps, first = 256, 256, 3000
    coords, _, poses, focal = synth.make_batch(first, frames)
    seeds = list(range(1300, 1310))
    med_t, med_r = {}, {}
    oracle_1305 = None
    for seed in seeds:
        errs = []
        est = []
        for b in range(frames):
            o = tier2.forward_rgb(coords[b], hyps, PARAMS['thr'], float(focal[b]), 360., 240., PARAMS['alpha'],
                                  PARAMS['max_reproj'], 8, seed=seed, image=first + b)
            errs.append(synth.pose_errors(poses[b], o['pose']))
            est.append(o['pose'])
        med_t[seed] = float(np.median([e[0] for e in errs]))
        med_r[seed] = float(np.median([e[1] for e in errs]))
        if seed == 1305:
            oracle_1305 = np.stack(est)
    got = []
    for lo in range(0, frames, 64):   # batches of 64 frames through the batched entry
        c = torch.from_numpy(coords[lo:lo + 64]).cuda()
        pose = torch.zeros(c.size(0), 4, 4, dtype=torch.float32, device='cuda')
        dsac.forward_rgb_batch(c, pose, hyps, PARAMS['thr'], torch.from_numpy(focal[lo:lo + 64]).cuda(), 360., 240.,
                               PARAMS['alpha'], PARAMS['max_reproj'], 8, seed=1305, image_base=first + lo)
        got.append(pose.cpu().numpy())
    got = np.concatenate(got)
    errs = [synth.pose_errors(poses[b], got[b]) for b in range(frames)]
    gpu_t, gpu_r = float(np.median([e[0] for e in errs])), float(np.median([e[1] for e in errs]))
    assert min(med_t.values()) <= gpu_t <= max(med_t.values()), (gpu_t, med_t)
    assert min(med_r.values()) <= gpu_r <= max(med_r.values()), (gpu_r, med_r)
    # the seed spread is a real spread (the criterion is not vacuous) and the GPU reproduces the oracle of its own seed
    assert max(med_t.values()) > min(med_t.values())
    close = np.abs(got - oracle_1305).reshape(frames, -1).max(1) < 1e-3 * np.abs(oracle_1305).reshape(frames, -1).max(1)
    assert close.mean() >= 0.99, close.mean()
