"""CPU suite: pins the oracles against cv2 and the committed golden vectors (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import cv2
import numpy as np
import pytest

from crossloc_b200 import rng, synth
from oracle import dsac_oracle_c as tier2
from oracle import dsac_oracle_py as tier1
from tests.util import GOLDEN_CASES, PARAMS, ROOT, golden_case, score_mismatch


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    assert rng.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert rng.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert rng.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_sample_cells_in_range_and_with_replacement():
    seen_dup = False
    for h in range(200):
        cells = rng.sample_cells(1305, 0, h, 0, 7, 5)
        assert all(0 <= x < 7 and 0 <= y < 5 for x, y in cells)
        seen_dup |= len(set(cells)) < 4
    assert seen_dup   # drawn with replacement (dsacstar_util.h:168-173)


def _random_minimal_sets(count, seed):
    g = np.random.default_rng(seed)
    for it in range(count):
        co = synth.make_scene(it % 5)['coords']
        xs, ys = g.integers(0, 90, 4), g.integers(0, 60, 4)
        img = np.array([[x * 8 + 4, y * 8 + 4] for x, y in zip(xs, ys)], dtype=np.float32)
        obj = np.array([co[:, y, x] for x, y in zip(xs, ys)], dtype=np.float32)
        yield obj, img, len({(x, y) for x, y in zip(xs, ys)}) < 4


def test_p3p_matches_cv2():
    """Tier-2 P3P picks the same root as cv2's SOLVEPNP_P3P and reprojects the minimal set at least as well."""
    k = tier1.cam_mat(480, 360, 240)
    both = 0
    for obj, img, dup in _random_minimal_sets(600, 0):
        ok1, r1, t1 = tier1._safe_solve_pnp(obj, img, k, None, None, False, cv2.SOLVEPNP_P3P)
        ok1 = ok1 and np.isfinite(r1).all() and np.isfinite(t1).all()   # cv2 can return True with a NaN pose
        ok2, r2, t2 = tier2.p3p(obj, img, 480, 360, 240)
        assert ok1 == ok2
        if not ok1 or dup:
            continue
        both += 1
        p1 = cv2.projectPoints(obj.astype(np.float64), r1, t1, k.astype(np.float64), None)[0].reshape(-1, 2)
        p2 = cv2.projectPoints(obj.astype(np.float64), r2, t2, k.astype(np.float64), None)[0].reshape(-1, 2)
        # both fit the three defining points; cv2's own residual is ~1e-5 px
        assert np.abs(p1[:3] - img[:3]).max() < 1e-3
        assert np.abs(p2[:3] - img[:3]).max() < 1e-6
        # same root: the 4th point lands in the same place
        assert np.abs(p1[3] - p2[3]).max() < 1e-2 * max(1.0, np.abs(p1[3] - img[3]).max())
    assert both > 400


def test_lm_matches_cv2_iterative():
    """CvLevMarq restatement == cv2.solvePnP(ITERATIVE, useExtrinsicGuess=True) to round-off."""
    k = tier1.cam_mat(480, 360, 240)
    samp = tier1.create_sampling(90, 60, 8)
    for i in range(3):
        s = synth.make_scene(i)
        r = tier1.forward_rgb(s['coords'], 8, 10., 480., 360., 240., 100., 100., 8, image=i, refine=False)
        rv, tv = r['rvec'].reshape(3, 1), r['tvec'].reshape(3, 1)
        e = tier1.repro_errs(s['coords'], rv, tv, samp, k, 100.)
        ys, xs = np.nonzero((e < 10).T)[1], np.nonzero((e < 10).T)[0]
        img = samp[ys, xs].astype(np.float32)
        obj = np.ascontiguousarray(s['coords'][:, ys, xs].T)
        ok, r1, t1 = cv2.solvePnP(obj, img, k, None, rv.copy(), tv.copy(), True, cv2.SOLVEPNP_ITERATIVE)
        ok2, r2, t2 = tier2.lm(obj, img, 480., 360., 240., rv, tv)
        assert ok and ok2
        assert np.abs(r1.ravel() - r2).max() < 1e-10
        assert np.abs(t1.ravel() - t2).max() < 1e-8


def _check_against(ref, out, cells=None):
    assert int(ref['best']) == out['best']
    assert (ref['tries'] == out['tries']).all()
    bad = score_mismatch(ref['scores'], out['scores'])
    # a minimal set that holds a cell twice makes P3P's disambiguation a tie
    assert bad.sum() <= 2 and not bad[int(ref['best'])]
    assert list(ref['counts']) == list(out['refine_counts'])
    assert np.abs(ref['pose'] - out['pose']).max() < 1e-4 * max(1.0, np.abs(ref['pose']).max())


@pytest.mark.parametrize('ci', range(len(GOLDEN_CASES)))
def test_tier2_matches_golden(ci):
    idx, hyps, scene, (cx, cy), ref = golden_case(ci)
    out = tier2.forward_rgb(scene['coords'], hyps, PARAMS['thr'], scene['focal'], cx, cy, PARAMS['alpha'],
                            PARAMS['max_reproj'], PARAMS['sub_sampling'], seed=PARAMS['seed'], image=idx)
    _check_against(ref, out)


def test_tier1_reproduces_golden():
    """The committed fixture is what the cv2 oracle produces today (guards against a silent cv2 change)."""
    idx, hyps, scene, (cx, cy), ref = golden_case(3)
    out = tier1.forward_rgb(scene['coords'], hyps, PARAMS['thr'], scene['focal'], cx, cy, PARAMS['alpha'],
                            PARAMS['max_reproj'], PARAMS['sub_sampling'], seed=PARAMS['seed'], image=idx)
    _check_against(ref, out)


@pytest.mark.parametrize('idx,kw', [
    (100, {}), (101, {'outlier_ratio': 0.5}), (102, {'noise_sigma': 2.0}), (103, {'height': 200, 'width': 312}),
    (104, {'outlier_ratio': 0.1, 'noise_sigma': 0.1}), (105, {'subsample': 16}), (106, {}), (107, {'outlier_ratio': 0.7}),
])
def test_tier2_matches_tier1_on_scenes_outside_the_fixture(idx, kw):
    """The C restatement against the cv2-based one, live, on scenes the committed fixture does not contain: same winner,
    same number of tries per hypothesis, same refinement inlier counts, same pose."""
    s = synth.make_scene(idx, **kw)
    h, w, sub = kw.get('height', 480), kw.get('width', 720), kw.get('subsample', 8)
    args = (s['coords'], 24, 10., s['focal'], w / 2, h / 2, 100., 100., sub)
    ref = tier1.forward_rgb(*args, seed=1305, image=idx)
    out = tier2.forward_rgb(*args, seed=1305, image=idx)
    assert int(ref['best']) == out['best']
    assert (np.asarray(ref['tries']) == np.asarray(out['tries'])).all()
    bad = score_mismatch(ref['scores'], out['scores'])
    assert bad.sum() <= 2 and not bad[int(ref['best'])]
    assert list(ref['refine_counts']) == list(out['refine_counts'])
    assert np.abs(ref['pose'] - out['pose']).max() < 1e-4 * max(1.0, np.abs(ref['pose']).max())


def test_ground_truth_map_gives_zero_error():
    """The authors' debug hook (test_single_task.py:361): GT coordinates in, ~0 pose error out."""
    s = synth.make_scene(7, noise_sigma=0.0, outlier_ratio=0.0)
    out = tier2.forward_rgb(s['coords'], 64, 10., s['focal'], 360., 240., 100., 100., 8, image=7)
    t_err, r_err = synth.pose_errors(s['pose'], out['pose'])
    assert t_err < 1e-3 and r_err < 1e-3
    assert out['refine_counts'][0] == 5400


def test_forced_samples_replay():
    s = synth.make_scene(1)
    free = tier2.forward_rgb(s['coords'], 16, 10., 480., 360., 240., 100., 100., 8, image=1)
    cells = np.stack([rng.sample_cells_array(1305, 1, 16, int(t) - 1, 90, 60)[h] for h, t in enumerate(free['tries'])])
    replay = tier2.forward_rgb(s['coords'], 16, 10., 480., 360., 240., 100., 100., 8, image=1, forced_samples=cells)
    assert np.allclose(free['hyps_rt'], replay['hyps_rt'], rtol=0, atol=0)
    assert free['best'] == replay['best']


def test_degenerate_constant_map_terminates():
    co = np.full((3, 60, 90), 2.0, dtype=np.float32)
    out = tier2.forward_rgb(co, 8, 10., 480., 360., 240., 100., 100., 8, max_tries=50)
    assert (out['tries'] == 50).all()
    assert np.isfinite(out['pose']).all()


def test_host_math_of_cuda_solver_matches_tier2(tmp_path):
    """The CUDA solver's geometry header, compiled for the host, agrees with the tier-2 oracle."""
    so = str(tmp_path / 'libhostmath.so')
    subprocess.check_call(['/usr/bin/g++', '-O2', '-shared', '-fPIC', '-o', so,
                           os.path.join(ROOT, 'tests', 'host_math_harness.cpp')])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    cells = (ctypes.c_int * 8)()
    for (i, h, t) in [(0, 0, 0), (3, 17, 5), (100, 255, 999)]:
        lib.hm_sample_cells(ctypes.c_uint64(1305), i, h, t, 90, 60, cells)
        assert [tuple(cells[2 * j:2 * j + 2]) for j in range(4)] == rng.sample_cells(1305, i, h, t, 90, 60)
    mism = 0
    for obj, img, dup in _random_minimal_sets(400, 1):
        obj, img = obj.astype(np.float64), img.astype(np.float64)
        ok2, r2, t2 = tier2.p3p(obj, img, 480, 360, 240)
        r, t = np.zeros(3), np.zeros(3)
        ok = lib.hm_p3p(obj.ctypes.data_as(dp), img.ctypes.data_as(dp), ctypes.c_double(480), ctypes.c_double(360),
                        ctypes.c_double(240), r.ctypes.data_as(dp), t.ctypes.data_as(dp))
        assert bool(ok) == ok2
        if ok2 and not dup and (np.abs(r - r2).max() > 1e-9 or np.abs(t - t2).max() > 1e-7):
            mism += 1
    assert mism <= 2


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/crossloc_b200.h declares."""
    from crossloc_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'crossloc_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(cl_\w+)\s*\(', header))
    assert declared and declared == set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert b'sm_100a' in lib.cl_version()


def test_rigid_motion_equivariance():
    """Size-independent property of the path: moving the whole scene by a rigid motion G moves the estimated
    camera-to-world pose by G (same samples, same winner; the solver only sees relative geometry)."""
    s = synth.make_scene(11, outlier_ratio=0.2, noise_sigma=0.05)
    rng_np = np.random.default_rng(5)
    axis = rng_np.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = 0.7
    kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    rot = np.eye(3) + np.sin(ang) * kx + (1 - np.cos(ang)) * kx @ kx
    shift = np.array([120.0, -45.0, 300.0])
    moved = (rot @ s['coords'].reshape(3, -1).astype(np.float64) + shift[:, None]).reshape(s['coords'].shape).astype(np.float32)
    a = tier2.forward_rgb(s['coords'], 32, 10., s['focal'], 360., 240., 100., 100., 8, image=11)
    b = tier2.forward_rgb(moved, 32, 10., s['focal'], 360., 240., 100., 100., 8, image=11)
    assert a['best'] == b['best'] and (np.asarray(a['tries']) == np.asarray(b['tries'])).all()
    g = np.eye(4)
    g[:3, :3], g[:3, 3] = rot, shift
    expect = g @ a['pose'].astype(np.float64)
    assert np.abs(expect[:3, :3] - b['pose'][:3, :3]).max() < 1e-4
    assert np.abs(expect[:3, 3] - b['pose'][:3, 3]).max() < 2e-2     # fp32 scene coordinates ~1e3 m: 1e-4 m resolution
