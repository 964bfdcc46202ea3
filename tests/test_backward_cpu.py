"""CPU checks of the DSAC* backward pass: the tier-1 oracle against its stored fixture and
against numerical derivatives, and the CUDA path's per-hypothesis algebra (csrc/dsac_backward_math.cuh, compiled for
the host by tests/host_math_harness.cpp) against the OpenCV calls the reference makes."""
import ctypes
import math
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import dsac_backward_py as bwd
from oracle import dsac_oracle_py as fwd
from tests.util import ROOT, backward_case, backward_module

DP = ctypes.POINTER(ctypes.c_double)
D, F = ctypes.c_double, ctypes.c_float


def _p(a):
    return a.ctypes.data_as(DP)


@pytest.fixture(scope='module')
def hostmath(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('hm') / 'libhostmath.so')
    subprocess.check_call(['/usr/bin/g++', '-O2', '-shared', '-fPIC', '-o', so, os.path.join(ROOT, 'tests', 'host_math_harness.cpp')])
    lib = ctypes.CDLL(so)
    lib.hm_pose_loss.restype = D
    return lib


def _random_pose(rs):
    r = rs.normal(size=3) * rs.uniform(0.01, 2.5)
    t = rs.normal(size=3) * 100
    return r, t


def test_oracle_reproduces_stored_fixture():
    """The committed vectors are what the oracle computes today (smallest case; the generator covers all four)."""
    idx, hyps, scene, gt, cxcy, p, ref = backward_case(3)
    r = bwd.backward_rgb(scene['coords'], gt, hyps, p['thr'], scene['focal'], cxcy[0], cxcy[1], p['w_rot'], p['w_trans'],
                         p['soft_clamp'], p['alpha'], p['max_reproj'], p['sub_sampling'], seed=p['seed'], image=idx)
    assert r['loss'] == pytest.approx(float(ref['loss']), rel=1e-12)
    assert np.abs(r['grad'] - ref['grad']).max() <= 1e-12 * np.abs(ref['grad']).max()
    assert (np.asarray(r['tries']) == ref['tries']).all()


def test_oracle_loss_jacobian_is_the_derivative_of_its_loss():
    """dLoss (dsacstar_loss.h:96-212) differentiates the loss of the INVERTED poses; check it numerically below the clamp."""
    rs = np.random.default_rng(3)
    for _ in range(20):
        r, t = _random_pose(rs)
        gr, gtt = r + rs.normal(size=3) * 0.05, t + rs.normal(size=3)
        gt = (gr.reshape(3, 1), gtt.reshape(3, 1))

        def val(x):
            est = fwd.pose2trans(x[:3].reshape(3, 1), x[3:].reshape(3, 1)).astype(np.float64)
            rot_diff = est[:3, :3].T @ fwd.pose2trans(*gt)[:3, :3].astype(np.float64)   # degrees with CV_PI, as dLoss
            tr = min(3.0, max(-1.0, float(np.trace(rot_diff))))
            return 180 * math.acos((tr - 1) / 2) / math.pi + float(np.linalg.norm(est[:3, 3] - fwd.pose2trans(*gt)[:3, 3]))

        x0 = np.concatenate([r, t])
        num = np.array([(val(x0 + e) - val(x0 - e)) / 2e-6 for e in np.eye(6) * 1e-6])
        jac = bwd.d_loss((r.reshape(3, 1), t.reshape(3, 1)), gt, 1.0, 1.0, 1e9).ravel()
        assert np.abs(num - jac).max() < 1e-4 * max(1.0, np.abs(jac).max())


def test_cuda_backward_algebra_matches_opencv(hostmath):
    lib = hostmath
    rs = np.random.default_rng(0)
    k = fwd.cam_mat(480., 360., 240.)
    worst = dict(loss=0., jac=0., t2p=0., dpo=0., row=0., pinv=0.)
    for _ in range(150):
        r, t = _random_pose(rs)
        gr = r + rs.normal(size=3) * rs.choice([1e-3, 0.05, 1.0])
        gtt = t + rs.normal(size=3) * rs.choice([0.01, 1, 50])
        est = (r.reshape(3, 1).copy(), t.reshape(3, 1).copy())
        gt_t = fwd.pose2trans(gr.reshape(3, 1), gtt.reshape(3, 1)).astype(np.float32).astype(np.float64)
        cut = float(rs.choice([100.0, 5.0]))
        rt = np.concatenate([r, t])
        a = bwd.loss(fwd.pose2trans(*est).astype(np.float64), gt_t, 1.0, 1.0, cut)
        b = lib.hm_pose_loss(_p(rt), _p(np.ascontiguousarray(gt_t)), D(1.0), D(1.0), D(cut))
        worst['loss'] = max(worst['loss'], abs(a - b) / max(1, abs(a)))
        hgt = bwd.trans2pose(gt_t)
        hgt6 = np.concatenate([hgt[0].ravel(), hgt[1].ravel()])
        rt2 = np.zeros(6)
        lib.hm_trans_to_pose(_p(np.ascontiguousarray(gt_t)), _p(rt2))
        worst['t2p'] = max(worst['t2p'], np.abs(rt2 - hgt6).max() / 100)
        ja = bwd.d_loss(est, hgt, 1.0, 1.0, cut).ravel()
        jb = np.zeros(6)
        lib.hm_pose_loss_jacobian(_p(rt), _p(hgt6), D(1.0), D(1.0), D(cut), _p(jb))
        worst['jac'] = max(worst['jac'], np.abs(ja - jb).max() / max(1e-9, np.abs(ja).max()))
        rot, _ = cv2.Rodrigues(r)
        x = (rot.T @ (np.array([rs.uniform(-3, 3), rs.uniform(-2, 2), rs.uniform(5, 50)]) - t)).astype(np.float32)
        pt = np.array([rs.integers(0, 90) * 8 + 4, rs.integers(0, 60) * 8 + 4], dtype=np.float32)
        oa = bwd.d_project_d_obj(pt, x, rot, t.reshape(3, 1), k, 100.0).ravel()
        ob = np.zeros(3)
        lib.hm_d_project_d_obj(F(pt[0]), F(pt[1]), F(x[0]), F(x[1]), F(x[2]), _p(np.ascontiguousarray(rot)), _p(t), D(480.),
                               D(360.), D(240.), F(100.), _p(ob))
        worst['dpo'] = max(worst['dpo'], np.abs(oa - ob).max() / max(1e-12, np.abs(oa).max()))
        proj, jac = cv2.projectPoints(x.reshape(1, 3), r.reshape(3, 1), t.reshape(3, 1), k, None)
        proj = proj.reshape(2).astype(np.float32)
        jac = np.asarray(jac)[:, :6]
        diff = (proj - pt).astype(np.float64)
        err = max(math.sqrt(float((diff ** 2).sum())), 1e-8)
        ra = np.zeros(6) if err > 100 else (diff / err) @ jac
        rb = np.zeros(6)
        lib.hm_residual_row(_p(rt), D(480.), D(360.), D(240.), F(x[0]), F(x[1]), F(x[2]), int(pt[0]), int(pt[1]), F(100.), _p(rb))
        worst['row'] = max(worst['row'], np.abs(ra - rb).max() / max(1e-12, np.abs(ra).max()))
        j = rs.normal(size=(40, 6)) * np.array([100, 100, 100, 1, 1, 1.])
        a66 = j.T @ j
        _, inv = cv2.invert(a66, flags=cv2.DECOMP_SVD)
        pm = np.zeros((6, 6))
        lib.hm_sym6_pinv(_p(np.ascontiguousarray(a66)), _p(pm))
        worst['pinv'] = max(worst['pinv'], np.abs(inv - pm).max() / np.abs(inv).max())
    assert worst['loss'] < 1e-9 and worst['t2p'] < 1e-12 and worst['jac'] < 1e-7, worst
    assert worst['dpo'] < 1e-11 and worst['row'] < 1e-11 and worst['pinv'] < 1e-11, worst


def test_sym6_pinv_drops_the_null_space_like_cv_invert(hostmath):
    """Rank-deficient normal matrix: cv::invert(DECOMP_SVD) returns the pseudo-inverse, so does the device routine."""
    rs = np.random.default_rng(2)
    j = rs.normal(size=(30, 4)) @ rs.normal(size=(4, 6))
    a66 = j.T @ j
    _, inv = cv2.invert(a66, flags=cv2.DECOMP_SVD)
    pm = np.zeros((6, 6))
    hostmath.hm_sym6_pinv(_p(np.ascontiguousarray(a66)), _p(pm))
    assert np.abs(inv - pm).max() < 1e-8 * np.abs(inv).max()


def test_backward_fixture_table_in_sync():
    mod = backward_module()
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'dsac_backward_golden.npz'))
    assert {k.split('_')[0] for k in g.files} == {'case%d' % i for i in range(len(mod.CASES))}
