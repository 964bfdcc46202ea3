"""TEST INFRASTRUCTURE ONLY -- tier-1 CPU oracle for the DSAC* RGB backward pass.

A line-by-line restatement of `dsacstar_rgb_backward` on top of the Python `cv2` module, calling the OpenCV entry
points the reference C++ calls (projectPoints with its Jacobian, solvePnP P3P / ITERATIVE, Rodrigues with its
Jacobian, invert(DECOMP_SVD)).  Nothing under crossloc_b200/, dsacstar/, networks/ or loss/ imports this file; only
tests/ and tests/golden/make_dsac_backward_golden.py do.

PARITY UNPINNED, for the same reason as oracle/dsac_oracle_py.py: the reference holds no vectors for this path and its
extension cannot be built here (OpenCV 3.4.2 C++ absent), so the anchor is the reference's own call sequence executed
through cv2 4.13.

Followed reference code (all under /root/reference/dsacstar/):
  dsacstar.cpp:200-483           dsacstar_rgb_backward         -> backward_rgb
  dsacstar_util.h:356-446        getReproErrs (calcJ = true)    -> repro_errs_jacobian
  dsacstar_util.h:777-790        trans2pose                     -> trans2pose
  dsacstar_util.h:821-834        getMax                         -> get_max
  dsacstar_loss.h:47-84          calcAngularDistance / loss     -> loss
  dsacstar_loss.h:96-212         dLoss                          -> d_loss
  dsacstar_derivative.h:50-109   dProjectdObj                   -> d_project_d_obj
  dsacstar_derivative.h:137-196  dPNP                           -> d_pnp
  dsacstar_derivative.h:214-331  dScore                         -> d_score
  dsacstar_derivative.h:351-414  dSMScore                       -> d_sm_score
The random source is the Philox stream of crossloc_b200/rng.py, as in the forward oracle.
"""
import math

import cv2
import numpy as np

from oracle import dsac_oracle_py as fwd

PROB_THRESH = 0.001    # dsacstar_derivative.h:36
MAXLOSS = 10000000.0   # dsacstar_loss.h:35
EPS = fwd.EPS
PI = 3.1415926         # dsacstar_util.h:46


def get_max(mat):
    """dsacstar_util.h:821-834: largest absolute entry (-1 for an empty matrix)."""
    return float(np.abs(mat).max()) if mat.size else -1.0


def trans2pose(trans):
    """dsacstar_util.h:777-790: camera-to-world 4x4 -> (rvec, tvec) of the scene-to-camera transform."""
    inv = np.linalg.inv(np.asarray(trans, dtype=np.float64))
    rvec, _ = cv2.Rodrigues(inv[:3, :3])
    return rvec.reshape(3, 1), inv[:3, 3].reshape(3, 1).copy()


def loss(trans1, trans2, w_rot=1.0, w_trans=1.0, cut=100.0):
    """dsacstar_loss.h:47-84."""
    rot_diff = trans2[:3, :3] @ trans1[:3, :3].T
    trace = min(3.0, max(-1.0, float(np.trace(rot_diff))))
    rot_err = 180 * math.acos((trace - 1.0) / 2.0) / PI
    t_err = float(np.linalg.norm(trans1[:3, 3] - trans2[:3, 3]))
    val = w_rot * rot_err + w_trans * t_err
    if val > cut:
        val = math.sqrt(cut * val)
    return min(val, MAXLOSS)


def d_loss(est, gt, w_rot=1.0, w_trans=1.0, cut=100.0):
    """dsacstar_loss.h:96-212: 1x6 Jacobian of the pose loss w.r.t. the estimated (rvec, tvec)."""
    rot1, d_rod = cv2.Rodrigues(est[0])           # d_rod: 3 x 9
    rot2, _ = cv2.Rodrigues(gt[0])
    inv_rot1, inv_rot2 = rot1.T.copy(), rot2.T.copy()
    diff_rot = rot1 @ inv_rot2
    trace = min(3.0, max(-1.0, float(np.trace(diff_rot))))
    rot_err = 180 * math.acos((trace - 1.0) / 2.0) / math.pi
    inv_t1 = inv_rot1 @ est[1].reshape(3, 1)
    inv_t2 = inv_rot2 @ gt[1].reshape(3, 1)
    t_err = float(np.linalg.norm(inv_t1 - inv_t2))
    jac = np.zeros((1, 6))
    val = w_rot * rot_err + w_trans * t_err
    cut_loss = False
    if val > cut:
        val = math.sqrt(val)
        cut_loss = True
    if val > MAXLOSS:
        return jac
    if (t_err + rot_err) < EPS:
        return jac
    d_dist_d_inv_t1 = ((inv_t1 - inv_t2) / t_err).reshape(1, 3)
    jac[:, 3:6] += d_dist_d_inv_t1 @ inv_rot1 * w_trans
    d_inv_t1_d_inv_rot1 = np.zeros((3, 9))
    t = est[1].reshape(3)
    for i in range(3):
        d_inv_t1_d_inv_rot1[i, i] = t[0]
        d_inv_t1_d_inv_rot1[i, i + 3] = t[1]
        d_inv_t1_d_inv_rot1[i, i + 6] = t[2]
    d_rod = d_rod.T                               # 9 x 3
    jac[:, 0:3] += d_dist_d_inv_t1 @ d_inv_t1_d_inv_rot1 @ d_rod * w_trans
    d_rot_diff = np.zeros((9, 9))
    for blk in range(3):
        for r in range(3):
            d_rot_diff[3 * blk + r, 3 * blk:3 * blk + 3] = inv_rot2[r]
    d_rot_diff = d_rot_diff.T
    d_trace = np.zeros((1, 9))
    d_trace[0, 0] = d_trace[0, 4] = d_trace[0, 8] = 1
    with np.errstate(divide='ignore', invalid='ignore'):
        d_angle = (180 / math.pi * -1 / np.sqrt(3 - trace * trace + 2 * trace)) * d_trace @ d_rot_diff @ d_rod
    jac[:, 0:3] += d_angle * w_rot
    if cut_loss:
        jac *= 0.5 / val
    if np.isnan(jac).any():
        return np.zeros((1, 6))
    return jac


def d_project_d_obj(pt, obj, rot, trans, k, max_repro_err):
    """dsacstar_derivative.h:50-109: 1x3 Jacobian of the reprojection error w.r.t. the 3-D point."""
    f, ppx, ppy = float(k[0, 0]), float(k[0, 2]), float(k[1, 2])
    o = rot @ np.asarray(obj, dtype=np.float64).reshape(3, 1) + trans.reshape(3, 1)
    x, y, z = float(o[0, 0]), float(o[1, 0]), float(o[2, 0])
    if abs(z) < EPS:
        return np.zeros((1, 3))
    px, py = f * x / z + ppx, f * y / z + ppy
    ptx, pty = float(pt[0]), float(pt[1])
    err = math.sqrt((ptx - px) * (ptx - px) + (pty - py) * (pty - py))
    if err > max_repro_err:
        return np.zeros((1, 3))
    err += EPS
    out = np.zeros((1, 3))
    for j in range(3):
        pxd = f * rot[0, j] / z - f * x / z / z * rot[2, j]
        pyd = f * rot[1, j] / z - f * y / z / z * rot[2, j]
        out[0, j] = 0.5 / err * (2 * (ptx - px) * -pxd + 2 * (pty - py) * -pyd)
    return out


def d_pnp(img_pts, obj_pts, k, eps=np.float32(0.001)):
    """dsacstar_derivative.h:137-196 for the 4-point (P3P) case: 6x12 central-difference Jacobian of the pose w.r.t. the
    minimal set (the 4th point only resolves the ambiguity: its columns stay zero)."""
    obj = np.array(obj_pts, dtype=np.float32)
    img = np.array(img_pts, dtype=np.float32)
    jac = np.zeros((6, obj.shape[0] * 3))
    for i in range(3):
        for j in range(3):
            obj[i, j] += eps
            ok, fr, ft = fwd._safe_solve_pnp(obj, img, k, None, None, False, cv2.SOLVEPNP_P3P)
            if not ok:
                return np.zeros((6, obj.shape[0] * 3))
            obj[i, j] -= 2 * eps
            ok, br, bt = fwd._safe_solve_pnp(obj, img, k, None, None, False, cv2.SOLVEPNP_P3P)
            if not ok:
                return np.zeros((6, obj.shape[0] * 3))
            obj[i, j] += eps
            col = np.concatenate([(fr - br) / (2 * float(eps)), (ft - bt) / (2 * float(eps))]).reshape(6)
            jac[:, i * 3 + j] = col
            if np.isnan(col).any():
                return np.zeros((6, obj.shape[0] * 3))
    return jac


def repro_errs_jacobian(coords, rvec, tvec, sampling, k, max_reproj):
    """dsacstar_util.h:356-446 with calcJ = true.  Returns (errs f32 [H, W], jacobeanHyp f64 [H, W, 6])."""
    h, w = coords.shape[1:]
    pts3 = coords.reshape(3, -1).T.astype(np.float32)
    pts2 = sampling.reshape(-1, 2).astype(np.float32)
    proj, jac = cv2.projectPoints(pts3, rvec, tvec, k, None)
    proj = proj.reshape(-1, 2).astype(np.float32)
    jac = np.asarray(jac, dtype=np.float64)[:, 0:6].reshape(-1, 2, 6)
    diff = (proj - pts2).astype(np.float64)                   # Point2f difference
    err = np.maximum(np.sqrt(diff[:, 0] ** 2 + diff[:, 1] ** 2), EPS)
    jh = np.zeros((pts3.shape[0], 6))
    keep = ~(err > max_reproj)
    dndp = diff[keep] / err[keep, None]
    jh[keep] = np.einsum('nk,nkj->nj', dndp, jac[keep])
    errs = np.minimum(fwd._norm2f(pts2 - proj).astype(np.float32), np.float32(max_reproj))
    return errs.reshape(h, w), jh.reshape(h, w, 6)


def refine_hyp_with_inliers(coords, errs, sampling, k, thr, max_ref_steps, max_reproj, rvec, tvec):
    """dsacstar_util.h:522-597 keeping the inlier map of the last accepted step (the one the backward pass needs)."""
    thr = np.float32(thr)
    local = errs.copy()
    best = 4
    inlier_map = None
    for _ in range(max_ref_steps):
        mask = local < thr
        ys, xs = np.nonzero(mask.T)[1], np.nonzero(mask.T)[0]
        n = len(xs)
        if n <= best:
            break
        best = n
        img = sampling[ys, xs].astype(np.float32)
        obj = coords[:, ys, xs].T.astype(np.float32)
        flag = cv2.SOLVEPNP_ITERATIVE if n > 4 else cv2.SOLVEPNP_P3P
        ok, r_new, t_new = fwd._safe_solve_pnp(np.ascontiguousarray(obj), np.ascontiguousarray(img), k, rvec, tvec, True, flag)
        if not ok:
            break
        rvec, tvec = r_new, t_new
        inlier_map = mask.copy()
        local = fwd.repro_errs(coords, rvec, tvec, sampling, k, max_reproj)
    return rvec, tvec, inlier_map


def backward_rgb(coords, gt_pose, hyps, thr, focal, cx, cy, w_loss_rot, w_loss_trans, soft_clamp, alpha, max_reproj,
                 sub_sampling, sampler=None, seed=1305, image=0, max_tries=fwd.MAX_HYPOTHESES_TRIES):
    """dsacstar.cpp:200-483.  coords f32 [3, Hc, Wc]; gt_pose [4, 4] camera-to-world.
    Returns dict(loss = expected pose loss, grad f64 [3, Hc, Wc], probs, losses, hyps_rt, ref_rt, tries)."""
    coords = np.asarray(coords, dtype=np.float32)
    if coords.ndim == 4:
        coords = coords[0]
    hc, wc = coords.shape[1:]
    k = fwd.cam_mat(focal, cx, cy)
    sampling = fwd.create_sampling(wc, hc, sub_sampling)
    if sampler is None:
        sampler = fwd.default_sampler(seed, image, wc, hc)
    gt_trans = np.asarray(gt_pose, dtype=np.float32).astype(np.float64)

    init, cells, tries = fwd.sample_hypotheses(coords, sampling, k, hyps, max_tries, thr, sampler)
    errs, jac_hyp = [], []
    for (r, t) in init:
        e, j = repro_errs_jacobian(coords, r, t, sampling, k, max_reproj)
        errs.append(e)
        jac_hyp.append(j)
    scores = fwd.hyp_scores(errs, thr, alpha)
    probs = fwd.soft_max(scores)

    ref, inlier_maps = [], []
    for h in range(hyps):
        r, t = init[h][0].copy(), init[h][1].copy()
        imap = None
        if probs[h] >= PROB_THRESH:
            r, t, imap = refine_hyp_with_inliers(coords, errs[h], sampling, k, thr, fwd.MAX_REF_STEPS, max_reproj, r, t)
        ref.append((r, t))
        inlier_maps.append(imap)

    losses = np.zeros(hyps)
    expected = 0.0
    for h in range(hyps):
        losses[h] = loss(fwd.pose2trans(*ref[h]), gt_trans, w_loss_rot, w_loss_trans, soft_clamp)
        expected += probs[h] * losses[h]

    grad = np.zeros((hc, wc, 3))
    hyp_gt = trans2pose(gt_trans)
    # ---- path I: through the refined hypotheses
    for h in range(hyps):
        if probs[h] < PROB_THRESH:
            continue
        imap = inlier_maps[h]
        if imap is None:
            continue
        ys, xs = np.nonzero(imap.T)[1], np.nonzero(imap.T)[0]       # x outer, y inner (dsacstar.cpp:366-379)
        if len(xs) < 4:
            continue
        img = sampling[ys, xs].astype(np.float32)
        obj = coords[:, ys, xs].T.astype(np.float32)
        proj, pj = cv2.projectPoints(obj, ref[h][0], ref[h][1], k, None)
        proj = proj.reshape(-1, 2).astype(np.float32)
        pj = np.asarray(pj, dtype=np.float64)[:, 0:6].reshape(-1, 2, 6)
        diff = (proj - img).astype(np.float64)
        err = np.maximum(np.sqrt(diff[:, 0] ** 2 + diff[:, 1] ** 2), EPS)
        jr = np.zeros((len(xs), 6))
        keep = ~(err > max_reproj)
        jr[keep] = np.einsum('nk,nkj->nj', diff[keep] / err[keep, None], pj[keep])
        ok, inv = cv2.invert(jr.T @ jr, flags=cv2.DECOMP_SVD)
        jr = -inv @ jr.T                                           # 6 x n
        if get_max(jr) > 10:
            jr = np.zeros_like(jr)
        rot, _ = cv2.Rodrigues(ref[h][0])
        d_loss_d_hyp = d_loss(ref[h], hyp_gt, w_loss_rot, w_loss_trans, soft_clamp)      # 1 x 6
        for p in range(len(xs)):
            dndo = d_project_d_obj(img[p], obj[p], rot, ref[h][1], k, max_reproj)        # 1 x 3
            d_hyp_d_obj = jr[:, p:p + 1] @ dndo                                           # 6 x 3
            grad[ys[p], xs[p]] += probs[h] * (d_loss_d_hyp @ d_hyp_d_obj).reshape(3)
    # ---- path II: through the scores
    thr32 = np.float32(thr)
    beta = np.float32(5) / thr32
    score_grads = np.zeros(hyps)
    for i in range(hyps):
        if probs[i] < PROB_THRESH:
            continue
        score_grads[i] = probs[i] * losses[i] - probs[i] * float(np.dot(probs, losses))
    for h in range(hyps):
        if probs[h] < PROB_THRESH:
            continue
        soft = (beta * (errs[h] - thr32)).astype(np.float64)
        soft = 1 / (1 + np.exp(-soft))
        d_repro = -soft * (1 - soft) * float(beta) * score_grads[h]
        d_repro = d_repro * float(np.float32(alpha) / np.float32(wc) / np.float32(hc))
        pts = [tuple(c) for c in cells[h]]
        img4 = np.array([sampling[y, x] for (x, y) in pts], dtype=np.float32)
        obj4 = np.array([coords[:, y, x] for (x, y) in pts], dtype=np.float32)
        dhdo = d_pnp(img4, obj4, k)
        if get_max(dhdo) > 10:
            dhdo = np.zeros_like(dhdo)
        rot, _ = cv2.Rodrigues(init[h][0])
        support = np.zeros((1, 12))
        for x in range(wc):
            for y in range(hc):
                pt = sampling[y, x].astype(np.float32)
                dpdo = d_project_d_obj(pt, coords[:, y, x], rot, init[h][1], k, max_reproj) * d_repro[y, x]
                grad[y, x] += dpdo.reshape(3)
                support += d_repro[y, x] * jac_hyp[h][y, x].reshape(1, 6) @ dhdo
        for i, (x, y) in enumerate(pts):
            grad[y, x] += support[0, i * 3:i * 3 + 3]
    return {
        'loss': expected, 'grad': np.ascontiguousarray(grad.transpose(2, 0, 1)), 'probs': probs, 'losses': losses,
        'hyps_rt': np.array([np.concatenate([r.reshape(3), t.reshape(3)]) for (r, t) in init]),
        'ref_rt': np.array([np.concatenate([r.reshape(3), t.reshape(3)]) for (r, t) in ref]),
        'tries': tries, 'cells': cells, 'scores': scores,
    }
