"""TEST INFRASTRUCTURE ONLY -- tier-1 CPU oracle for the DSAC* RGB forward path.

A line-by-line restatement of the reference solver on top of the Python `cv2` module,
calling the very OpenCV entry points the reference C++ calls (solvePnP P3P / ITERATIVE,
projectPoints, Rodrigues).  Nothing under crossloc_b200/, dsacstar/, networks/ or loss/
imports this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for this path
 and its extension cannot be built here (OpenCV 3.4.2 C++ headers and
libraries are absent; pinned in /root/reference/setup/environment.yml:11).  This oracle
therefore anchors on the reference's own call sites, executed through cv2 4.13 -- the
same algorithm family, but not the pinned version.

Followed reference code (all under /root/reference/dsacstar/):
  dsacstar.cpp:63-178          dsacstar_rgb_forward      -> forward_rgb
  dsacstar_util.h:59-76        createSampling            -> create_sampling
  dsacstar_util.h:91-120       safeSolvePnP              -> _safe_solve_pnp
  dsacstar_util.h:135-221      sampleHypotheses          -> sample_hypotheses
  dsacstar_util.h:316-343      getHypScores              -> hyp_scores
  dsacstar_util.h:356-446      getReproErrs              -> repro_errs
  dsacstar_util.h:522-597      refineHyp                 -> refine_hyp
  dsacstar_util.h:684-752      softMax / entropy / draw  -> soft_max / entropy / draw_argmax
  dsacstar_util.h:759-770      pose2trans                -> pose2trans

The one deliberate difference is the random source: cells are drawn by a caller-supplied
``sampler(hyp, try) -> [(x, y)] * 4`` (default: crossloc_b200.rng's Philox stream) instead of
the per-thread mt19937 (thread_rand.cpp:13-71), so that another implementation can replay
exactly the same draws.
"""
import math

import cv2
import numpy as np

MAX_REF_STEPS = 100              # dsacstar.cpp:47
MAX_HYPOTHESES_TRIES = 1000000   # dsacstar.cpp:48
EPS = 0.00000001                 # dsacstar_util.h:45


def create_sampling(out_w, out_h, sub_sampling, shift_x=0, shift_y=0):
    """dsacstar_util.h:59-76 -- int pixel position of every cell, [out_h, out_w, 2] (x, y)."""
    xs = np.arange(out_w, dtype=np.int32) * sub_sampling + sub_sampling // 2 - shift_x
    ys = np.arange(out_h, dtype=np.int32) * sub_sampling + sub_sampling // 2 - shift_y
    grid = np.empty((out_h, out_w, 2), dtype=np.int32)
    grid[..., 0] = xs[None, :]
    grid[..., 1] = ys[:, None]
    return grid


def cam_mat(focal, cx, cy):
    """dsacstar.cpp:86-90 -- float32 calibration matrix."""
    k = np.eye(3, dtype=np.float32)
    k[0, 0] = focal
    k[1, 1] = focal
    k[0, 2] = cx
    k[1, 2] = cy
    return k


def _safe_solve_pnp(obj_pts, img_pts, k, rvec, tvec, extrinsic_guess, flag):
    """dsacstar_util.h:91-120 -- returns (ok, rvec[3,1] f64, tvec[3,1] f64); zeros on failure."""
    try:
        if extrinsic_guess:
            ok, r, t = cv2.solvePnP(obj_pts, img_pts, k, None, rvec.copy(), tvec.copy(), True, flag)
        else:
            ok, r, t = cv2.solvePnP(obj_pts, img_pts, k, None, flags=flag)
    except cv2.error:
        ok = False
    if not ok:
        return False, np.zeros((3, 1)), np.zeros((3, 1))
    return True, np.asarray(r, dtype=np.float64).reshape(3, 1), np.asarray(t, dtype=np.float64).reshape(3, 1)


def _project(obj_pts, rvec, tvec, k):
    """cv::projectPoints into vector<Point2f> (float32 result), no distortion."""
    proj, _ = cv2.projectPoints(obj_pts.astype(np.float32), rvec, tvec, k, None)
    return proj.reshape(-1, 2).astype(np.float32)


def _norm2f(diff):
    """cv::norm(Point2f): sqrt in double of float components."""
    d = diff.astype(np.float64)
    return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])


def sample_hypotheses(coords, sampling, k, hyps, max_tries, thr, sampler):
    """dsacstar_util.h:135-221.  coords [3,H,W] f32.  Returns list of (rvec, tvec), cells [hyps,4,2], tries."""
    out = []
    cells = np.zeros((hyps, 4, 2), dtype=np.int32)
    tries = np.zeros(hyps, dtype=np.int64)
    thr = np.float32(thr)
    for h in range(hyps):
        rvec = np.zeros((3, 1))
        tvec = np.zeros((3, 1))
        t = 0
        while t < max_tries:
            pts = sampler(h, t)
            t += 1
            img = np.array([sampling[y, x] for (x, y) in pts], dtype=np.float32)
            obj = np.array([coords[:, y, x] for (x, y) in pts], dtype=np.float32)
            cells[h] = np.array(pts, dtype=np.int32)
            ok, rvec, tvec = _safe_solve_pnp(obj, img, k, None, None, False, cv2.SOLVEPNP_P3P)
            if not ok:
                continue
            proj = _project(obj, rvec, tvec, k)
            if np.all(_norm2f(img - proj) < thr):   # strict <, dsacstar_util.h:210
                break
        tries[h] = t
        out.append((rvec, tvec))
    return out, cells, tries


def repro_errs(coords, rvec, tvec, sampling, k, max_reproj):
    """dsacstar_util.h:356-446 (calcJ = false).  Returns float32 [H, W]."""
    h, w = coords.shape[1:]
    pts3 = coords.reshape(3, -1).T.astype(np.float32)            # row-major order; order is irrelevant here
    pts2 = sampling.reshape(-1, 2).astype(np.float32)
    proj = _project(pts3, rvec, tvec, k)
    err = _norm2f(pts2 - proj).astype(np.float32)
    return np.minimum(err, np.float32(max_reproj)).reshape(h, w)


def hyp_scores(errs_list, thr, alpha):
    """dsacstar_util.h:316-343 -- float beta, double accumulation, float scale."""
    thr = np.float32(thr)
    beta = np.float32(5) / thr
    scores = []
    for e in errs_list:
        soft = (beta * (e - thr)).astype(np.float64)          # float arithmetic, widened to double
        soft = 1.0 / (1.0 + np.exp(-soft))
        s = float(np.sum(1.0 - soft))
        scale = np.float32(alpha) / np.float32(e.shape[1]) / np.float32(e.shape[0])
        scores.append(s * float(scale))
    return np.array(scores, dtype=np.float64)


def soft_max(scores):
    """dsacstar_util.h:684-704."""
    m = np.max(scores)
    sf = np.exp(scores - m)
    return sf / np.sum(sf)


def entropy(dist):
    """dsacstar_util.h:711-719."""
    d = dist[dist > 0]
    return float(-np.sum(d * np.log2(d)))


def draw_argmax(probs):
    """dsacstar_util.h:727-752 with training = false: first maximal entry among probs >= EPS."""
    max_prob, max_idx = -1.0, 0
    for i, p in enumerate(probs):
        if p < EPS:
            continue
        if max_prob < 0 or p > max_prob:
            max_prob, max_idx = p, i
    return max_idx


def refine_hyp(coords, errs, sampling, k, thr, max_ref_steps, max_reproj, rvec, tvec):
    """dsacstar_util.h:522-597.  Returns refined (rvec, tvec), list of inlier counts per step."""
    thr = np.float32(thr)
    local = errs.copy()
    best = 4
    counts = []
    for _ in range(max_ref_steps):
        mask = local < thr                                           # [H, W]
        # column-major gathering order (x outer, y inner), dsacstar_util.h:547-559
        ys, xs = np.nonzero(mask.T)[1], np.nonzero(mask.T)[0]
        n = len(xs)
        counts.append(n)
        if n <= best:
            break
        best = n
        img = sampling[ys, xs].astype(np.float32)
        obj = coords[:, ys, xs].T.astype(np.float32)
        flag = cv2.SOLVEPNP_ITERATIVE if n > 4 else cv2.SOLVEPNP_P3P
        ok, r_new, t_new = _safe_solve_pnp(np.ascontiguousarray(obj), np.ascontiguousarray(img), k,
                                           rvec, tvec, True, flag)
        if not ok:
            break
        rvec, tvec = r_new, t_new
        local = repro_errs(coords, rvec, tvec, sampling, k, max_reproj)
    return rvec, tvec, counts


def pose2trans(rvec, tvec):
    """dsacstar_util.h:759-770 -- inverse of [R(rvec) t; 0 1], i.e. camera-to-world."""
    rot, _ = cv2.Rodrigues(rvec)
    trans = np.eye(4)
    trans[:3, :3] = rot
    trans[:3, 3] = tvec.reshape(3)
    return np.linalg.inv(trans)


def default_sampler(seed, image, width, height):
    from crossloc_b200.rng import sample_cells
    return lambda h, t: sample_cells(seed, image, h, t, width, height)


def forward_rgb(coords, hyps, thr, focal, cx, cy, alpha, max_reproj, sub_sampling,
                sampler=None, seed=1305, image=0, max_tries=MAX_HYPOTHESES_TRIES, refine=True):
    """dsacstar.cpp:63-178.  coords: float32 [3, Hc, Wc] (or [1,3,Hc,Wc]).

    Returns dict(pose f32 [4,4] camera-to-world, best, scores f64 [hyps], hyps_rt [hyps,6],
    cells, tries, refine_counts, rvec, tvec).
    """
    coords = np.asarray(coords, dtype=np.float32)
    if coords.ndim == 4:
        assert coords.shape[0] == 1, 'the reference supports batch size 1 only (dsacstar.cpp:52)'
        coords = coords[0]
    hc, wc = coords.shape[1:]
    k = cam_mat(focal, cx, cy)
    sampling = create_sampling(wc, hc, sub_sampling)
    if sampler is None:
        sampler = default_sampler(seed, image, wc, hc)

    hyp_list, cells, tries = sample_hypotheses(coords, sampling, k, hyps, max_tries, thr, sampler)
    errs = [repro_errs(coords, r, t, sampling, k, max_reproj) for (r, t) in hyp_list]
    scores = hyp_scores(errs, thr, alpha)
    probs = soft_max(scores)
    best = draw_argmax(probs)

    rvec, tvec = hyp_list[best]
    counts = []
    if refine:
        rvec, tvec, counts = refine_hyp(coords, errs[best], sampling, k, thr, MAX_REF_STEPS, max_reproj, rvec, tvec)
    pose = pose2trans(rvec, tvec).astype(np.float32)
    return {
        'pose': pose, 'best': best, 'scores': scores, 'entropy': entropy(probs),
        'hyps_rt': np.array([np.concatenate([r.reshape(3), t.reshape(3)]) for (r, t) in hyp_list]),
        'cells': cells, 'tries': tries, 'refine_counts': counts,
        'rvec': rvec.reshape(3).copy(), 'tvec': tvec.reshape(3).copy(),
    }
