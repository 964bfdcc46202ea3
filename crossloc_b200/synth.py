"""Synthetic "naturescape-shaped" localization scenes.

The reference ships no data and no fixtures, so every parity test and the bench use
scenes generated here: a smooth height-field terrain seen from a drone-like camera,
rendered into the [3, Hc, Wc] scene-coordinate map the network is trained to regress
(cell centres at (x*S + S/2, y*S + S/2), /root/reference/dsacstar/dsacstar_util.h:59-76;
principal point at the image centre, /root/reference/utils/evaluation.py:168-169).

Pure numpy: the generator is host-side test/bench plumbing, not part of the hot path.
"""
import numpy as np

NATURESCAPE_MEAN = np.array([-455.934, 417.50, 520.31])  # /root/reference/utils/learning.py:92


def _terrain(x, y):
    return (NATURESCAPE_MEAN[2]
            + 18.0 * np.sin(x * 0.011 + 0.3) * np.cos(y * 0.013 - 0.2)
            + 7.0 * np.sin(x * 0.031 + y * 0.027))


def _rodrigues(rvec):
    theta = np.linalg.norm(rvec)
    if theta < 1e-12:
        return np.eye(3)
    k = rvec / theta
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(theta) * kx + (1 - np.cos(theta)) * (kx @ kx)


def make_scene(index, height=480, width=720, subsample=8, focal=480.0,
               noise_sigma=0.5, outlier_ratio=0.3, outlier_sigma=50.0, nodata_ratio=0.05):
    """One synthetic frame.

    Returns a dict with
      coords   float32 [3, Hc, Wc]  noisy scene-coordinate map fed to the solver
      gt       float32 [3, Hc, Wc]  clean ground truth, nodata cells set to -1
      pose     float64 [4, 4]       camera-to-world ground truth
      focal    float
    """
    rng = np.random.default_rng(1000 + index)
    hc = -(-height // subsample)
    wc = -(-width // subsample)

    centre = NATURESCAPE_MEAN[:2] + rng.uniform(-250.0, 250.0, size=2)
    altitude = rng.uniform(80.0, 150.0)
    cam_pos = np.array([centre[0], centre[1], _terrain(centre[0], centre[1]) + altitude])

    # nadir-looking OpenCV camera (x right, y down, z forward = world -Z), tilted by up to 30 deg
    base = np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    tilt = _rodrigues(axis * np.deg2rad(rng.uniform(0.0, 30.0)))
    rot = tilt @ base  # camera-to-world rotation

    xs = np.arange(wc) * subsample + subsample // 2
    ys = np.arange(hc) * subsample + subsample // 2
    u, v = np.meshgrid(xs, ys)
    rays_cam = np.stack([(u - width / 2) / focal, (v - height / 2) / focal, np.ones_like(u, dtype=np.float64)], 0)
    rays = np.einsum('ij,jhw->ihw', rot, rays_cam)

    # ray / height-field intersection by fixed-point iteration on the depth
    depth = np.full((hc, wc), altitude, dtype=np.float64)
    for _ in range(40):
        p = cam_pos[:, None, None] + depth * rays
        depth = depth + (p[2] - _terrain(p[0], p[1])) / np.maximum(-rays[2], 0.2)
        depth = np.clip(depth, 1.0, 2000.0)
    gt = cam_pos[:, None, None] + depth * rays

    coords = gt + rng.normal(scale=noise_sigma, size=gt.shape)
    outlier = rng.random((hc, wc)) < outlier_ratio
    coords = np.where(outlier[None], gt + rng.normal(scale=outlier_sigma, size=gt.shape), coords)

    gt_out = gt.copy()
    nodata = rng.random((hc, wc)) < nodata_ratio
    gt_out[:, nodata] = -1.0

    pose = np.eye(4)
    pose[:3, :3] = rot
    pose[:3, 3] = cam_pos
    return {
        'coords': coords.astype(np.float32),
        'gt': gt_out.astype(np.float32),
        'pose': pose,
        'focal': float(focal),
    }


def make_batch(first_index, count, **kw):
    """Stack ``count`` scenes: coords [B,3,Hc,Wc] f32, gt [B,3,Hc,Wc] f32, poses [B,4,4] f64, focal [B] f32."""
    scenes = [make_scene(first_index + i, **kw) for i in range(count)]
    return (np.stack([s['coords'] for s in scenes]),
            np.stack([s['gt'] for s in scenes]),
            np.stack([s['pose'] for s in scenes]),
            np.array([s['focal'] for s in scenes], dtype=np.float32))


def pose_errors(gt_pose, est_pose):
    """Translation (m) and rotation (deg) error as /root/reference/utils/evaluation.py:121-132."""
    gt_pose = np.asarray(gt_pose, dtype=np.float64)
    est_pose = np.asarray(est_pose, dtype=np.float64)
    t_err = float(np.linalg.norm(gt_pose[:3, 3] - est_pose[:3, 3]))
    r = est_pose[:3, :3].T @ gt_pose[:3, :3]
    # angle of the relative rotation as cv2.Rodrigues gives it: atan2 of the skew part keeps small angles exact
    sin = 0.5 * np.sqrt((r[2, 1] - r[1, 2]) ** 2 + (r[0, 2] - r[2, 0]) ** 2 + (r[1, 0] - r[0, 1]) ** 2)
    cos = (np.trace(r) - 1.0) / 2.0
    return t_err, float(np.degrees(np.arctan2(sin, cos)))
