"""Generates tests/golden/dsac_backward_golden.npz from the tier-1 backward oracle (oracle/dsac_backward_py.py, cv2).

As for the forward pass (make_golden.py) the reference holds no vectors for dsacstar_rgb_backward and its extension
cannot be built here, so these fixtures pin the restated algorithm executed through the OpenCV entry points the
reference calls.  Inputs are regenerated deterministically by crossloc_b200.synth; only outputs are stored.

    python tests/golden/make_backward_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from crossloc_b200 import synth  # noqa: E402

# (scene index, hypotheses, scene kwargs, solver kwargs) -- must stay in sync with tests/util.py BACKWARD_CASES
CASES = [
    (0, 64, {}, {}),
    (1, 32, {'outlier_ratio': 0.4}, {'alpha': 10.0}),                 # flat distribution: many hypotheses count
    (5, 32, {'height': 240, 'width': 368}, {'alpha': 20.0}),          # ragged size: 30 x 46 cells
    (6, 16, {'height': 240, 'width': 368}, {'soft_clamp': 0.01, 'w_rot': 2.0, 'w_trans': 0.5, 'alpha': 5.0}),   # clamped loss branch
]
PARAMS = dict(thr=10.0, alpha=100.0, max_reproj=100.0, sub_sampling=8, seed=1305, w_rot=1.0, w_trans=1.0, soft_clamp=100.0)


def gt_pose_for(scene, idx):
    """Ground truth = the scene's pose moved by a small rigid offset, so that the loss and dLoss are not degenerate."""
    rs = np.random.default_rng(1000 + idx)
    ax = rs.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(0.5)
    k = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    d = np.eye(4)
    d[:3, :3] = np.eye(3) + np.sin(ang) * k + (1 - np.cos(ang)) * k @ k
    d[:3, 3] = rs.normal(size=3) * 0.3
    return (np.asarray(scene['pose'], dtype=np.float64) @ d).astype(np.float32)


def run_case(ci):
    idx, hyps, kw, skw = CASES[ci]
    s = synth.make_scene(idx, **kw)
    h, w = kw.get('height', 480), kw.get('width', 720)
    p = dict(PARAMS)
    p.update(skw)
    gt = gt_pose_for(s, idx)
    from oracle import dsac_backward_py as tier1
    r = tier1.backward_rgb(s['coords'], gt, hyps, p['thr'], s['focal'], w / 2, h / 2, p['w_rot'], p['w_trans'],
                           p['soft_clamp'], p['alpha'], p['max_reproj'], p['sub_sampling'], seed=p['seed'], image=idx)
    return s, gt, p, r


def main():
    out = {}
    for ci in range(len(CASES)):
        s, gt, p, r = run_case(ci)
        out['case%d_loss' % ci] = np.float64(r['loss'])
        out['case%d_grad' % ci] = r['grad']
        out['case%d_probs' % ci] = r['probs']
        out['case%d_losses' % ci] = r['losses']
        out['case%d_ref_rt' % ci] = r['ref_rt']
        out['case%d_hyps_rt' % ci] = r['hyps_rt']
        out['case%d_tries' % ci] = np.asarray(r['tries'], dtype=np.int32)
        out['case%d_cells' % ci] = np.asarray(r['cells'], dtype=np.int32).reshape(-1, 4, 2)
        print(ci, 'loss %.6f' % r['loss'], 'hyps >= 1e-3:', int((r['probs'] >= 1e-3).sum()), 'max |grad| %.4g' % np.abs(r['grad']).max(),
              'nonzero cells', int((np.abs(r['grad']).sum(0) > 0).sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dsac_backward_golden.npz'), **out)


if __name__ == '__main__':
    main()
