"""Generates tests/golden/dsac_golden.npz from the tier-1 oracle (oracle/dsac_oracle_py.py on cv2).

The reference ships no golden vectors and its extension cannot be built here, so
these fixtures pin the *restated* algorithm executed through the very OpenCV entry points the reference
calls.  Inputs are regenerated deterministically by crossloc_b200.synth; only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from crossloc_b200 import synth  # noqa: E402
from oracle import dsac_oracle_py as tier1  # noqa: E402

CASES = [
    # (scene index, hypotheses, scene kwargs)
    (0, 64, {}),
    (1, 64, {}),
    (2, 64, {}),
    (3, 32, {'noise_sigma': 0.0, 'outlier_ratio': 0.0}),      # ground-truth map: known answer
    (4, 64, {'outlier_ratio': 0.6}),                         # hard: many retries
    (5, 64, {'height': 240, 'width': 368}),                  # ragged size: 30 x 46 cells
]
PARAMS = dict(thr=10.0, alpha=100.0, max_reproj=100.0, sub_sampling=8, seed=1305)


def main():
    out = {}
    for ci, (idx, hyps, kw) in enumerate(CASES):
        s = synth.make_scene(idx, **kw)
        h = kw.get('height', 480)
        w = kw.get('width', 720)
        r = tier1.forward_rgb(s['coords'], hyps, PARAMS['thr'], s['focal'], w / 2, h / 2, PARAMS['alpha'],
                              PARAMS['max_reproj'], PARAMS['sub_sampling'], seed=PARAMS['seed'], image=idx)
        out['case%d_pose' % ci] = r['pose']
        out['case%d_best' % ci] = np.int32(r['best'])
        out['case%d_scores' % ci] = r['scores']
        out['case%d_hyps_rt' % ci] = r['hyps_rt']
        out['case%d_tries' % ci] = r['tries'].astype(np.int32)
        out['case%d_counts' % ci] = np.array(r['refine_counts'], dtype=np.int32)
        out['case%d_rt' % ci] = np.concatenate([r['rvec'], r['tvec']])
        print(ci, idx, 'best', r['best'], 'counts', r['refine_counts'], 'err', synth.pose_errors(s['pose'], r['pose']))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dsac_golden.npz'), **out)


if __name__ == '__main__':
    main()
