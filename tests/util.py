"""Shared helpers of the test-suite."""
import os

import numpy as np

from crossloc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'dsac_golden.npz')

# must stay in sync with tests/golden/make_golden.py
GOLDEN_CASES = [
    (0, 64, {}),
    (1, 64, {}),
    (2, 64, {}),
    (3, 32, {'noise_sigma': 0.0, 'outlier_ratio': 0.0}),
    (4, 64, {'outlier_ratio': 0.6}),
    (5, 64, {'height': 240, 'width': 368}),
]
PARAMS = dict(thr=10.0, alpha=100.0, max_reproj=100.0, sub_sampling=8, seed=1305)


def golden_case(ci):
    idx, hyps, kw = GOLDEN_CASES[ci]
    g = np.load(GOLDEN)
    scene = synth.make_scene(idx, **kw)
    h, w = kw.get('height', 480), kw.get('width', 720)
    ref = {k[len('case%d_' % ci):]: g[k] for k in g.files if k.startswith('case%d_' % ci)}
    return idx, hyps, scene, (w / 2, h / 2), ref


def has_duplicate_cells(cells):
    """True when a minimal set holds the same cell twice: P3P's 4th-point disambiguation is then a tie."""
    c = np.asarray(cells).reshape(4, 2)
    return len({(int(x), int(y)) for x, y in c}) < 4


def score_mismatch(a, b, rel=1e-5):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b) > rel * max(1.0, float(np.max(np.abs(a))))


BACKWARD_GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'dsac_backward_golden.npz')


def backward_module():
    """tests/golden/make_backward_golden.py as a module (case table, parameters, ground-truth pose rule)."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location('make_backward_golden', os.path.join(ROOT, 'tests', 'golden', 'make_backward_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    spec.loader.exec_module(mod)
    sys.path[:] = saved
    return mod


def backward_case(ci):
    """Inputs (regenerated) and stored tier-1 outputs of backward fixture `ci` (tests/golden/make_backward_golden.py)."""
    mod = backward_module()
    idx, hyps, kw, skw = mod.CASES[ci]
    scene = synth.make_scene(idx, **kw)
    h, w = kw.get('height', 480), kw.get('width', 720)
    p = dict(mod.PARAMS)
    p.update(skw)
    g = np.load(BACKWARD_GOLDEN)
    ref = {k[len('case%d_' % ci):]: g[k] for k in g.files if k.startswith('case%d_' % ci)}
    return idx, hyps, scene, mod.gt_pose_for(scene, idx), (w / 2, h / 2), p, ref
