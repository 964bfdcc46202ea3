"""Debug aid: GPU backward pass vs the stored tier-1 fixtures, with the largest differences located."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_dsac_backward_gpu import _minimal_set_mask, _run  # noqa: E402
from tests.util import backward_case  # noqa: E402

for ci in range(4):
    idx, hyps, scene, gt, cxcy, p, ref = backward_case(ci)
    loss, grad, dbg = _run(scene, gt, hyps, cxcy, p, idx)
    keep = ref['probs'] >= 1e-3
    d = np.abs(ref['grad'] - grad)
    print(ci, 'loss', float(ref['loss']), loss, 'probs', np.abs(ref['probs'] - dbg['probs']).max(),
          'ref_rt', np.abs(ref['ref_rt'] - dbg['ref_rt'])[keep].max(), 'grad rel', d.max() / np.abs(ref['grad']).max())
    m = _minimal_set_mask(ref, grad.shape[1:])
    print('   rel diff outside the minimal sets %.3g, inside %.3g' % (d[:, ~m].max() / np.abs(ref['grad']).max(), d[:, m].max() / np.abs(ref['grad']).max()))
    order = np.argsort((d * ~m[None]).ravel())[::-1][:4]
    elig = {tuple(c): h for h in np.nonzero(keep)[0] for c in ref['cells'][h]}
    for o in order:
        c, y, x = np.unravel_index(o, d.shape)
        print('   cell (x=%d, y=%d) ch %d  ref %.6g  gpu %.6g  minimal-set of hyp %s' % (x, y, c, ref['grad'][c, y, x], grad[c, y, x], elig.get((x, y))))
