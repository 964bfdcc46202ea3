"""Generates tests/golden/loss_golden.npz by running the REFERENCE loss (/root/reference/loss/coord.py) on CPU.

Authoring container only (the reference tree does not travel).  Inputs are regenerated from seeds by
tests/test_loss_cpu.py::make_inputs; only the reference's outputs (loss, valid rate, gradient checksum) are stored.

    python tests/golden/make_loss_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.test_loss_cpu import CASES, make_inputs  # noqa: E402


def load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    import types
    # the reference module imports utils.learning / utils.io, which pull in packages missing here: stub them
    utils = types.ModuleType('utils')
    learning = types.ModuleType('utils.learning')
    io = types.ModuleType('utils.io')

    def pick_valid_points(coord_input, nodata_value, boolean=False):   # utils/learning.py:49-71 (boolean path)
        return torch.sum(coord_input == nodata_value, dim=1) == 0
    learning.pick_valid_points = pick_valid_points
    io.safe_printout = lambda words: None
    sys.modules['utils'], sys.modules['utils.learning'], sys.modules['utils.io'] = utils, learning, io
    spec = importlib.util.spec_from_file_location('ref_loss_coord', '/root/reference/loss/coord.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for k in ('utils', 'utils.learning', 'utils.io'):
        del sys.modules[k]
    return mod


def main():
    ref = load_reference()
    out = {}
    for name, case in CASES.items():
        args, kwargs, coords = make_inputs(case)
        loss, rate = ref.scene_coords_regression_loss(*args, **kwargs)
        total = loss.sum()
        total.backward()
        out[name + '_loss'] = loss.detach().numpy()
        out[name + '_rate'] = np.float64(rate)
        out[name + '_grad_abs_sum'] = np.float64(coords.grad.abs().sum().item())
        out[name + '_grad_sample'] = coords.grad.reshape(-1)[::97].numpy().copy()
        print(name, loss.detach().numpy(), rate)
    np.savez_compressed(os.path.join(HERE, 'loss_golden.npz'), **out)


if __name__ == '__main__':
    main()
