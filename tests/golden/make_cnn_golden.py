"""Generates tests/golden/cnn_golden.npz by importing the REFERENCE network (/root/reference/networks/networks.py).

Runs only in the authoring container (the reference tree does not travel to the GPU box).  The weights are
not stored (107 MB): they are the default initialisation under a fixed torch seed, which the twin module in
networks/networks.py reproduces bit for bit because it creates the same layers in the same order; a few
parameter checksums are stored to detect any drift of that assumption.

    python tests/golden/make_cnn_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (class, ctor args, ctor kwargs, weight seed, input shape, input seed)
    'transpose_default': ('TransPoseNet', (torch.zeros(3), False, False, 2, 2, 3, 1), {}, 2021, (1, 3, 64, 96), 0),
    'transpose_ragged': ('TransPoseNet', (torch.tensor([1., -2., 3.]), False, False, 2, 2, 3, 1), {}, 11, (2, 3, 52, 76), 1),
    'transpose_tiny_gray': ('TransPoseNet', (torch.tensor([1., 2., 3.]), True, True, 1, 0, 3, 0), {}, 5, (2, 1, 40, 56), 2),
    'network_vanilla': ('Network', (torch.tensor([1., 2., 3.]), False), {}, 7, (1, 1, 48, 64), 3),
    'network_tiny': ('Network', (torch.tensor([0., 0., 0.]), True), {}, 8, (1, 1, 48, 64), 4),
    # full-size DUC head (ragged size: the bilinear resize is not the identity) and MLR
    'transpose_fullsize_ragged': ('TransPoseNet', (torch.tensor([1., -2., 3.]), False, False, 1, 1, 3, 1),
                                  {'full_size_output': True}, 13, (2, 3, 52, 76), 5),
    'transpose_fullsize_even': ('TransPoseNet', (torch.zeros(3), False, False, 0, 1, 3, 1),
                                {'full_size_output': True}, 15, (1, 3, 64, 96), 7),
    'transpose_mlr3_tiny': ('TransPoseNet', (torch.tensor([0.5, 0., -1.]), True, False, 1, 1, 3, 1),
                            {'num_mlr': 3}, 14, (1, 3, 48, 64), 6),
}


def load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self      # the reference constructors call .cuda()
    sys.path.insert(0, '/root/reference')
    spec = importlib.util.spec_from_file_location('ref_networks', '/root/reference/networks/networks.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    out = {}
    for name, (cls, args, kwargs, wseed, shape, iseed) in CASES.items():
        torch.manual_seed(wseed)
        net = getattr(ref, cls)(*args, **kwargs).eval()
        g = torch.Generator().manual_seed(iseed)
        x = torch.rand(*shape, generator=g)
        with torch.no_grad():
            y = net(x)
        out[name + '_out'] = y.numpy()
        sd = net.state_dict()
        out[name + '_wsum'] = np.array([float(v.double().abs().sum()) for v in sd.values()])
        out[name + '_nkeys'] = np.int32(len(sd))
        print(name, tuple(y.shape), len(sd))
    np.savez_compressed(os.path.join(HERE, 'cnn_golden.npz'), **out)


if __name__ == '__main__':
    main()
