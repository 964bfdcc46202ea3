"""CPU suite for the network twin: state-dict surface and forward_reference vs fixtures made from the reference."""
import os

import numpy as np
import pytest
import torch

import networks.networks as nets
from tests.util import ROOT

GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'cnn_golden.npz'))

CASES = {
    'transpose_default': ('TransPoseNet', (torch.zeros(3), False, False, 2, 2, 3, 1), 2021, (1, 3, 64, 96), 0),
    'transpose_ragged': ('TransPoseNet', (torch.tensor([1., -2., 3.]), False, False, 2, 2, 3, 1), 11, (2, 3, 52, 76), 1),
    'transpose_tiny_gray': ('TransPoseNet', (torch.tensor([1., 2., 3.]), True, True, 1, 0, 3, 0), 5, (2, 1, 40, 56), 2),
    'network_vanilla': ('Network', (torch.tensor([1., 2., 3.]), False), 7, (1, 1, 48, 64), 3),
    'network_tiny': ('Network', (torch.tensor([0., 0., 0.]), True), 8, (1, 1, 48, 64), 4),
    'transpose_fullsize_ragged': ('TransPoseNet', (torch.tensor([1., -2., 3.]), False, False, 1, 1, 3, 1), 13, (2, 3, 52, 76), 5,
                                  {'full_size_output': True}),
    'transpose_fullsize_even': ('TransPoseNet', (torch.zeros(3), False, False, 0, 1, 3, 1), 15, (1, 3, 64, 96), 7,
                                {'full_size_output': True}),
    'transpose_mlr3_tiny': ('TransPoseNet', (torch.tensor([0.5, 0., -1.]), True, False, 1, 1, 3, 1), 14, (1, 3, 48, 64), 6,
                            {'num_mlr': 3}),
}


def build_case(name, device='cpu'):
    cls, args, wseed, shape, iseed = CASES[name][:5]
    kwargs = CASES[name][5] if len(CASES[name]) > 5 else {}
    torch.manual_seed(wseed)
    net = getattr(nets, cls)(*args, **kwargs).eval().to(device)
    g = torch.Generator().manual_seed(iseed)
    x = torch.rand(*shape, generator=g).to(device)
    return net, x


@pytest.mark.parametrize('name', list(CASES))
def test_twin_reproduces_reference_fixture(name):
    """Same seeded weights (checksummed) and the same output as the reference module, bit for bit on CPU."""
    net, x = build_case(name)
    sd = net.state_dict()
    assert len(sd) == int(GOLD[name + '_nkeys'])
    wsum = np.array([float(v.double().abs().sum()) for v in sd.values()])
    assert np.allclose(wsum, GOLD[name + '_wsum'], rtol=1e-12, atol=0)
    with torch.no_grad():
        y = net.forward_reference(x)
    assert np.abs(y.numpy() - GOLD[name + '_out']).max() <= 1e-6 * np.abs(GOLD[name + '_out']).max()


def test_state_dict_surface():
    """Key names and shapes the reference checkpoints carry."""
    net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1)
    sd = net.state_dict()
    assert len(sd) == 116
    assert tuple(sd['encoder.conv1.weight'].shape) == (32, 3, 3, 3)
    assert tuple(sd['encoder.enc_add_res_block1.0.weight'].shape) == (512, 512, 3, 3)
    assert tuple(sd['encoder.enc_add_res_block1.1.weight'].shape) == (512,)
    assert tuple(sd['decoder.fc3.weight'].shape) == (4, 512, 1, 1)
    assert tuple(sd['mean'].shape) == (3,) and tuple(sd['decoder.mean'].shape) == (3,)
    assert net.OUTPUT_SUBSAMPLE == 8 and net.num_task_channel == 3 and net.num_pos_channel == 1
    assert sum(p.numel() for p in net.parameters()) == 26836996
    assert len(nets.Network(torch.zeros(3), False).state_dict()) == 35
    assert nets.Network.OUTPUT_SUBSAMPLE == 8
    for attr in ('encoder', 'decoder', 'mlr_encoder_ls', 'encoder_ls', 'mlr_ls', 'decoder_ls'):
        assert hasattr(net, attr)


def test_native_path_refuses_cpu_tensors():
    net = nets.TransPoseNet(torch.zeros(3), True, False).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match='no CPU fallback'):
        net(torch.rand(1, 3, 32, 32))
