"""GPU parity of the training-step convolutions (forward, dgrad, wgrad) against torch autograd in fp32."""
import pytest
import torch
import torch.nn.functional as F

from crossloc_b200 import train

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


SHAPES = [
    # cin, cout, k, stride, B, H, W
    (256, 256, 3, 1, 2, 9, 14),
    (512, 512, 1, 1, 2, 9, 14),
    (256, 512, 3, 1, 1, 12, 10),
    (64, 128, 3, 2, 2, 20, 28),
    (64, 128, 3, 2, 2, 21, 27),      # odd input size
    (32, 64, 3, 2, 2, 16, 24),       # Cin = 32: dgrad output padded to 64 channels
    (128, 256, 3, 2, 1, 30, 44),
]


@pytest.mark.parametrize('shape', SHAPES)
def test_conv_forward_dgrad_wgrad_match_autograd(shape):
    cin, cout, k, stride, b, h, w = shape
    torch.manual_seed(sum(shape))
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).to(DEV)
    x = torch.randn(b, cin, h, w, device=DEV).relu().requires_grad_(True)
    y_ref = conv(x)
    gy = torch.randn_like(y_ref) * 1e-4          # small gradients: exercises the power-of-two rescaling
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(y_ref, (x, conv.weight, conv.bias), gy)

    x2 = x.detach().clone().requires_grad_(True)
    y = train.NativeConv2d.apply(x2, conv.weight, conv.bias, stride)
    gx, gw, gb = torch.autograd.grad(y, (x2, conv.weight, conv.bias), gy)
    assert rel(y, y_ref) < 2e-5
    assert rel(gx, gx_ref) < 5e-5
    assert rel(gw, gw_ref) < 5e-5
    assert rel(gb, gb_ref) < 1e-5


def test_training_step_matches_reference_autograd():
    """One TransPoseNet training step (coord MLE loss): loss and every gradient vs the stock-torch definition."""
    import networks.networks as nets
    from loss.coord import scene_coords_regression_loss
    from tests.test_loss_cpu import pixel_grid
    torch.manual_seed(3)
    net = nets.TransPoseNet(torch.tensor([0., 0., 50.]), True, False, 1, 1, 3, 1).to(DEV).train()
    x = torch.rand(2, 3, 64, 96, device=DEV)
    gt = torch.randn(2, 3, 8, 12, device=DEV) * 5 + torch.tensor([0., 0., 50.], device=DEV)[None, :, None, None]
    pose = torch.eye(4, device=DEV).repeat(2, 1, 1)
    cam = torch.eye(3, device=DEV)
    cam[0, 0] = cam[1, 1] = 60.0
    cam[0, 2], cam[1, 2] = 48.0, 32.0

    probe = torch.randn(2, 4, 8, 12, device=DEV)

    def grads(forward, objective):
        net.zero_grad()
        out = forward(x)
        if objective == 'mle':
            coords, unc = torch.split(out, [3, 1], dim=1)
            loss, _ = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', pixel_grid().to(DEV), -1, cam,
                                                   coords, unc, pose, gt)
        else:   # smooth functional of the output: the only kinks left are the network's own ReLUs
            loss = (out * probe).sum()
        loss.backward()
        return loss.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}

    def worst_error(g_nat, g_ref):
        # parameters whose gradient is analytically zero (a bias in front of a per-channel GroupNorm) only carry
        # round-off noise: measure every error against the larger of the parameter's own and 1e-4 of the global scale
        scale = max(float(g.double().norm()) for g in g_ref.values())
        errs = {n: float((g_nat[n].double() - g_ref[n].double()).norm()) / max(float(g_ref[n].double().norm()), 1e-4 * scale)
                for n in g_ref}
        worst = max(errs, key=errs.get)
        return worst, errs[worst]

    loss_ref, g_ref = grads(net.forward_reference, 'probe')
    loss_nat, g_nat = grads(net.forward_train, 'probe')
    assert abs(float(loss_nat) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    # Every kernel agrees with autograd to ~1e-6 in isolation (test above), but this toy network has only 24k
    # activations per layer: a single ReLU whose pre-activation lies within the 1e-5 forward difference of zero
    # flips and moves every upstream gradient by 1 / sqrt(24576) = 0.6 % (measured: 1 flip, 1.2 %).
    name, err = worst_error(g_nat, g_ref)
    assert err < 5e-2, (name, err)
    tail = [n for n in g_ref if n.startswith('decoder.fc')]          # downstream of every flip candidate but two
    assert max(float((g_nat[n] - g_ref[n]).norm() / g_ref[n].norm()) for n in tail if g_ref[n].dim() == 4) < 1e-3

    # the reference's loss (train_single_task.py:279-283): same value; its gradient is piecewise (validity masks,
    # soft clamp) over only 192 cells here, so one cell changing side moves every gradient by ~0.5 %
    loss_ref, g_ref = grads(net.forward_reference, 'mle')
    loss_nat, g_nat = grads(net.forward_train, 'mle')
    assert abs(float(loss_nat) - float(loss_ref)) < 1e-5 * abs(float(loss_ref))
    name, err = worst_error(g_nat, g_ref)
    assert err < 5e-2, (name, err)
    # and it is what forward() itself runs when autograd is on
    out = net(x)
    assert out.requires_grad
