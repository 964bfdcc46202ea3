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


def _pf_rows(t):
    """NCHW fp32 -> fp32 padded-flat matrix [B*(H+2)*(W+2)][C] with zero borders."""
    return F.pad(t, (1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(-1, t.size(1)).contiguous()


@pytest.mark.parametrize('case', [(64, 2, 9, 14, 'merge'), (256, 2, 7, 10, 'merge'), (512, 1, 6, 9, 'plain'),
                                  (128, 2, 9, 13, 'phased'), (64, 1, 8, 8, 'nonorm')])
def test_gn_backward_stage_matches_autograd(case):
    """cl_gn_backward (both passes) vs autograd of out = relu(res + relu(GroupNorm(raw))) / relu(GroupNorm(raw))."""
    from crossloc_b200 import layout
    from crossloc_b200.cnn import _Geometry
    from crossloc_b200.train_plan import TrainPlan, _Src
    c, b, h, w, kind = case
    torch.manual_seed(c + h)
    norm = None if kind == 'nonorm' else torch.nn.GroupNorm(32, c).to(DEV)
    if norm is not None:
        with torch.no_grad():
            norm.weight.uniform_(0.5, 1.5)
            norm.bias.uniform_(-0.5, 0.5)
    raw = (torch.randn(b, c, h, w, device=DEV) * 2 + 0.3).requires_grad_(True)
    res = torch.rand(b, c, h, w, device=DEV) - 0.3
    y = F.relu(raw if norm is None else norm(raw))
    out = F.relu(res + y) if kind == 'merge' else y
    g1, g2 = torch.randn_like(out) * 3e-4, torch.randn_like(out) * 1e-4
    params = [raw] + ([] if norm is None else [norm.weight, norm.bias])
    ref = torch.autograd.grad(out, params, g1 * 0.5 + g2 * 2.0)

    geo = _Geometry(b, h, w)
    raw_pf = _pf_rows(raw.detach())
    groups = 32
    stats = None
    if norm is not None:
        xr = raw.detach().double().reshape(b, groups, -1)
        stats = torch.stack([xr.sum(-1), (xr * xr).sum(-1)], -1).contiguous()
    half, two = torch.tensor([0.5], device=DEV), torch.tensor([2.0], device=DEV)
    if kind == 'phased':   # gradient handed over as 4 parity phases at half resolution, 64 floats of row pitch more
        hh, wh = (h + 1) // 2, (w + 1) // 2
        ph = torch.zeros(4, b, c + 64, hh, wh, device=DEV)
        for a in range(2):
            for bb in range(2):
                sub = (g1 * 0.5 + g2 * 2.0)[:, :, a::2, bb::2]
                ph[a * 2 + bb, :, :c, :sub.size(2), :sub.size(3)] = sub
        buf = torch.cat([_pf_rows(ph[i]) for i in range(4)], 0)
        srcs = [_Src(buf, c + 64, None, None, phased=True)]
    else:
        srcs = [_Src(_pf_rows(g1), c, half, None), _Src(_pf_rows(g2), c, two, torch.ones(1, device=DEV))]
    mask = layout.to_pf(out.detach(), 1, 2) if kind == 'merge' else None
    plan = TrainPlan.__new__(TrainPlan)
    plan._pool = {}
    from crossloc_b200 import _lib
    d_raw, scale_out, ab, dbias, g_buf = plan._gn_backward(
        _lib.load(), torch.cuda.current_stream().cuda_stream, geo, c, {'raw': raw_pf, 'stats': stats}, norm, True, srcs, mask,
        kind == 'merge' or len(srcs) > 1)
    torch.cuda.synchronize()
    got = layout.from_pf(d_raw, b, h, w, 2) * scale_out[1]
    assert rel(got, ref[0]) < 2e-5
    assert abs(float(scale_out[0] * scale_out[1]) - 1.0) < 1e-6
    assert rel(dbias.float(), ref[0].sum((0, 2, 3))) < 1e-4 or float(ref[0].sum((0, 2, 3)).abs().max()) < 1e-7
    if norm is not None:
        assert rel(ab.sum(0)[:, 1].float(), ref[1]) < 1e-5
        assert rel(ab.sum(0)[:, 0].float(), ref[2]) < 1e-5
    if kind == 'merge':
        g_ref = (g1 * 0.5 + g2 * 2.0) * (out.detach() > 0)
        assert rel(layout.raw_to_nchw(g_buf, b, h, w), g_ref) < 1e-6


@pytest.mark.parametrize('fused', [True, False])
def test_training_step_matches_reference_autograd(fused, monkeypatch):
    """One TransPoseNet training step (coord MLE loss): loss and every gradient vs the stock-torch definition, through the
    fused plan (crossloc_b200.train_plan) and through the per-layer path (crossloc_b200.train).  The strict comparison runs
    the plan in its most accurate arithmetic (e4m3 forward corrections, fp16x3 gradients); the default arithmetic (block-scaled
    e2m1 corrections in the forward and in the data gradient) is compared at the end with the bounds of that scheme."""
    import networks.networks as nets
    from crossloc_b200 import train_plan
    monkeypatch.setattr(train_plan, 'FORWARD', 'fp16+fp8')
    monkeypatch.setattr(train_plan, 'BACKWARD', 'fp16x3')
    monkeypatch.setattr(train_plan, 'WGRAD', 'fp16x3')
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
    loss_nat, g_nat = grads(lambda t: net.forward_train(t, fused=fused), 'probe')
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
    loss_nat, g_nat = grads(lambda t: net.forward_train(t, fused=fused), 'mle')
    assert abs(float(loss_nat) - float(loss_ref)) < 1e-5 * abs(float(loss_ref))
    name, err = worst_error(g_nat, g_ref)
    assert err < 5e-2, (name, err)
    # and it is what forward() itself runs when autograd is on
    out = net(x)
    assert out.requires_grad
    if fused:
        # default arithmetic of the plan: fp16 + fp4 forward and data gradients, one-pass weight gradients
        monkeypatch.undo()
        object.__setattr__(net, '_train_plan', None)
        loss_ref, g_ref = grads(net.forward_reference, 'probe')
        loss_nat, g_nat = grads(lambda t: net.forward_train(t, fused=True), 'probe')
        assert net._train_plan.engine.precision == train_plan.FORWARD and net._train_plan.backward_mode == train_plan.BACKWARD
        assert abs(float(loss_nat) - float(loss_ref)) < 1e-3 * abs(float(loss_ref))
        name, err = worst_error(g_nat, g_ref)
        assert err < 1e-1, (name, err)
        assert max(float((g_nat[n] - g_ref[n]).norm() / g_ref[n].norm()) for n in tail if g_ref[n].dim() == 4) < 5e-3


def test_fused_plan_refuses_stale_backward():
    """Two forwards before a backward would silently differentiate through overwritten activations: it raises instead."""
    import networks.networks as nets
    torch.manual_seed(1)
    net = nets.TransPoseNet(torch.zeros(3), True, False, 0, 0, 3, 1).to(DEV).train()
    x = torch.rand(1, 3, 32, 48, device=DEV)
    first = net.forward_train(x, fused=True)
    second = net.forward_train(x, fused=True)
    second.sum().backward()          # the latest forward is fine
    with pytest.raises(RuntimeError, match='ONE forward'):
        first.sum().backward()


@pytest.mark.parametrize('case', [
    # tiny, class, height, width, extra blocks, forward arithmetic
    (False, 'TransPoseNet', 50, 70, 1, 'fp16+fp8'),   # full width (e4m3 forward terms, CTA-pair weight gradient), ragged size
    (False, 'TransPoseNet', 64, 96, 0, 'fp16x3'),
    (True, 'TransPoseNet', 41, 59, 1, 'fp16+fp8'),    # odd sizes at every level of the strided ladder
    (False, 'Network', 48, 64, 0, 'fp16+fp8'),        # vanilla DSAC* network: no GroupNorm, 1-channel input
    (False, 'TransPoseNet', 50, 70, 1, 'fp16+fp4'),   # block-scaled e2m1 corrections in the forward and the data gradient
])
def test_fused_plan_gradients_on_more_shapes(case):
    """Fused training plan vs stock autograd on a smooth objective: loss value, and every parameter gradient measured
    against the larger of its own norm and 1e-4 of the global gradient scale (ReLU flips bound the agreement)."""
    import networks.networks as nets
    from crossloc_b200 import train_plan
    tiny, cls, h, w, extra, fwd = case
    torch.manual_seed(h + w)
    if cls == 'Network':
        net = nets.Network(torch.tensor([0., 0., 5.]), tiny).to(DEV).train()
        x = torch.rand(2, 1, h, w, device=DEV)
    else:
        net = nets.TransPoseNet(torch.tensor([0., 0., 5.]), tiny, False, extra, extra, 3, 1).to(DEV).train()
        x = torch.rand(2, 3, h, w, device=DEV)
    with torch.no_grad():
        shape = net.forward_reference(x).shape
    probe = torch.randn(shape, device=DEV)

    def grads(forward):
        net.zero_grad()
        loss = (forward(x) * probe).sum()
        loss.backward()
        return loss.detach(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}

    loss_ref, g_ref = grads(net.forward_reference)
    fp4 = fwd == 'fp16+fp4'
    loss_nat, g_nat = grads(lambda t: train_plan.forward_train(net, t, backward='fp16+fp4' if fp4 else 'fp16x3', forward=fwd,
                                                              wgrad='fp16x1' if fp4 else 'fp16x3'))
    assert abs(float(loss_nat) - float(loss_ref)) < (1e-3 if fp4 else 2e-4) * max(1.0, abs(float(loss_ref)))
    scale = max(float(g.double().norm()) for g in g_ref.values())
    errs = {n: float((g_nat[n].double() - g_ref[n].double()).norm()) / max(float(g_ref[n].double().norm()), 1e-4 * scale)
            for n in g_ref}
    worst = max(errs, key=errs.get)
    assert errs[worst] < (1e-1 if fp4 else 5e-2), (worst, errs[worst])
    # the typical parameter agrees much better than the worst one; with e4m3 forward terms the forward differs by 3e-5
    # instead of 1e-5 from fp32 (e2m1 terms: 1.7e-4), more pre-activations change sign and the toy-sized maps (63 cells)
    # feel every flip
    assert sorted(errs.values())[len(errs) // 2] < (5e-2 if fp4 else (2e-2 if (fwd == 'fp16+fp8' and not tiny) else 5e-3))
    # TF32-grade gradient GEMMs (one fp16 pass): same gradients to 10-bit operand precision
    _, g_fast = grads(lambda t: train_plan.forward_train(net, t, backward='fp16x1', forward=fwd))
    errs = {n: float((g_fast[n].double() - g_ref[n].double()).norm()) / max(float(g_ref[n].double().norm()), 1e-4 * scale)
            for n in g_ref}
    assert max(errs.values()) < 1e-1 and sorted(errs.values())[len(errs) // 2] < (5e-2 if fp4 else 2e-2)


def test_default_plan_gradients_are_closer_to_fp32_than_stock_tf32():
    """Whole-network gradients on frames large enough that single ReLU flips no longer dominate (2 x 480 x 720, full-width
    TransPoseNet, the reference's coord MLE loss): the default arithmetic of the fused plan (fp16 + fp4 forward and data
    gradients, one-pass weight gradients) against stock autograd in fp32 (TF32 off) -- within 1e-3 relative L2 over all
    parameters, and closer than stock autograd with TF32 on, i.e. than what `train_single_task.py` runs by default."""
    import networks.networks as nets
    from crossloc_b200 import synth, train_plan
    from loss.coord import scene_coords_regression_loss
    from tests.test_loss_cpu import pixel_grid
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.tensor(synth.NATURESCAPE_MEAN, dtype=torch.float32), False, False, 1, 1, 3, 1).to(DEV).train()
    batch, h, w = 2, 480, 720
    _, gt, poses, _ = synth.make_batch(0, batch, height=h, width=w)
    images = torch.rand(batch, 3, h, w, device=DEV)
    gt = torch.from_numpy(gt).to(DEV)
    poses = torch.from_numpy(poses).float().to(DEV)
    cam = torch.eye(3, device=DEV)
    cam[0, 0] = cam[1, 1] = 480.0
    cam[0, 2], cam[1, 2] = w / 2, h / 2
    grid = pixel_grid().to(DEV)

    def grads(forward):
        net.zero_grad()
        c, u = torch.split(forward(images), [3, 1], dim=1)
        loss, _ = scene_coords_regression_loss(0.1, 100.0, 1000.0, 50.0, 'MLE', grid, -1, cam, c, u, poses, gt)
        loss.backward()
        return {n: p.grad.detach().double().clone() for n, p in net.named_parameters()}

    def err(g, ref):
        return (sum(float((g[n] - ref[n]).norm()) ** 2 for n in ref) ** 0.5) / (sum(float(ref[n].norm()) ** 2 for n in ref) ** 0.5)

    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        ref = grads(net.forward_reference)
        torch.backends.cudnn.allow_tf32 = True
        stock_tf32 = err(grads(net.forward_reference), ref)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    object.__setattr__(net, '_train_plan', None)
    native = err(grads(lambda t: train_plan.forward_train(net, t)), ref)
    assert native < 1e-3, native
    assert native < stock_tf32, (native, stock_tf32)
