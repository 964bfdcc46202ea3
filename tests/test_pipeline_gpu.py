"""GPU parity of the end-to-end localizer (crossloc_b200.pipeline): network + solver on device-resident maps,
the body of the reference's evaluation loop (/root/reference/test_single_task.py:328-370), for the sub-sampled
network and for the full-size DUC variant (OUTPUT_SUBSAMPLE = 1)."""
import numpy as np
import pytest
import torch

from crossloc_b200 import synth
from oracle import dsac_oracle_c as tier2

pytestmark = pytest.mark.gpu


def _localize(net, height, width, batch, hyps, subsample, first, focal_px):
    from crossloc_b200.pipeline import Localizer
    dev = torch.device('cuda', 0)
    images = torch.rand(batch, 3, height, width, generator=torch.Generator().manual_seed(first)).to(dev)
    coords, _, poses, focal = synth.make_batch(first, batch, height=height, width=width, focal=focal_px,
                                               subsample=subsample)
    offsets = torch.from_numpy(coords).to(dev)
    loc = Localizer(net, hyps=hyps, device=dev)
    assert loc.subsample == subsample
    pose, dbg = loc.localize_device(images, torch.from_numpy(focal).to(dev), offsets, image_base=first, debug=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        native = net(images)
        ref = net.forward_reference(images)
    return images, offsets, focal, poses, pose.cpu().numpy(), dbg, native, ref


@pytest.mark.parametrize('full_size', [False, True])
def test_localizer_matches_oracle_on_the_same_map(full_size):
    import networks.networks as nets
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(2021)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 1, 1, 3, 1, full_size_output=full_size).eval().cuda()
    height, width, batch, hyps = (64, 96, 2, 32) if full_size else (96, 144, 2, 64)
    sub = 1 if full_size else 8
    images, offsets, focal, gt, pose, dbg, native, ref = _localize(net, height, width, batch, hyps, sub, 60, 120.0)
    assert tuple(native.shape) == ((batch, 4, height, width) if full_size else (batch, 4, height // 8, width // 8))
    rel = float((native[:, :3] - ref[:, :3]).norm() / ref[:, :3].norm())
    assert rel < 1e-3, rel   # north-star tolerance on the regressed map
    solver_in = (native[:, :3] + offsets).cpu().numpy()
    for b in range(batch):
        o = tier2.forward_rgb(solver_in[b], hyps, 10.0, float(focal[b]), width / 2, height / 2, 100.0, 100.0, sub,
                              seed=1305, image=60 + b)
        assert o['best'] == int(dbg['best'][b])
        assert (np.asarray(o['tries']) == dbg['tries'][b].numpy()).all()
        assert np.abs(o['pose'] - pose[b]).max() < 1e-3 * max(1.0, np.abs(o['pose']).max())
        t_err, r_err = synth.pose_errors(gt[b], pose[b])
        assert t_err < 5.0 and r_err < 5.0   # the random-init network output perturbs the scene by up to a few metres


def test_localizer_host_pipeline_equals_device_entry():
    """submit()/result() with pinned host frames (the e2e path of bench.py) returns what localize_device returns."""
    import networks.networks as nets
    from crossloc_b200.pipeline import Localizer
    torch.manual_seed(3)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 0, 0, 3, 1).eval().cuda()
    dev = torch.device('cuda', 0)
    height, width, batch = 64, 96, 3
    coords, _, _, focal = synth.make_batch(10, batch, height=height, width=width, focal=100.0)
    offsets, focal_d = torch.from_numpy(coords).to(dev), torch.from_numpy(focal).to(dev)
    loc = Localizer(net, hyps=32, device=dev)
    frames = [torch.rand(batch, 3, height, width, generator=torch.Generator().manual_seed(s)).pin_memory() for s in (1, 2, 3)]
    host = []
    loc.submit(frames[0], focal_d, offsets, image_base=0)
    for k in (1, 2):   # two batches in flight at most: a slot's host buffer is reused two submits later
        loc.submit(frames[k], focal_d, offsets, image_base=100 * k)
        host.append(loc.result().clone())
    host.append(loc.result().clone())
    for k, f in enumerate(frames):
        direct = loc.localize_device(f.to(dev), focal_d, offsets, image_base=100 * k)
        torch.cuda.synchronize()
        assert torch.equal(direct.cpu(), host[k])


def test_uint8_frames_match_torchvision_transforms():
    """cl_frames_to_nchw vs the host transform of the reference's dataloader (ToPILImage -> Resize(480) -> ToTensor
    [-> Normalize], dataloader/dataloader.py:189-212) on frames that already have the network height: bit for bit."""
    from torchvision import transforms
    from crossloc_b200.pipeline import frames_to_network_input
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (3, 48, 72, 3), dtype=torch.uint8, generator=g)
    mean, std = [0.4245, 0.4375, 0.3836], [0.1823, 0.1701, 0.1854]
    raw = transforms.Compose([transforms.ToPILImage(), transforms.Resize(48), transforms.ToTensor()])
    norm = transforms.Compose([transforms.ToPILImage(), transforms.Resize(48), transforms.ToTensor(),
                               transforms.Normalize(mean=mean, std=std)])
    ref_raw = torch.stack([raw(f.numpy()) for f in frames])
    ref_norm = torch.stack([norm(f.numpy()) for f in frames])
    assert torch.equal(frames_to_network_input(frames.cuda()).cpu(), ref_raw)
    assert torch.equal(frames_to_network_input(frames.cuda(), mean, std).cpu(), ref_norm)


def test_localizer_accepts_uint8_host_frames():
    """submit() with pinned uint8 HWC frames gives the poses of the fp32 frames ToTensor would have produced."""
    import networks.networks as nets
    from crossloc_b200.pipeline import Localizer
    torch.manual_seed(4)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 0, 0, 3, 1).eval().cuda()
    dev = torch.device('cuda', 0)
    height, width, batch = 64, 96, 2
    coords, _, _, focal = synth.make_batch(20, batch, height=height, width=width, focal=100.0)
    offsets, focal_d = torch.from_numpy(coords).to(dev), torch.from_numpy(focal).to(dev)
    frames = torch.randint(0, 256, (batch, height, width, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(6))
    loc = Localizer(net, hyps=32, device=dev)
    from_u8 = loc.localize(frames.pin_memory(), focal_d, offsets, image_base=7).clone()
    as_f32 = (frames.permute(0, 3, 1, 2).float() / 255).contiguous().pin_memory()
    from_f32 = loc.localize(as_f32, focal_d, offsets, image_base=7).clone()
    assert torch.equal(from_u8, from_f32)
