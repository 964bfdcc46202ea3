"""Device side of the input-frame path: Resize on the GPU == torchvision Resize on a PIL image,
and PNG files -> network output through crossloc_b200.frames == the reference's loader transforms -> network."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from crossloc_b200 import frames

pytestmark = pytest.mark.gpu


def _png(arr):
    buf = io.BytesIO()
    Image.fromarray(arr, 'RGB').save(buf, format='PNG')
    return buf.getvalue()


@pytest.mark.parametrize('hw,size', [((600, 900), 480), ((960, 1440), 480), ((300, 500), 480), ((777, 555), 480), ((480, 720), 480),
                                     ((1080, 1920), 480)])
def test_resize_kernel_is_bit_identical_to_torchvision_on_pil(hw, size):
    import torchvision.transforms as T
    rng = np.random.default_rng(hw[1])
    imgs = rng.integers(0, 256, size=(3,) + hw + (3,), dtype=np.uint8)
    got = frames.resize_frames(torch.from_numpy(imgs).cuda(), size).cpu().numpy()
    for b in range(3):
        want = np.asarray(T.Resize(size)(Image.fromarray(imgs[b], 'RGB')))
        assert got[b].shape == want.shape
        assert np.array_equal(got[b], want)


def test_png_files_to_network_output_match_the_loader_transforms():
    """dataloader.py:199-211 (ToPILImage -> Resize(480) -> ToTensor -> Normalize) + network vs decode_png_batch ->
    resize_frames -> forward_frames: identical frames, so identical arithmetic from the first convolution on."""
    import torchvision.transforms as T
    import networks.networks as nets
    torch.manual_seed(11)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 0, 0, 3, 1).eval().cuda()
    rng = np.random.default_rng(3)
    raw = [rng.integers(0, 256, size=(120, 180, 3), dtype=np.uint8) for _ in range(2)]
    mean, std = [0.4245, 0.4375, 0.3836], [0.1823, 0.1701, 0.1854]
    tf = T.Compose([T.ToPILImage(), T.Resize(96), T.ToTensor(), T.Normalize(mean=mean, std=std)])
    ref_in = torch.stack([tf(r) for r in raw]).cuda()
    dev_frames, focal = frames.load_frames([_png(r) for r in raw], image_height=96, focal_lengths=[480.0, 500.0])
    assert tuple(dev_frames.shape) == (2, 96, 144, 3) and focal == [480.0 * 0.8, 500.0 * 0.8]
    with torch.no_grad():
        want = net(ref_in)
        got = net.forward_frames(dev_frames, mean, std)
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
