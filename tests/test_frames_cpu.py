"""Input-frame host logic against Pillow: the PNG decoder, and the resize coefficient tables with
a numpy emulation of the two device passes -- bit for bit what PIL.Image.resize(BILINEAR) (= torchvision Resize on a PIL
image, /root/reference/dataloader/dataloader.py:199-202) returns."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from crossloc_b200 import frames


def _png(arr, mode, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr, mode).save(buf, format='PNG', **kw)
    return buf.getvalue()


def _reference_rgb(data):
    """dataloader.py:306-316 with PIL as the decoder (io.imread's PNG plugin): gray2rgb / drop alpha."""
    img = np.asarray(Image.open(io.BytesIO(data)))
    if img.ndim < 3:
        img = np.stack([img] * 3, -1)
    if img.shape[-1] == 4:
        img = img[:, :, :3]
    if img.shape[-1] == 2:
        img = np.stack([img[:, :, 0]] * 3, -1)
    return img


@pytest.mark.parametrize('mode,shape', [('RGB', (37, 53, 3)), ('RGBA', (20, 31, 4)), ('L', (25, 40)), ('LA', (9, 14, 2))])
@pytest.mark.parametrize('level', [0, 6, 9])
def test_png_decoder_matches_pillow(mode, shape, level):
    rng = np.random.default_rng(len(mode) + level)
    smooth = (np.add.outer(np.arange(shape[0]), np.arange(shape[1])) * 3 % 256).astype(np.uint8)   # exercises Sub/Up/Avg/Paeth
    arr = rng.integers(0, 256, size=shape, dtype=np.uint8)
    if arr.ndim == 3:
        arr[..., 0] = smooth
    else:
        arr = smooth
    data = _png(arr, mode, compress_level=level)
    got = frames.decode_png_batch([data], pin=False)[0].numpy()
    assert got.shape == (shape[0], shape[1], 3)
    assert np.array_equal(got, _reference_rgb(data))


def test_png_palette_and_batch_and_errors():
    rng = np.random.default_rng(5)
    arr = rng.integers(0, 256, size=(16, 24, 3), dtype=np.uint8)
    pal = Image.fromarray(arr, 'RGB').convert('P', palette=Image.ADAPTIVE, colors=64)
    buf = io.BytesIO()
    pal.save(buf, format='PNG')
    got = frames.decode_png_batch([buf.getvalue()], pin=False)[0].numpy()
    assert np.array_equal(got, np.asarray(pal.convert('RGB')))
    batch = [_png(rng.integers(0, 256, size=(12, 18, 3), dtype=np.uint8), 'RGB') for _ in range(7)]
    out = frames.decode_png_batch(batch, threads=3, pin=False)
    for i, d in enumerate(batch):
        assert np.array_equal(out[i].numpy(), _reference_rgb(d))
    with pytest.raises(RuntimeError, match='not a PNG'):
        frames.decode_png_batch([b'definitely not a png file, just some bytes to get past the length check'], pin=False)
    with pytest.raises(RuntimeError, match='size differs'):
        frames.decode_png_batch([batch[0], _png(arr, 'RGB')], pin=False)
    sixteen = io.BytesIO()
    Image.fromarray((rng.integers(0, 65535, size=(8, 8))).astype(np.uint16)).save(sixteen, format='PNG')
    with pytest.raises(RuntimeError, match='8-bit'):
        frames.decode_png_batch([sixteen.getvalue()], pin=False)


def _emulate(img, ho, wo):
    """The two passes of csrc/frames.cu in numpy (int32 arithmetic, uint8 intermediate)."""
    def one_axis(x, out_size, axis):
        x = np.moveaxis(x, axis, 0).astype(np.int64)
        bounds, kk = frames.resize_coeffs(x.shape[0], out_size)
        out = np.empty((out_size,) + x.shape[1:], dtype=np.uint8)
        for i in range(out_size):
            lo, n = bounds[i]
            acc = (1 << 21) + np.tensordot(kk[i, :n].astype(np.int64), x[lo:lo + n], axes=(0, 0))
            out[i] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)
    h, w = img.shape[:2]
    mid = one_axis(img, wo, 1) if wo != w else img
    return one_axis(mid, ho, 0) if ho != h else mid


@pytest.mark.parametrize('hw,size', [((600, 900), 480), ((960, 1440), 480), ((480, 720), 480), ((300, 500), 480), ((777, 555), 480),
                                     ((1080, 1920), 480)])
def test_resize_tables_reproduce_pillow(hw, size):
    rng = np.random.default_rng(hw[0])
    img = rng.integers(0, 256, size=hw + (3,), dtype=np.uint8)
    ho, wo = frames.resized_shape(hw[0], hw[1], size)
    import torchvision.transforms as T
    want = np.asarray(T.Resize(size)(Image.fromarray(img, 'RGB')))
    assert want.shape == (ho, wo, 3)
    assert np.array_equal(_emulate(img, ho, wo), want)
