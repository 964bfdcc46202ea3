"""Input frames without DataLoader workers: PNG decode on host threads, the dataset's Resize on the device.

Host-side mirror of the image path of the reference's data loader (/root/reference/dataloader/dataloader.py):
`_fetch_datapoint` (:306-323: io.imread, gray2rgb, RGBA -> RGB, focal length scaled by image_height / H) and
`image_transform` (:189-212: ToPILImage -> Resize(image_height) -> ToTensor [-> Normalize]).  The arithmetic is in
csrc/frames.cu (decode, resize) and csrc/cnn_pointwise.cu (ToTensor / Normalize inside cl_net_forward_frames).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


def png_size(data):
    """(height, width, channels stored in the file) of a PNG file image given as bytes."""
    lib = _lib.load()
    h, w, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
    _lib.check(lib.cl_png_info(ctypes.addressof(buf), len(data), ctypes.byref(h), ctypes.byref(w), ctypes.byref(c)))
    return h.value, w.value, c.value


def decode_png_batch(files, threads=None, pin=None):
    """PNG files (paths or bytes objects, all of one size) -> uint8 tensor [B, H, W, 3] (pinned when CUDA is available):
    what `io.imread` + gray2rgb / `[:, :, :3]` of dataloader.py:306-316 return, for a whole batch on host threads."""
    lib = _lib.load()
    blobs = [f if isinstance(f, (bytes, bytearray)) else open(f, 'rb').read() for f in files]
    if not blobs:
        raise RuntimeError('decode_png_batch: no files')
    h, w, _ = png_size(blobs[0])
    pin = torch.cuda.is_available() if pin is None else pin
    out = torch.empty(len(blobs), h, w, 3, dtype=torch.uint8, pin_memory=pin)
    keep = [(ctypes.c_char * len(b)).from_buffer_copy(b) for b in blobs]
    ptrs = (ctypes.c_void_p * len(blobs))(*[ctypes.addressof(k) for k in keep])
    sizes = (ctypes.c_size_t * len(blobs))(*[len(b) for b in blobs])
    threads = threads or min(len(blobs), os.cpu_count() or 1)
    _lib.check(lib.cl_decode_png_batch(ptrs, sizes, len(blobs), out.data_ptr(), h, w, int(threads)))
    return out


def resized_shape(height, width, size):
    """torchvision.transforms.Resize(size) with an int: the shorter side becomes `size`, the other int(size * long / short)."""
    if height <= width:
        return size, int(size * width / height)
    return int(size * height / width), size


def resize_coeffs(in_size, out_size):
    """(bounds [out, 2], kk [out, ksize]) -- Pillow's fixed-point bilinear table of one axis (host only, for checks)."""
    lib = _lib.load()
    ks = ctypes.c_int32()
    _lib.check(lib.cl_resize_coeffs(in_size, out_size, ctypes.byref(ks), None, None))
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ks.value), dtype=np.int32)
    _lib.check(lib.cl_resize_coeffs(in_size, out_size, ctypes.byref(ks), bounds.ctypes.data_as(_lib._i32p),
                                    kk.ctypes.data_as(_lib._i32p)))
    return bounds, kk


def resize_frames(frames, size):
    """uint8 CUDA frames [B, H, W, 3] -> [B, H', W', 3]: transforms.Resize(size) of dataloader.py:201, bit for bit
    (Pillow's antialiased bilinear resample).  `size` is an int (shorter side) or (H', W')."""
    lib = _lib.load()
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.size(3) != 3 or not frames.is_cuda:
        raise RuntimeError('resize_frames expects a uint8 CUDA tensor [B, H, W, 3]')
    frames = frames.contiguous()
    b, h, w, _ = frames.shape
    ho, wo = resized_shape(h, w, size) if isinstance(size, int) else size
    out = torch.empty(b, ho, wo, 3, dtype=torch.uint8, device=frames.device)
    need = lib.cl_resize_workspace_bytes(b, h, w, ho, wo)
    ws = torch.empty(max(int(need), 16), dtype=torch.uint8, device=frames.device)
    _lib.check(lib.cl_resize_frames(frames.data_ptr(), b, h, w, out.data_ptr(), ho, wo, ws.data_ptr(), ws.numel(),
                                    torch.cuda.current_stream(frames.device).cuda_stream))
    return out


def load_frames(files, image_height=480, device=None, focal_lengths=None):
    """The image part of `CamLocDataset._fetch_datapoint` for a batch: decoded, resized uint8 frames on the device (feed them
    to `net.forward_frames(frames[, mean, std])`, which applies ToTensor / Normalize) and the focal lengths scaled by
    image_height / H (dataloader.py:320-322)."""
    host = decode_png_batch(files)
    device = device or torch.device('cuda', torch.cuda.current_device())
    frames = host.to(device, non_blocking=True)
    scale = image_height / host.size(1)
    if (host.size(1), host.size(2)) != resized_shape(host.size(1), host.size(2), image_height):
        frames = resize_frames(frames, image_height)
    focal = None if focal_lengths is None else [float(f) * scale for f in focal_lengths]
    return frames, focal
