#!/usr/bin/env python
"""Measurements of the SURVEY section 8 "next" rows built in round 2 (one JSON line):
  f3  input frames: PNG decode on host threads (frames/s vs Pillow on one core), device Resize (GB/s of pixels moved)
  f4  dsacstar.backward_rgb: device ms per image at 60 x 90 cells vs the tier-1 (cv2, Python) oracle on the host
"""
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crossloc_b200 import dsac, frames, synth  # noqa: E402


def timed_cuda(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def frames_rows():
    from PIL import Image
    rs = np.random.default_rng(0)
    # natural-image-like content (smooth + noise) so that deflate has realistic work
    base = rs.integers(0, 255, size=(60, 80, 3)).astype(np.uint8)
    img = np.asarray(Image.fromarray(base).resize((960, 720), Image.BICUBIC)).copy()
    img = np.clip(img.astype(np.int16) + rs.integers(-6, 7, size=img.shape), 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format='PNG')
    blob = buf.getvalue()
    files = [blob] * 64
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    for _ in range(3):
        host = frames.decode_png_batch(files)
    dec = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(8):
        np.asarray(Image.open(io.BytesIO(blob)).convert('RGB'))
    pil = (time.perf_counter() - t0) / 8
    dev = host.cuda()
    ms = timed_cuda(lambda: frames.resize_frames(dev, 480))
    out = frames.resize_frames(dev, 480)
    moved = dev.numel() + out.numel()
    return {'png_bytes': len(blob), 'frame': '720x960x3 -> 480x640x3', 'decode_frames_per_s': len(files) / dec, 'decode_threads': threads,
            'pillow_frames_per_s_one_core': 1 / pil, 'resize_ms_per_64_frames': ms, 'resize_GBps': moved / (ms * 1e-3) / 1e9,
            'replaces': 'dataloader.py:306-346 (io.imread, transforms.Resize) in <= 6 DataLoader workers (utils/evaluation.py:74)'}


def backward_rows():
    from oracle import dsac_backward_py as tier1
    out = {}
    s = synth.make_scene(3)
    gt = np.asarray(s['pose'], dtype=np.float32)
    for hyps in (64, 256):
        c = torch.from_numpy(np.ascontiguousarray(s['coords'])).unsqueeze(0).cuda()
        g = torch.zeros_like(c)
        gtt = torch.from_numpy(gt).reshape(1, 4, 4)
        fn = lambda: dsac.backward_rgb_batch(c, g, gtt, hyps, 10., s['focal'], 360., 240., 1., 1., 100., 100., 100., 8, seed=1305, image_base=3)
        ms = timed_cuda(fn, n=10)
        out['hyps%d_device_ms_per_image' % hyps] = ms
        cb = c.repeat(12, 1, 1, 1).contiguous()
        gb = torch.zeros_like(cb)
        gtb = gtt.repeat(12, 1, 1)
        fnb = lambda: dsac.backward_rgb_batch(cb, gb, gtb, hyps, 10., s['focal'], 360., 240., 1., 1., 100., 100., 100., 8, seed=1305, image_base=3)
        out['hyps%d_device_ms_per_12_images' % hyps] = timed_cuda(fnb, n=5)
    t0 = time.perf_counter()
    r = tier1.backward_rgb(s['coords'], gt, 64, 10., s['focal'], 360., 240., 1., 1., 100., 100., 100., 8, seed=1305, image=3)
    out['hyps64_tier1_oracle_s_per_image'] = time.perf_counter() - t0
    out['hyps64_hypotheses_refined'] = int((r['probs'] >= 1e-3).sum())
    out['note'] = 'device time of cl_dsac_backward_rgb incl. sampling and scoring; the reference (OpenMP C++, not buildable here) is only represented by the Python/cv2 oracle'
    return out


if __name__ == '__main__':
    print(json.dumps({'frames': frames_rows(), 'backward_rgb': backward_rows()}))
