"""ctypes binding of oracle/dsac_oracle.c (tier-2 oracle, also the timed CPU baseline).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libdsac_oracle.so')
_lib = None
MAX_REF_STEPS = 100


def build(force=False):
    src = os.path.join(_HERE, 'dsac_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-B', '-C', _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.ora_forward_rgb.restype = ctypes.c_int
        _lib.ora_p3p.restype = ctypes.c_int
        _lib.ora_lm.restype = ctypes.c_int
        _lib.ora_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def num_threads():
    return lib().ora_num_threads()


def p3p(obj, img, f, cx, cy):
    obj = np.ascontiguousarray(obj, dtype=np.float64).reshape(12)
    img = np.ascontiguousarray(img, dtype=np.float64).reshape(8)
    r = np.zeros(3)
    t = np.zeros(3)
    ok = lib().ora_p3p(_p(obj, ctypes.c_double), _p(img, ctypes.c_double), ctypes.c_double(f), ctypes.c_double(cx),
                       ctypes.c_double(cy), _p(r, ctypes.c_double), _p(t, ctypes.c_double))
    return bool(ok), r, t


def lm(obj, img, f, cx, cy, rvec, tvec):
    obj = np.ascontiguousarray(obj, dtype=np.float32)
    img = np.ascontiguousarray(img, dtype=np.float32)
    r = np.array(rvec, dtype=np.float64).reshape(3).copy()
    t = np.array(tvec, dtype=np.float64).reshape(3).copy()
    ok = lib().ora_lm(ctypes.c_int(len(obj)), _p(obj, ctypes.c_float), _p(img, ctypes.c_float), ctypes.c_double(f),
                      ctypes.c_double(cx), ctypes.c_double(cy), _p(r, ctypes.c_double), _p(t, ctypes.c_double))
    return bool(ok), r, t


def forward_rgb(coords, hyps, thr, focal, cx, cy, alpha, max_reproj, sub_sampling, seed=1305, image=0,
                max_tries=1000000, forced_samples=None, refine=True):
    """One image, coords float32 [3, Hc, Wc].  Returns the same dict layout as the tier-1 oracle."""
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    if coords.ndim == 4:
        assert coords.shape[0] == 1
        coords = coords[0]
    hc, wc = coords.shape[1:]
    pose = np.zeros(16, dtype=np.float32)
    best = ctypes.c_int32(0)
    scores = np.zeros(hyps, dtype=np.float64)
    hyps_rt = np.zeros((hyps, 6), dtype=np.float64)
    tries = np.zeros(hyps, dtype=np.int32)
    counts = np.zeros(MAX_REF_STEPS, dtype=np.int32)
    rt = np.zeros(6, dtype=np.float64)
    forced = None
    if forced_samples is not None:
        forced = np.ascontiguousarray(forced_samples, dtype=np.int32).reshape(hyps, 4, 2)
    lib().ora_forward_rgb(
        _p(coords, ctypes.c_float), ctypes.c_int(hc), ctypes.c_int(wc), _p(pose, ctypes.c_float), ctypes.c_int(hyps),
        ctypes.c_float(thr), ctypes.c_float(focal), ctypes.c_float(cx), ctypes.c_float(cy), ctypes.c_float(alpha),
        ctypes.c_float(max_reproj), ctypes.c_int(sub_sampling), ctypes.c_uint64(seed), ctypes.c_uint32(image),
        ctypes.c_uint(max_tries), _p(forced, ctypes.c_int32), ctypes.c_int(1 if refine else 0), ctypes.byref(best),
        _p(scores, ctypes.c_double), _p(hyps_rt, ctypes.c_double), _p(tries, ctypes.c_int32),
        _p(counts, ctypes.c_int32), _p(rt, ctypes.c_double))
    return {'pose': pose.reshape(4, 4), 'best': int(best.value), 'scores': scores, 'hyps_rt': hyps_rt,
            'tries': tries, 'refine_counts': [int(c) for c in counts if c >= 0], 'rvec': rt[:3].copy(),
            'tvec': rt[3:].copy()}
