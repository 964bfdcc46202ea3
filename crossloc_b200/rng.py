"""Counter-based RNG shared by the CUDA sampler, the oracles and the tests.

The reference draws the four correspondences of every hypothesis try from a
per-OpenMP-thread ``std::mt19937`` seeded once per process
(/root/reference/dsacstar/thread_rand.cpp:13-71, dsacstar_util.h:168-173), which
makes its output depend on thread count and call order.
This build replaces it by Philox4x32-10 keyed by ``seed`` with the counter
``(try, hypothesis, image, lane)``, so every (seed, image, hypothesis, try) names
one fixed 4-tuple of cells, independently of scheduling.  This file is the
specification; ``csrc/dsac_common.cuh`` and ``oracle/dsac_oracle.c`` restate it.

Mapping to a cell coordinate is the multiply-high range reduction
``(u32 * n) >> 32`` (Lemire, without rejection); x and y are drawn independently
and with replacement exactly as in the reference.
"""
import numpy as np

_M0 = 0xD2511F53
_M1 = 0xCD9E8D57
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Philox4x32 with 10 rounds. ``ctr``: 4 uint32, ``key``: 2 uint32 -> 4 uint32."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in ctr)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def sample_cells(seed, image, hyp, tr, width, height):
    """The four (x, y) cells of try ``tr`` of hypothesis ``hyp`` of image ``image``.

    Two Philox blocks (lane 0, lane 1) give eight 32-bit words r0..r7; point j uses
    x = (r[2j] * width) >> 32, y = (r[2j+1] * height) >> 32.
    """
    key = (seed & _MASK, (seed >> 32) & _MASK)
    r = philox4x32_10((tr, hyp, image, 0), key) + philox4x32_10((tr, hyp, image, 1), key)
    return [((r[2 * j] * width) >> 32, (r[2 * j + 1] * height) >> 32) for j in range(4)]


def sample_cells_array(seed, image, hyps, tr, width, height):
    """[hyps, 4, 2] int32 array of try ``tr`` for all hypotheses (x, y order)."""
    out = np.empty((hyps, 4, 2), dtype=np.int32)
    for h in range(hyps):
        out[h] = sample_cells(seed, image, h, tr, width, height)
    return out
