"""Drop-in replacement of the reference's `dsacstar` extension module.

The reference builds a pybind11 CPU extension exporting forward_rgb, backward_rgb, forward_rgbd and
backward_rgbd (/root/reference/dsacstar/dsacstar.cpp:887-892); its only call site is
/root/reference/utils/evaluation.py:162-172 (`dsacstar.forward_rgb`).  Here forward_rgb and backward_rgb
 run the hand-written sm_100a solver in libcrossloc_b200.so.  The RGB-D entry points are never
called by any Python file of the reference and are outside this hot path.

`import torch` must come first, as with the reference (/root/reference/README.md:51).
"""
from crossloc_b200.dsac import backward_rgb, backward_rgb_batch, forward_rgb, forward_rgb_batch, set_seed  # noqa: F401


def forward_rgbd(*args, **kwargs):
    raise NotImplementedError('dsacstar.forward_rgbd is out of scope: no RGB-D data in CrossLoc (SURVEY.md section 2, row 11)')


def backward_rgbd(*args, **kwargs):
    raise NotImplementedError('dsacstar.backward_rgbd is out of scope: no RGB-D data in CrossLoc (SURVEY.md section 2, row 11)')
