"""`loss` package of the drop-in: only `loss.coord` is on the localization hot path.

The reference's `loss/` is a directory WITHOUT `__init__.py` (a namespace portion) that also holds `depth.py`,
`normal.py` and `semantics.py`, imported unconditionally by /root/reference/train_single_task.py:12-15 and
finetune_decoder_single_task.py:12-15.  A regular package shadows namespace portions wherever they sit on
`sys.path`, so with this repository on PYTHONPATH `import loss.depth` would fail.  The package path is therefore
extended with every other `loss/` directory on `sys.path`: `loss.coord` resolves to this twin (first entry of
`__path__`), the other task losses to the reference checkout.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _entry in list(sys.path):
    _cand = os.path.join(os.path.abspath(_entry or os.getcwd()), 'loss')
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)
