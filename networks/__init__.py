"""Regular package on purpose: it must win over the reference's `networks/` directory.

The reference keeps `networks/networks.py` in a directory without `__init__.py`.  Were this one a namespace
portion too, `networks.networks` would resolve to whichever portion comes first on `sys.path` -- and a script run
from the CrossLoc checkout has its own directory at `sys.path[0]`, so the stock cuDNN network would silently be
used.  A regular package is found before any namespace portion, whatever the order.
"""
