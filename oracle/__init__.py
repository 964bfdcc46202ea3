"""CPU oracles for the localization hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (crossloc_b200/, dsacstar/, networks/, loss/) never does.
"""
