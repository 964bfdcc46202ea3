"""crossloc_b200 -- B200-native implementation of CrossLoc's per-image localization hot path.

  crossloc_b200.dsac     DSAC* pose solver (hand-written sm_100a kernels behind a C ABI)
  crossloc_b200.cnn      scene-coordinate CNN engine (TMA + tcgen05 implicit-GEMM convolutions)
  crossloc_b200.synth    synthetic scene generator used by the tests and the bench

The reference-facing drop-in modules live at the repository root under the reference's own names:
`dsacstar`, `networks.networks`, `loss.coord`.
"""
__version__ = '0.1.0'
