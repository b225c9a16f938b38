"""ofb_b200 — B200-native (sm_100a) implementation of the Once-for-Both bi-mask DeiT search step.

The directory is named ``once-for-both_b200`` (not an importable identifier); ``import ofb_b200`` (the loader module
at the repository root) registers this package under the name ``ofb_b200``.
"""
__version__ = "0.1.0"
