"""tclight_b200 — B200-native (sm_100a) implementation of TC-Light's two hot paths.

Sub-modules import ``_lib`` (ctypes binding of libtclight.so) on first use; there is no
CPU or PyTorch fallback for the arithmetic.
"""
__all__ = ["_lib", "ops"]
__version__ = "0.1.0"
