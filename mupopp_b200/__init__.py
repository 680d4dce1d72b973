"""mupopp_b200 -- B200-native (sm_100a) implementation of the MuPoPP per-timestep hot path.

Host side in Python (torch for device buffers/streams only); numerics in hand-written CUDA
kernels behind the C-ABI of include/mupopp_b200.h (libmupopp_b200.so, loaded with ctypes).
No CPU fallback: without the shared library or a CUDA device every compute entry point raises.
"""
from .mesh import BoxMesh, Mesh, RectangleMesh, UnitCubeMesh, UnitSquareMesh  # noqa: F401

__version__ = "0.1.0"
