"""centroidal_mpc_b200 — B200-native SCP hot path of ahmadgazar/centroidal-MPC.

Host package (Python, mirrors the reference's ``src`` / ``config`` names) over a C-ABI
shared library of hand-written sm_100a CUDA kernels (``csrc/``, ``include/cmpc.h``).
There is no CPU or numpy fallback: entry points raise when the library or a GPU is missing.
"""
__version__ = "0.1.0"
