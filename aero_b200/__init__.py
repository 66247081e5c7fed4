"""aero_b200 -- B200-native (sm_100a) LDE / blake2s row commitment / DEEP / FRI core of the
Miden-Winterfell prover that starkoracles/Aero drives, behind a C ABI (include/aero_b200.h).

The Python layer is a ctypes face for tests and benchmarks; the product is libaero_b200.so.
"""
from ._lib import (AERO_ERR_BUFFER, AERO_ERR_CUDA, AERO_ERR_INVALID, AERO_ERR_NOMEM, AERO_ERR_STATE,
                   AERO_ERR_UNSUPPORTED, AERO_FORM_CANONICAL, AERO_FORM_MONTGOMERY, AERO_OK, Divisor, ProofOptions,
                   load)
from .prover import (AeroError, AirProgramBuilder, Context, FriProver, Group, RandomCoin, Segment, host_blake2s, host_hash_elements,
                     make_divisor, miden_options)

__all__ = [n for n in dir() if not n.startswith("_")]
