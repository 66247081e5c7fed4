"""ctypes binding of libaero_b200.so (include/aero_b200.h + include/aero_prover.h).

The library is the product; this module only declares prototypes.  There is no CPU fallback: if the
shared object is missing and cannot be built, importing fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p

from . import build as _build

AERO_OK = 0
AERO_ERR_INVALID = 1
AERO_ERR_CUDA = 2
AERO_ERR_NOMEM = 3
AERO_ERR_STATE = 4
AERO_ERR_UNSUPPORTED = 5
AERO_ERR_BUFFER = 6
AERO_FORM_MONTGOMERY = 0
AERO_FORM_CANONICAL = 1

STATUS_NAMES = {0: "AERO_OK", 1: "AERO_ERR_INVALID", 2: "AERO_ERR_CUDA", 3: "AERO_ERR_NOMEM", 4: "AERO_ERR_STATE",
                5: "AERO_ERR_UNSUPPORTED", 6: "AERO_ERR_BUFFER"}

p_u64 = POINTER(c_uint64)
pp_u64 = POINTER(p_u64)
p_u8 = POINTER(c_uint8)


class Divisor(ctypes.Structure):
    """aero_divisor: (x^a - b) / prod (x - exemptions[k])."""

    _fields_ = [("a", c_uint64), ("b", c_uint64), ("n_exemptions", c_uint32), ("exemptions", c_uint64 * 8)]


class ProofOptions(ctypes.Structure):
    """aero_proof_options; defaults = Miden's ProofOptions::with_96_bit_security."""

    _fields_ = [("num_queries", c_uint8), ("blowup_factor", c_uint8), ("grinding_factor", c_uint8),
                ("hash_fn", c_uint8), ("field_extension", c_uint8), ("fri_folding_factor", c_uint8),
                ("fri_max_remainder_size", c_uint16)]


class AirNode(ctypes.Structure):
    """aero_air_node: op (AERO_AIR_*), operands a, b."""

    _fields_ = [("op", c_uint32), ("a", c_uint32), ("b", c_uint32)]


class AirProgram(ctypes.Structure):
    """aero_air_program: transition constraints as an arithmetic program + single-value assertions."""

    _fields_ = [("nodes", POINTER(AirNode)), ("n_nodes", c_uint32), ("consts", POINTER(c_uint64)), ("n_consts", c_uint32),
                ("n_transition", c_uint32), ("transition_out", POINTER(c_uint32)), ("transition_adj", POINTER(c_uint64)),
                ("n_boundary", c_uint32), ("boundary_col", POINTER(c_uint32)), ("boundary_value", POINTER(c_uint64)),
                ("boundary_adj", POINTER(c_uint64)), ("boundary_div", POINTER(c_uint32)),
                ("n_periodic", c_uint32), ("periodic_len", POINTER(c_uint32)), ("periodic_values", POINTER(c_uint64))]


AERO_AIR_CUR, AERO_AIR_NEXT, AERO_AIR_CONST, AERO_AIR_ADD, AERO_AIR_SUB, AERO_AIR_MUL, AERO_AIR_PERIODIC = range(7)

AUX_BUILDER = ctypes.CFUNCTYPE(c_int, c_void_p, p_u64, c_uint32, pp_u64)
CONSTRAINT_EVALUATOR = ctypes.CFUNCTYPE(c_int, c_void_p, pp_u64, c_uint32, c_uint64, p_u64, c_uint32, pp_u64)


HOST_BARRIER = ctypes.CFUNCTYPE(c_int, c_void_p)


class ProveInputs(ctypes.Structure):
    _fields_ = [("options", ProofOptions), ("trace_len", c_uint64), ("main_width", c_uint32), ("aux_width", c_uint32),
                ("aux_rands", c_uint32), ("inputs_on_device", c_int), ("main_cols", pp_u64), ("aux_cols", pp_u64),
                ("ce_cols", pp_u64), ("divisors", POINTER(Divisor)), ("n_div", c_uint32),
                ("n_constraint_coeffs", c_uint32), ("ce_blowup", c_uint32), ("aux_builder", AUX_BUILDER),
                ("constraint_evaluator", CONSTRAINT_EVALUATOR), ("user", c_void_p), ("pub_inputs_bytes", p_u8),
                ("pub_inputs_len", c_size_t), ("trace_meta", p_u8), ("trace_meta_len", c_uint16),
                ("air_program", POINTER(AirProgram))]


# name -> (restype, argtypes).  Every symbol include/*.h declares is listed here; tests assert that
# the shared object exports all of them.
PROTOTYPES = {
    # context
    "aero_ctx_create": (c_int, [POINTER(c_int), c_int, POINTER(c_void_p)]),
    "aero_ctx_destroy": (None, [c_void_p]),
    "aero_last_error": (c_char_p, [c_void_p]),
    "aero_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "aero_ctx_set_form": (c_int, [c_void_p, c_int]),
    "aero_ctx_get_form": (c_int, [c_void_p]),
    "aero_ctx_set_option": (c_int, [c_void_p, c_char_p, ctypes.c_longlong]),
    "aero_ctx_set_error": (None, [c_void_p, c_char_p]),
    "aero_ctx_profile_enable": (c_int, [c_void_p, c_int]),
    "aero_ctx_profile_filter": (c_int, [c_void_p, c_char_p]),
    "aero_ctx_profile_read": (c_int, [c_void_p, c_char_p, POINTER(c_size_t)]),
    "aero_launch_count": (c_uint64, []),
    "aero_version": (c_char_p, []),
    "aero_ctx_window_create": (c_int, [c_void_p, c_size_t, p_u8]),
    "aero_ctx_window_attach": (c_int, [c_void_p, c_int, p_u8]),
    "aero_ctx_window_attach_local": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "aero_ctx_window_ranks": (c_int, [c_void_p]),
    "aero_ctx_set_host_barrier": (c_int, [c_void_p, HOST_BARRIER, c_void_p]),
    "aero_ctx_shard_begin": (c_int, [c_void_p, c_char_p]),
    "aero_ctx_shard_end": (c_int, [c_void_p, c_int]),
    "aero_window_barrier": (c_int, [c_void_p]),
    "aero_upload_start": (c_int, [c_void_p, pp_u64, c_uint32, c_uint64, c_int, c_int, POINTER(c_void_p)]),
    "aero_upload_wait": (c_int, [c_void_p, POINTER(c_void_p)]),
    "aero_upload_free": (None, [c_void_p]),
    # segments
    "aero_segment_commit": (c_int, [c_void_p, pp_u64, c_uint32, c_uint64, c_uint32, c_int, POINTER(c_void_p), p_u8]),
    "aero_segment_commit_device": (c_int, [c_void_p, c_void_p, c_size_t, c_uint32, c_uint64, c_uint32, c_int,
                                           POINTER(c_void_p), p_u8]),
    "aero_ctx_set_shard": (c_int, [c_void_p, c_int, c_int]),
    "aero_segment_destroy": (None, [c_void_p]),
    "aero_segment_info": (c_int, [c_void_p, POINTER(c_uint32), POINTER(c_uint64), POINTER(c_uint32)]),
    "aero_segment_download_lde": (c_int, [c_void_p, pp_u64]),
    "aero_segment_download_polys": (c_int, [c_void_p, pp_u64]),
    "aero_segment_download_leaves": (c_int, [c_void_p, p_u8]),
    "aero_segment_open": (c_int, [c_void_p, p_u64, c_uint32, p_u64, p_u8, POINTER(c_size_t)]),
    # constraints
    "aero_constraints_into_poly": (c_int, [c_void_p, pp_u64, POINTER(Divisor), c_uint32, c_uint64, c_uint64,
                                           POINTER(c_void_p)]),
    "aero_constraints_into_poly_device": (c_int, [c_void_p, c_void_p, c_size_t, POINTER(Divisor), c_uint32, c_uint64,
                                                  c_uint64, POINTER(c_void_p)]),
    "aero_constraints_evaluate_device": (c_int, [c_void_p, POINTER(c_void_p), c_uint32, POINTER(AirProgram), p_u64, c_uint32,
                                                 c_uint32, c_uint32, c_void_p, c_size_t]),
    "aero_constraints_evaluate_into_poly": (c_int, [c_void_p, POINTER(c_void_p), c_uint32, POINTER(AirProgram), p_u64,
                                                    c_uint32, c_uint32, POINTER(Divisor), c_uint32, POINTER(c_void_p)]),
    "aero_periodic_column_table": (c_int, [p_u64, c_uint64, c_uint64, c_uint32, p_u64]),
    "aero_segment_commit_polys": (c_int, [c_void_p, c_uint32, p_u8]),
    # OOD + DEEP
    "aero_ood_eval": (c_int, [c_void_p, POINTER(c_void_p), c_uint32, c_void_p, c_uint64, p_u64, p_u64]),
    "aero_deep_compose": (c_int, [c_void_p, POINTER(c_void_p), c_uint32, c_void_p, c_uint64, p_u64, p_u64, p_u64,
                                  POINTER(c_void_p)]),
    "aero_fri_download_evaluations": (c_int, [c_void_p, p_u64, POINTER(c_uint64)]),
    # FRI
    "aero_fri_from_evaluations": (c_int, [c_void_p, p_u64, c_uint64, POINTER(c_void_p)]),
    "aero_fri_commit_layer": (c_int, [c_void_p, p_u8]),
    "aero_fri_fold": (c_int, [c_void_p, c_uint64]),
    "aero_fri_build_layers": (c_int, [c_void_p, p_u8, c_uint32, p_u8, p_u64]),
    "aero_fri_build_layers_grind": (c_int, [c_void_p, p_u8, c_uint32, c_uint32, p_u8, p_u64, p_u64]),
    "aero_segments_roots": (c_int, [c_void_p, POINTER(c_void_p), c_uint32, p_u8]),
    "aero_fri_open": (c_int, [c_void_p, p_u64, c_uint32, p_u8, POINTER(c_size_t)]),
    "aero_open_queries": (c_int, [c_void_p, c_void_p, POINTER(c_void_p), c_uint32, p_u64, c_uint32, p_u8,
                                  POINTER(c_size_t), POINTER(p_u64), POINTER(p_u8), POINTER(c_size_t)]),
    "aero_fri_destroy": (None, [c_void_p]),
    # auxiliary-segment construction
    "aero_running_product_columns": (c_int, [c_void_p, pp_u64, p_u64, c_uint32, c_uint64, pp_u64]),
    "aero_running_product_columns_device": (c_int, [c_void_p, c_void_p, c_size_t, p_u64, c_uint32, c_uint64, c_void_p,
                                                    c_size_t]),
    "aero_batch_inverse": (c_int, [c_void_p, p_u64, c_uint64, p_u64]),
    # grinding
    "aero_pow_min_nonce": (c_int, [c_void_p, p_u8, c_uint32, p_u64]),
    # standalone
    "aero_commit_rows_device": (c_int, [c_void_p, c_void_p, c_size_t, c_uint32, c_uint64, p_u8]),
    "aero_test_field_ops": (c_int, [c_void_p, p_u64, p_u64, c_size_t, p_u64]),
    "aero_measure_alu_peak": (c_int, [c_void_p, POINTER(ctypes.c_double)]),
    "aero_device_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "aero_device_free": (c_int, [c_void_p, c_void_p]),
    "aero_device_upload": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "aero_device_download": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "aero_device_sync": (c_int, [c_void_p]),
    # host driver (aero_prover.h)
    "aero_prove": (c_int, [c_void_p, POINTER(ProveInputs), p_u8, POINTER(c_size_t)]),
    "aero_group_create": (c_int, [POINTER(c_int), c_int, c_size_t, POINTER(c_void_p)]),
    "aero_group_destroy": (None, [c_void_p]),
    "aero_group_size": (c_int, [c_void_p]),
    "aero_group_ctx": (c_void_p, [c_void_p, c_int]),
    "aero_group_last_error": (c_char_p, [c_void_p]),
    "aero_group_prove": (c_int, [c_void_p, POINTER(ProveInputs), c_int, p_u8, POINTER(c_size_t)]),
    "aero_host_blake2s": (None, [p_u8, c_size_t, p_u8]),
    "aero_host_hash_elements": (None, [p_u64, c_size_t, p_u8]),
    "aero_coin_new": (c_void_p, [p_u8, c_size_t]),
    "aero_coin_free": (None, [c_void_p]),
    "aero_coin_reseed": (None, [c_void_p, p_u8]),
    "aero_coin_reseed_with_int": (None, [c_void_p, c_uint64]),
    "aero_coin_draw": (c_int, [c_void_p, p_u64]),
    "aero_coin_draw_integers": (c_int, [c_void_p, c_uint32, c_uint64, p_u64]),
    "aero_coin_leading_zeros": (c_uint32, [c_void_p]),
    "aero_coin_check_leading_zeros": (c_uint32, [c_void_p, c_uint64]),
    "aero_coin_seed": (None, [c_void_p, p_u8]),
}

_LIB = None


def load() -> ctypes.CDLL:
    """Loads (building first if the sources are newer) libaero_b200.so; raises if unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if _build.is_stale():
        if os.environ.get("AERO_B200_NO_BUILD") and os.path.exists(path):
            pass
        else:
            _build.build_library()
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
