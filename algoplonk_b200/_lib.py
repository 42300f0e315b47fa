"""ctypes binding of libb200plonk.so (include/b200plonk.h).

There is no CPU fallback: if the shared library is missing every call raises, and without a usable CUDA device
every proving / MSM / NTT entry point does (b2p_init fails).  The verification entry points (b2p_verify,
b2p_verify_batch, b2p_pairing_check, b2p_kzg_vk_load, b2p_g2_generate_unsafe), the persisted-key parsers
(b2p_gnark_*_parse, b2p_circuit_save) and the marshalling helpers are host
arithmetic by design -- plonk.Verify runs on the CPU in the reference too -- and need no device.  Build with `python -c "import
__graft_entry__ as g; g.build()"` or `make -C algoplonk_b200/csrc -j8`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200plonk.so")

B2P_BN254, B2P_BLS12_381 = 0, 1
ERR_ARG, ERR_CUDA, ERR_INTERNAL, ERR_VERIFY = -1, -2, -3, -4
BASIS_CANONICAL, BASIS_LAGRANGE = 0, 1
NTT_INVERSE, NTT_COSET = 1, 2
SOLVE_AUTO, SOLVE_HOST, SOLVE_DEVICE = 0, 1, 2
IPC_HANDLE_BYTES = 64
SHARD_HANDLES = 9
STAT_NAMES = ["total_ms", "msm_ms", "msm_accum_ms", "ntt_ms", "quotient_ms", "msm_calls", "msm_accum_adds",
              "h2d_bytes", "d2h_bytes", "launches"]
STAT_COUNT = 16

# every symbol include/b200plonk.h declares: (name, restype, argtypes)
_vp, _u64, _u32, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SYMBOLS = [
    ("b2p_init", _int, [_int]),
    ("b2p_last_error", C.c_char_p, []),
    ("b2p_version", C.c_char_p, []),
    ("b2p_launch_count", _u64, []),
    ("b2p_host_alloc", _int, [_u64, C.POINTER(_vp)]),
    ("b2p_host_free", None, [_vp]),
    ("b2p_srs_load", _int, [_int, _vp, _u64, _vp, _u64, C.POINTER(_vp)]),
    ("b2p_srs_load_compressed", _int, [_int, _vp, _u64, _u64, C.POINTER(_vp)]),
    ("b2p_srs_generate_unsafe", _int, [_int, _vp, _u64, C.POINTER(_vp)]),
    ("b2p_srs_generate_unsafe_range", _int, [_int, _vp, _u64, _u64, C.POINTER(_vp)]),
    ("b2p_srs_generate_unsafe_strided", _int, [_int, _vp, _u64, _u64, _u64, C.POINTER(_vp)]),
    ("b2p_srs_to_lagrange", _int, [_vp, _u64, _vp]),
    ("b2p_srs_get_points", _int, [_vp, _u64, _u64, _vp]),
    ("b2p_srs_size", _u64, [_vp]),
    ("b2p_srs_msm_params", _int, [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_u64)]),
    ("b2p_srs_free", None, [_vp]),
    ("b2p_msm_g1", _int, [_vp, _int, _vp, _u64, _vp]),
    ("b2p_msm_g1_dev", _int, [_vp, _int, _vp, _u64, _vp]),
    ("b2p_msm_g2", _int, [_int, _vp, _vp, _u64, _vp]),
    ("b2p_g1_sum", _int, [_int, _vp, _u64, _vp]),
    ("b2p_srs_stream", _vp, [_vp]),
    ("b2p_srs_set_commit_hook", _int, [_vp, _vp, _vp]),
    ("b2p_device_copy", _int, [_vp, _vp, _u64]),
    ("b2p_shard_group_create", _int, [_int, _u32, _u32, _u64, _vp, _u64, C.POINTER(_vp)]),
    ("b2p_shard_group_attach", _int, [_vp, _vp, _vp]),
    ("b2p_shard_group_export", _int, [_vp, _vp]),
    ("b2p_shard_group_connect", _int, [_vp, _vp]),
    ("b2p_shard_group_connect_local", _int, [C.POINTER(_vp), _u32]),
    ("b2p_shard_group_serve_proof", _int, [_vp, _u64]),
    ("b2p_shard_group_msm", _int, [_vp, _vp, _u64, _vp]),
    ("b2p_shard_group_serve_msm", _int, [_vp, _u64]),
    ("b2p_shard_group_free", None, [_vp]),
    ("b2p_ntt", _int, [_int, _vp, _u64, _int]),
    ("b2p_ntt_shard_create", _int, [_int, _u64, _u32, _u32, C.POINTER(_vp)]),
    ("b2p_ntt_shard_free", None, [_vp]),
    ("b2p_ntt_shard_local_size", _u64, [_vp]),
    ("b2p_ntt_shard_chunk_size", _u64, [_vp]),
    ("b2p_ntt_shard_forward_local", _int, [_vp, _vp, _u64, _int, _vp, _vp]),
    ("b2p_ntt_shard_forward_combine", _int, [_vp, C.POINTER(_vp), _vp, _vp]),
    ("b2p_ntt_shard_inverse_split", _int, [_vp, _vp, C.POINTER(_vp), _vp]),
    ("b2p_ntt_shard_inverse_local", _int, [_vp, _vp, _int, _vp, _vp]),
    ("b2p_peer_alloc", _int, [_u64, C.POINTER(_vp), _vp]),
    ("b2p_peer_open", _int, [_vp, C.POINTER(_vp)]),
    ("b2p_peer_close", _int, [_vp]),
    ("b2p_peer_free", _int, [_vp]),
    ("b2p_circuit_load", _int, [_vp, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _u64,
                                C.POINTER(_vp)]),
    ("b2p_circuit_vk_commitments", _int, [_vp, _vp]),
    ("b2p_circuit_free", None, [_vp]),
    ("b2p_proof_raw_size", _u64, [_int, _u32]),
    ("b2p_prove", _int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("b2p_prove_dev", _int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("b2p_circuit_stream", _vp, [_vp]),
    ("b2p_proof_marshal_size", _u64, [_int, _u32]),
    ("b2p_marshal_proof", _int, [_int, _u32, _vp, _vp, _vp]),
    ("b2p_marshal_public_inputs", _int, [_int, _vp, _u32, _vp]),
    ("b2p_verify", _int, [_int, _u64, _u32, _u32, C.POINTER(_u64), _vp, _vp, _vp, _vp, _u64, _vp, _u64]),
    ("b2p_verify_batch", _int, [_int, _u64, _u32, _u32, C.POINTER(_u64), _vp, _vp, _vp, _vp, _u64, _vp, _u64, _u64,
                          C.POINTER(_u64)]),
    ("b2p_verify_batch_dev", _int, [_int, _u64, _u32, _u32, C.POINTER(_u64), _vp, _vp, _vp, _vp, _u64, _vp, _u64, _u64,
                              C.POINTER(_u64)]),
    ("b2p_pairing_check", _int, [_int, _vp, _vp, _u64, C.POINTER(_int)]),
    ("b2p_kzg_vk_load", _int, [_int, _vp, _u64, _vp, _vp]),
    ("b2p_g2_generate_unsafe", _int, [_int, _vp, _vp]),
    ("b2p_solver_create", _int, [_int, _u64, _u32, _u64, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    ("b2p_solver_create_hinted", _int, [_int, _u64, _u32, _u64, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _u32, _vp, C.POINTER(_vp)]),
    ("b2p_solver_set_hint_fn", _int, [_vp, _vp, _vp]),
    ("b2p_solver_solve", _int, [_vp, _vp, _int, _vp, _vp, _vp]),
    ("b2p_solver_solve_dev", _int, [_vp, _vp, _int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    ("b2p_solver_info", _int, [_vp, C.POINTER(_u64)]),
    ("b2p_solver_free", None, [_vp]),
    ("b2p_gnark_file_parse", _int, [_vp, _u64, _vp]),
    ("b2p_gnark_vk_parse", _int, [_int, _vp, _u64, _vp]),
    ("b2p_gnark_pk_parse", _int, [_int, _vp, _u64, _vp]),
    ("b2p_circuit_save", _int, [C.c_char_p, _int, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _u64]),
    ("b2p_circuit_load_file", _int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    ("b2p_circuit_set_profiling", _int, [_vp, _int]),
    ("b2p_circuit_stats", _int, [_vp, C.POINTER(C.c_double)]),
]


MAX_COMMITMENTS = 8


class GnarkFile(C.Structure):
    """b2p_gnark_file"""
    _fields_ = [("curve", C.c_int32), ("ecc_id", C.c_uint32), ("ccs_off", _u64), ("ccs_len", _u64),
                ("pk_off", _u64), ("pk_len", _u64), ("vk_off", _u64), ("vk_len", _u64)]


class GnarkVk(C.Structure):
    """b2p_gnark_vk"""
    _fields_ = [("size", _u64), ("nb_public", _u64), ("k", _u32), ("has_lines", _u32), ("encoded_len", _u64),
                ("commitment_indexes", _u64 * MAX_COMMITMENTS),
                ("size_inv", C.c_uint8 * 32), ("generator", C.c_uint8 * 32), ("coset_shift", C.c_uint8 * 32),
                ("points", C.c_uint8 * ((8 + MAX_COMMITMENTS) * 96)),
                ("kzg_g1", C.c_uint8 * 96), ("kzg_g2", C.c_uint8 * (2 * 192))]


class GnarkPk(C.Structure):
    """b2p_gnark_pk"""
    _fields_ = [("vk", GnarkVk), ("kzg_off", _u64), ("kzg_count", _u64), ("lagrange_off", _u64),
                ("lagrange_count", _u64)]


class Hint(C.Structure):
    """b2p_hint"""
    _fields_ = [("id", _u32), ("n_in", _u32), ("n_out", _u32), ("in_vars", C.POINTER(_u32)), ("out_vars", C.POINTER(_u32))]


# b2p_hint_fn: int fn(void* ctx, uint32_t id, const void* inputs, uint32_t n_in, void* outputs, uint32_t n_out)
HINT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32)


# b2p_commit_fn: int fn(void* ctx, const void* d_scalars, uint64_t n, void* out_affine)
COMMIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)


class B200PlonkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200plonk error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Loads the shared library and binds every declared symbol (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200PlonkError(rc, load().b2p_last_error().decode())


_initialised = False


def init(device: int = -1) -> None:
    """b2p_init: raises B200PlonkError when no CUDA device is usable."""
    global _initialised
    check(load().b2p_init(device))
    _initialised = True
