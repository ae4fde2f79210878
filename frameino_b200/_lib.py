"""ctypes loader for the C-ABI library ``libframeino_b200.so`` (declared in ``include/frameino_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``frameino_b200/csrc/Makefile``. There is no CPU
fallback: if the library is missing, or a compute entry point is called without a GPU, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# FINO_LIB_PATH: A/B runs of two builds of the library inside one GPU session (tools/kbench.py); never set in production
LIB_PATH = os.environ.get("FINO_LIB_PATH") or os.path.join(_HERE, "libframeino_b200.so")

_lib = None

_P = c_void_p
_I = c_int
_L = c_int64
_F = c_float

# name -> (restype, argtypes); mirrors include/frameino_b200.h one to one
SIGNATURES = {
    "fino_abi_version": (_I, []),
    "fino_last_error": (c_char_p, []),
    "fino_set_device": (_I, [_I]),
    "fino_launch_count": (_L, []),
    "fino_gemm_bf16": (_I, [_P, _L, _P, _L, _P, _P, _L, _L, _I, _I, _I, _I, _I, _P, _L, _P, _L, _P, _L, _P]),
    "fino_gemm_set_mode": (_I, [_I]),
    "fino_gemm_set_split": (_I, [_I]),
    "fino_gemm_plan": (_I, [_L, _I, _I, _I, _I, _P, _P]),
    "fino_attention_fwd": (_I, [_P, _P, _P, _P, _I, _I, _L, _L, _I, _L, _L, _L, _L, _L, _L, _L, _L, _F, _P]),
    "fino_attention_set_variant": (_I, [_I]),
    "fino_attention_set_split": (_I, [_I]),
    "fino_attention_plan": (_I, [_L, _L, _I, _I, _I, _I, _P, _P]),
    "fino_attention_plan_hd": (_I, [_L, _L, _I, _I, _I, _I, _I, _P, _P, _P]),
    "fino_rows_set_variant": (_I, [_I, _I]),
    "fino_rows_set_tma": (_I, [_I]),
    "fino_ln_modulate": (_I, [_P, _P, _L, _I, _L, _L, _F, _P, _P, _P, _P, _L, _P, _L, _I, _P]),
    "fino_gate_residual": (_I, [_P, _P, _P, _L, _I, _L, _L, _L, _P, _L, _P, _L, _I, _P]),
    "fino_qk_norm_rope": (_I, [_P, _L, _L, _P, _P, _I, _P, _L, _L, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _L, _L, _P]),
    "fino_patchify": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _P]),
    "fino_unpatchify": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _I, _P]),
    "fino_timestep_embedding": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _P]),
    "fino_linear_small_m": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fino_timestep_dedup": (_I, [_P, _L, _P, _P, _P, _P]),
    "fino_build_mod_table": (_I, [_P, _P, _P, _I, _I, _I, _L, _P]),
    "fino_wan_pack_model_input": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _L, _P]),
    "fino_wan_cfg_euler_step": (_I, [_P, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "fino_conv3d_cl_bf16": (_I, [_P, _I, _I, _I, _I, _L, _L, _L, _P, _L, _P, _P, _I, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I,
                                 _I, _I, _I, _P, _I, _P]),
    "fino_rms_act_cl": (_I, [_P, _P, _L, _I, _L, _L, _P, _P, _I, _P]),
    "fino_upsample2x_cl": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "fino_dupup_add_cl": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fino_avgdown_add_cl": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "fino_softmax_rows": (_I, [_P, _P, _L, _I, _L, _L, _F, _P]),
    "fino_vae_to_cl": (_I, [_P, _I, _P, _I, _I, _I, _I, _L, _L, _L, _L, _I, _I, _P]),
    "fino_vae_from_cl": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _P]),
    "fino_swap01": (_I, [_P, _P, _L, _L, _L, _P]),
    "fino_peer_alloc": (_I, [_L, _P]),
    "fino_peer_free": (_I, [_P]),
    "fino_peer_export": (_I, [_P, _P]),
    "fino_peer_import": (_I, [_P, _P]),
    "fino_peer_release": (_I, [_P]),
    "fino_peer_barrier": (_I, [_P, _I, _I, c_uint32, _P]),
    "fino_peer_status": (_I, [_P, _P, _P]),
    "fino_halo_exchange": (_I, [_P, _I, _I, _L, _L, _P, _P, _P, c_uint32, _L, _I, _P]),
    "fino_qkv_norm_rope_scatter": (_I, [_P, _L, _L, _P, _P, _I, _I, _F, _P, _P, _P, _I, _I, _L, _L, _P]),
    "fino_qkv_ln_rope_scatter": (_I, [_P, _L, _L, _P, _P, _P, _P, _I, _I, _F, _P, _P, _L, _P, _I, _I, _L, _L, _P]),
    "fino_attention_fwd_scatter": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _L, _L, _I, _L, _L, _L, _L, _L, _L, _L, _L, _F,
                                        _P]),
}


def load() -> ctypes.CDLL:
    """Loads (once) and returns the library, with argtypes/restypes set for every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(frameino_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    # tuning hooks for A/B runs of a whole step (tools/step_breakdown.py, bench.py): same effect as the setters in ops
    for env, fn in (("FINO_ATTN_VARIANT", lib.fino_attention_set_variant), ("FINO_ATTN_SPLIT", lib.fino_attention_set_split),
                    ("FINO_GEMM_MODE", lib.fino_gemm_set_mode), ("FINO_GEMM_SPLIT", lib.fino_gemm_set_split)):
        if os.environ.get(env):
            if fn(int(os.environ[env])) != 0:
                raise RuntimeError(f"{env}={os.environ[env]}: {lib.fino_last_error().decode()}")
    return lib


class FinoError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().fino_last_error()
        raise FinoError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")
