"""ctypes binding of the C-ABI in include/vln_b200.h (csrc/libvln_b200.so).

There is no CPU fallback: if the shared library is missing, loading raises, and every op that
needs it fails loudly.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# VLN_LIB_VARIANT=stamps selects a debug build (-DVLN_CHAIN_STAMPS: the step-chain kernels record globaltimer stamps,
# tools/chain_stamps.py); the product library has none of it compiled in
VARIANT = os.environ.get("VLN_LIB_VARIANT", "")
LIB_PATH = os.path.join(CSRC, "libvln_b200%s.so" % (("_" + VARIANT) if VARIANT else ""))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "vln_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -shared -> csrc/libvln_b200.so"""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(p) for p in deps)):
        return LIB_PATH
    extra = ["-DVLN_CHAIN_STAMPS"] if VARIANT == "stamps" else []
    cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


_lib = None

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_u64 = C.c_uint64
_i64 = C.c_int64

_SIGS = {
    "vln_version": ([], _i),
    "vln_last_error": ([], C.c_char_p),
    "vln_ctx_create": ([C.POINTER(_p), _p, _i, _i], _i),
    "vln_ctx_destroy": ([_p], None),
    "vln_gather_pano": ([_p, _p, _p, _p, _p, _i, _p], _i),
    "vln_gather_cand": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p], _i),
    "vln_gather_action_feat": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p], _i),
    "vln_pano_attn": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _p, _u64, _i, _p], _i),
    "vln_pano_attn_ld": ([_p, _p, _p, _p, _p, _i, _p, _p, _i, _p, _i, _i, _i, _f, _p, _u64, _p, _i, _p], _i),
    "vln_feature_mask_bits": ([_p, _i64, _i, _f, _p, _u64, _u64, _p], _i),
    "vln_feature_mask_bits_ld": ([_p, _i64, _i64, _i64, _i, _f, _p, _u64, _u64, _p], _i),
    "vln_ctx_attn_fwd_ld": ([_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p], _i),
    "vln_ctx_attn_bwd_ld": ([_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p], _i),
    "vln_lstm_pointwise_drop_fwd": ([_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _u64, _p], _i),
    "vln_lstm_pointwise_drop_bwd": ([_p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _f, _p, _u64, _p], _i),
    "vln_envdrop_ctx_step_fwd": ([_p] * 6 + [_i] + [_p] * 4 + [_i, _i, _i, _f, _p, _u64, _i, _p], _i),
    "vln_envdrop_ctx_step_bwd": ([_p] * 5 + [_i] + [_p] * 8 + [_i, _i, _i, _f, _p, _u64, _p], _i),
    "vln_linear_state_fwd": ([_p, _p, _i, _i, _p, _i, _i, _p, _i, _p, _i, _p, _p, _f, _p, _u64, _u64, _p, _p], _i),
    "vln_linear_state_bwd": ([_p, _p, _i, _i, _p, _i, _i, _p, _i, _p, _p, _i, _p, _i, _p, _f, _p, _u64, _u64, _p, _p], _i),
    "vln_envdrop_state_fwd": ([_p, _i, _p, _i, _p, _p, _i, _i, _f, _p, _u64, _u64, _p], _i),
    "vln_envdrop_state_bwd": ([_p, _p, _i, _p, _p, _i, _i, _p, _i, _i, _f, _p, _u64, _u64, _p], _i),
    "vln_envdrop_act_fwd": ([_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _u64, _p], _i),
    "vln_envdrop_act_bwd": ([_p, _i, _p, _p, _i, _i, _i, _f, _p, _u64, _u64, _p], _i),
    "vln_policy_env_act_fwd": ([_p, _p, _i, _p, _u64] + [_p] * 5 + [_p] * 5 + [_p] * 7 + [_p] * 8 + [_p] * 5
                               + [_i, _i, _f, _u64, _i, _p], _i),
    "vln_cand_policy_env_act_fwd": ([_p] * 6 + [_f, _u64, _p, _i, _p, _u64] + [_p] * 5 + [_p] * 3 + [_p] * 7 + [_p] * 8
                                    + [_p] * 5 + [_i, _i, _f, _u64, _i, _p], _i),
    "vln_cand_logits_bwd_policy": ([_p] * 14 + [_i, _i, _f, _p, _u64, _u64, _p], _i),
    "vln_cand_logits_fwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _p, _u64, _p], _i),
    "vln_cand_logits_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _p, _u64, _p], _i),
    "vln_ctx_attn_fwd": ([_p, _p, _p, _p, _p, _i, _i, _i, _p], _i),
    "vln_ctx_attn_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p], _i),
    "vln_lstm_pointwise_fwd": ([_p, _p, _p, _p, _p, _i, _i, _p], _i),
    "vln_lstm_pointwise_bwd": ([_p, _p, _p, _p, _p, _p, _p, _i, _i, _p], _i),
    "vln_linear_bf16x3": ([_p, _p, _i, _i, _p, _i, _i, _p, _p, _i, _i, _i, _p], _i),
    "vln_linear_bf16x3_tall": ([_p, _p, _i, _i, _p, _i, _i, _p, _p, _i, _i, _p], _i),
    "vln_linear_bf16x3_pair": ([_p] * 8 + [_i] * 5 + [_p], _i),
    "vln_split_bf16": ([_p, _p, _p, _p, _p, _i, _i, _p], _i),
    "vln_lstm_seq_fwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p], _i),
    "vln_lstm_seq_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p], _i),
    "vln_lstm_set_variant": ([_i], _i),
    "vln_policy_fwd": ([_p, _p, _i, _p, _u64, _p, _p, _p, _p, _p, _i, _p], _i),
    "vln_policy_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _p], _i),
    "vln_dropout": ([_p, _p, _i64, _f, _p, _u64, _p], _i),
    "vln_embed_drop_fwd": ([_p, _p, _p, _i64, _i, _i, _f, _p, _u64, _p], _i),
    "vln_embed_drop_bwd": ([_p, _p, _p, _i64, _i, _i, _i, _f, _p, _u64, _p], _i),
    "vln_dropout_mask": ([_p, _i64, _f, _p, _u64, _p], _i),
    "vln_rng_advance": ([_p, _u64, _p], _i),
    "vln_env_step": ([_p] * 21 + [_i, _p], _i),
    "vln_a2c_fwd": ([_p] * 7 + [_f, _f] + [_p] * 4 + [_i, _i, _p], _i),
    "vln_a2c_bwd": ([_p] * 4 + [_f] + [_p] * 3 + [_i, _i, _p], _i),
    "vln_env_observe": ([_p] * 11 + [_i, _p], _i),
    "vln_wgrad_tf32": ([_p, _i, _p, _i, _i, _i, _i, _p, _i, _i, _p, _i64, _p], _i),
    "vln_chain_begin": ([_p, _i, _p], _i),
    "vln_chain_end": ([], _i),
    "vln_dgrad_tf32": ([_p, _i, _p, _i, _i, _i, _i, _p, _i, _p], _i),
    "vln_seq_outer_sum": ([_p, _i64, _i, _p, _i64, _i, _i, _i, _i, _i, _p, _i, _p], _i),
    "vln_eval_paths": ([_p, _p, _i, _p, _p, _i, _p, _p, _p, C.c_double, _p, _i, _p], _i),
    "vln_grad_sqnorm": ([_p, C.POINTER(_i64), _i, _p, _f, _p], _i),
    "vln_optim_step": ([_p, _p, _p, _p, C.POINTER(_i64), C.POINTER(_f), _i, _p, _f, _i, _f, _i, _p], _i),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback for its kernels.")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(L, name)            # AttributeError here = header/library mismatch
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"{what or 'vln_b200 call'} failed (rc={rc}): {lib().vln_last_error().decode()}")
