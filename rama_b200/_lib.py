"""ctypes loader for librama_b200.so (the C ABI declared in include/rama_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
calls raise.  The library is built in-tree by ``make -C rama_b200/csrc`` (see __graft_entry__.build).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "librama_b200.so")

T_COUNT = 14
K_COUNT = 9
PK_COUNT = 5
PREFILL_KINDS = ["gemm", "attn", "norm", "comm", "other"]
KERNEL_KINDS = ["embed", "qkv", "attn", "wo", "w13", "w2", "cls", "sample", "comm"]
STATE = ["x", "xb", "xb2", "hb", "hb2", "q", "k", "v", "att", "logits", "key_cache", "value_cache"]


class RamaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rama_b200 error {code}: {msg}")
        self.code = code


class CConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads",
                                         "vocab_size", "seq_len", "shared_weight")]


class CTp(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("nccl_id", C.c_uint8 * 128)]


_lib = None

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)
vp = C.c_void_p
sz = C.c_size_t

_SIGS = {
    "rama_abi_version": ([], C.c_int),
    "rama_last_error": ([], C.c_char_p),
    "rama_device_count": ([C.POINTER(C.c_int)], C.c_int),
    "rama_ctx_create": ([C.c_int, C.POINTER(CTp), C.POINTER(vp)], C.c_int),
    "rama_ctx_create_multi": ([C.c_int32, ip, C.POINTER(vp)], C.c_int),
    "rama_ctx_destroy": ([vp], C.c_int),
    "rama_tp_unique_id": ([C.POINTER(C.c_uint8)], C.c_int),
    "rama_ctx_load_file": ([vp, C.c_char_p], C.c_int),
    "rama_last_load_gbps": ([C.POINTER(C.c_double)], C.c_int),
    "rama_ctx_load_host": ([vp, C.POINTER(CConfig), C.POINTER(fp)], C.c_int),
    "rama_ctx_load_synthetic": ([vp, C.POINTER(CConfig), C.c_uint64, fp, fp, fp, fp], C.c_int),
    "rama_ctx_config": ([vp, C.POINTER(CConfig)], C.c_int),
    "rama_ctx_weight_to_host": ([vp, C.c_int, fp, sz, C.POINTER(sz)], C.c_int),
    "rama_ctx_weight_bytes": ([vp, C.POINTER(sz)], C.c_int),
    "rama_ctx_mem_info": ([vp, C.POINTER(sz), C.POINTER(sz)], C.c_int),
    "rama_session_create": ([vp, C.POINTER(vp)], C.c_int),
    "rama_session_reset": ([vp], C.c_int),
    "rama_session_destroy": ([vp], C.c_int),
    "rama_forward": ([vp, C.c_int32, C.c_int32], C.c_int),
    "rama_sample": ([vp, C.c_float, C.c_float, ip], C.c_int),
    "rama_generate": ([vp, ip, C.c_int32, C.c_int32, C.c_float, C.c_float, ip, fp], C.c_int),
    "rama_session_sync": ([vp], C.c_int),
    "rama_batch_create": ([vp, C.c_int32, C.POINTER(vp)], C.c_int),
    "rama_batch_destroy": ([vp], C.c_int),
    "rama_forward_batch": ([vp, C.POINTER(vp), ip, ip, C.c_int32], C.c_int),
    "rama_sample_batch": ([vp, C.POINTER(vp), C.c_int32, C.c_float, C.c_float, ip], C.c_int),
    "rama_generate_batch": ([vp, C.POINTER(vp), C.c_int32, C.POINTER(ip), ip, C.c_int32, C.c_float, C.c_float, ip, fp], C.c_int),
    "rama_batch_sync": ([vp], C.c_int),
    "rama_batch_launches_per_step": ([vp, ip], C.c_int),
    "rama_prefill": ([vp, ip, C.c_int32, C.c_int32, fp, fp, ip], C.c_int),
    "rama_session_set_prefill": ([vp, C.c_int32], C.c_int),
    "rama_state_to_host": ([vp, C.c_int, fp, sz, C.POINTER(sz)], C.c_int),
    "rama_logits_to_host": ([vp, fp, sz], C.c_int),
    "rama_session_set_debug": ([vp, C.c_int], C.c_int),
    "rama_session_launches_per_step": ([vp, C.POINTER(C.c_int)], C.c_int),
    "rama_step_trace": ([vp, C.c_int32, C.c_int32, C.POINTER(C.c_longlong), C.c_int32, ip], C.c_int),
    "rama_profile_step": ([vp, C.c_int32, C.c_int32, fp, ip], C.c_int),
    "rama_step_timeline": ([vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), ip, C.c_int32, ip], C.c_int),
    "rama_tokenizer_load": ([C.c_char_p, C.c_int32, C.POINTER(vp)], C.c_int),
    "rama_tokenizer_free": ([vp], C.c_int),
    "rama_tokenizer_info": ([vp, ip, ip], C.c_int),
    "rama_tokenizer_encode": ([vp, C.c_char_p, ip, C.c_int32, ip], C.c_int),
    "rama_tokenizer_decode": ([vp, C.c_int32, C.c_char_p, C.c_int32, ip], C.c_int),
    "rama_dev_alloc": ([vp, sz, C.POINTER(fp)], C.c_int),
    "rama_dev_free": ([vp, fp], C.c_int),
    "rama_dev_h2d": ([vp, fp, fp, sz], C.c_int),
    "rama_dev_d2h": ([vp, fp, fp, sz], C.c_int),
    "rama_ctx_sync": ([vp], C.c_int),
    "rama_op_array_add": ([vp, fp, fp, sz], C.c_int),
    "rama_op_array_mult": ([vp, fp, fp, sz], C.c_int),
    "rama_op_sinu": ([vp, fp, sz], C.c_int),
    "rama_op_multi_head_attention": ([vp, fp, fp, fp, fp, fp, C.POINTER(CConfig), C.c_int32, C.c_int32], C.c_int),
    "rama_op_copy_from_slice": ([vp, fp, fp, sz], C.c_int),
    "rama_op_rmsnorm": ([vp, fp, fp, fp, sz], C.c_int),
    "rama_op_apply_position": ([vp, fp, fp, fp, fp, sz], C.c_int),
    "rama_op_matmul": ([vp, fp, fp, fp, sz, sz, sz], C.c_int),
    "rama_op_matmul_nt": ([vp, fp, fp, fp, sz, sz, sz, C.c_int, C.c_int], C.c_int),
    "rama_bench_matmul_nt": ([vp, fp, fp, fp, sz, sz, sz, C.c_int, C.c_int, C.c_int, fp], C.c_int),
    "rama_debug_gemm_trace": ([vp, fp, fp, fp, sz, sz, sz, C.c_int, C.c_int, C.POINTER(C.c_longlong)], C.c_int),
    "rama_op_softmax": ([vp, fp, sz], C.c_int),
    "rama_op_sample": ([vp, fp, sz, C.c_float, C.c_float, ip], C.c_int),
    "rama_synth_fill": ([vp, fp, sz, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float], C.c_int),
    "rama_bench_gemv": ([vp, fp, fp, fp, sz, sz, sz, C.c_int, C.c_int, fp], C.c_int),
}

EXPORTS = sorted(_SIGS)


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RamaError(-2, f"{SO_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "or `make -C rama_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (args, res) in _SIGS.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RamaError(rc, (lib().rama_last_error() or b"").decode(errors="replace"))
