"""ctypes binding of the CPU oracle (oracle/ref_cpu.cpp).  TEST INFRASTRUCTURE.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (rama_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libref_cpu.so")

STATE = ["x", "xb", "xb2", "hb", "hb2", "q", "k", "v", "att", "logits", "key_cache", "value_cache"]


class CConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads",
                                       "vocab_size", "seq_len", "shared_weight")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ref_cpu.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.ref_model_from_host.restype = C.c_void_p
        L.ref_model_from_host.argtypes = [C.POINTER(CConfig), C.POINTER(fp)]
        L.ref_model_from_file.restype = C.c_void_p
        L.ref_model_from_file.argtypes = [C.c_char_p]
        L.ref_model_config.argtypes = [C.c_void_p, C.POINTER(CConfig)]
        L.ref_model_tensor.restype = fp
        L.ref_model_tensor.argtypes = [C.c_void_p, C.c_int]
        L.ref_model_free.argtypes = [C.c_void_p]
        L.ref_state_create.restype = C.c_void_p
        L.ref_state_create.argtypes = [C.c_void_p]
        L.ref_state_free.argtypes = [C.c_void_p]
        L.ref_state_ptr.restype = fp
        L.ref_state_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.ref_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_generate.restype = C.c_double
        L.ref_generate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.POINTER(C.c_int32), fp, fp]
        L.ref_sample.restype = C.c_int
        L.ref_sample.argtypes = [fp, C.c_int, C.c_float, C.c_float]
        L.ref_sample_top_q.restype = C.c_int
        L.ref_sample_top_q.argtypes = [fp, C.c_int, C.c_float, C.c_float]
        L.ref_chacha_first_f32.restype = C.c_float
        L.ref_chacha_first_f32.argtypes = [C.c_uint64]
        L.ref_chacha_first_u32.restype = C.c_uint32
        L.ref_chacha_first_u32.argtypes = [C.c_uint64]
        L.ref_matmul.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int]
        L.ref_rmsnorm.argtypes = [fp, fp, fp, C.c_int]
        L.ref_softmax.argtypes = [fp, C.c_int]
        L.ref_sinu.argtypes = [fp, C.c_int]
        L.ref_array_add.argtypes = [fp, fp, C.c_int]
        L.ref_array_mult.argtypes = [fp, fp, C.c_int]
        L.ref_copy_from_slice.argtypes = [fp, fp, C.c_int]
        L.ref_apply_position.argtypes = [fp, fp, fp, fp, C.c_int]
        L.ref_multi_head_attention.argtypes = [fp, fp, fp, fp, fp, C.POINTER(CConfig), C.c_int, C.c_int]
        L.ref_synth_fill.argtypes = [fp, C.c_int64, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
        L.ref_tok_load.restype = C.c_void_p
        L.ref_tok_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_tok_free.argtypes = [C.c_void_p]
        L.ref_tok_max_len.argtypes = [C.c_void_p]
        L.ref_tok_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32), C.c_int]
        L.ref_tok_decode.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_set_reduce_mode.argtypes = [C.c_int]
        _lib = L
    return _lib


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def cconfig(cfg) -> CConfig:
    return CConfig(*cfg.as_i32())


class Model:
    """≙ TransformerWeights<Vec<f32>> + Config (views into caller-owned numpy arrays)."""

    def __init__(self, cfg, tensors: Dict[str, np.ndarray]):
        from rama_b200.checkpoint import TENSORS
        self.cfg = cfg
        self._keep = [np.ascontiguousarray(tensors[n], dtype=np.float32) for n in TENSORS]
        arr = (C.POINTER(C.c_float) * 14)()
        for i, a in enumerate(self._keep):
            arr[i] = fptr(a) if a.size else C.POINTER(C.c_float)()
        self._cc = cconfig(cfg)
        self.h = lib().ref_model_from_host(C.byref(self._cc), arr)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_model_free(self.h)
            self.h = None


class FileModel:
    def __init__(self, path: str):
        from rama_b200.checkpoint import Config
        self.h = lib().ref_model_from_file(path.encode())
        if not self.h:
            raise IOError(path)
        cc = CConfig()
        lib().ref_model_config(self.h, C.byref(cc))
        self.cfg = Config(cc.dim, cc.hidden_dim, cc.n_layers, cc.n_heads, cc.n_kv_heads,
                          cc.vocab_size, cc.seq_len, bool(cc.shared_weight))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_model_free(self.h)
            self.h = None


class State:
    """≙ RunState<Vec<f32>>; fields are numpy views of the oracle's buffers."""

    def __init__(self, model):
        self.model = model
        self.h = lib().ref_state_create(model.h)
        n = C.c_int64()
        for i, name in enumerate(STATE):
            p = lib().ref_state_ptr(self.h, i, C.byref(n))
            setattr(self, name, np.ctypeslib.as_array(p, shape=(n.value,)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_state_free(self.h)
            self.h = None


def forward(model, state: State, token: int, pos: int) -> None:
    lib().ref_forward(model.h, state.h, token, pos)


def generate(model, state: State, prompt: Sequence[int], steps: int, temperature: float = 0.0,
             topp: float = 0.9, want_logits: bool = False):
    """Returns (tokens[steps], logits[steps,V] or None, min top1-top2 gap, elapsed seconds)."""
    pr = np.asarray(list(prompt), dtype=np.int32)
    out = np.zeros(steps, dtype=np.int32)
    lg = np.zeros((steps, model.cfg.vocab_size), dtype=np.float32) if want_logits else None
    gap = C.c_float()
    el = lib().ref_generate(model.h, state.h, pr.ctypes.data_as(C.POINTER(C.c_int32)), len(pr), steps,
                            temperature, topp, out.ctypes.data_as(C.POINTER(C.c_int32)),
                            fptr(lg.reshape(-1)) if want_logits else None, C.byref(gap))
    return out, lg, gap.value, el


def synth_fill(n: int, seed: int, tensor_id: int, scale: float, offset: float = 0.0,
               out: Optional[np.ndarray] = None) -> np.ndarray:
    if out is None:
        out = np.empty(n, dtype=np.float32)
    lib().ref_synth_fill(fptr(out), n, seed, tensor_id, scale, offset)
    return out


def synth_tensors(cfg, spec) -> Dict[str, np.ndarray]:
    """Fast (threaded C++) equivalent of rama_b200.checkpoint.synth_tensors."""
    from rama_b200.checkpoint import T, rope_tables
    cos, sin = rope_tables(cfg.seq_len, cfg.head_size)
    out = {}
    for name, n, kind, scale, offset in spec.plan(cfg):
        if kind == "synth":
            out[name] = synth_fill(n, spec.seed, T[name], scale, offset)
        elif kind == "rope_real":
            out[name] = cos
        elif kind == "rope_imag":
            out[name] = sin
        else:
            out[name] = np.zeros(0, dtype=np.float32)
    return out
