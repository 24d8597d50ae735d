"""Tensor-parallel shard plan (host mirror of set_config() in csrc/ctx.cu).

New relative to the reference, which is single-device (gpu.rs:215).  Classic Megatron layout on
llama2.c's row-major [out][in] matrices — the inverse of what the reference's
exporter does when it concatenates Meta's shards (engine/export/export.py:380-396):

  column-parallel (split `out` ⇒ contiguous row blocks):  wq, wk, wv (this rank's heads), w1, w3
  row-parallel    (split `in`  ⇒ column slices, repacked contiguous): wo, w2  → all-reduce after each
  classifier: vocabulary rows split (a window of the embedding when the classifier is shared)
  replicated: embedding, norms, RoPE tables
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .checkpoint import Config, TENSORS

# (Lc, R, C, r0, Rl, c0, Cl): local window [Lc][Rl][Cl] of the global tensor [Lc][R][C] at (r0, c0)
Plan = Tuple[int, int, int, int, int, int, int]


def check_divisible(cfg: Config, world: int) -> None:
    if cfg.n_heads % world or cfg.hidden_dim % world or cfg.vocab_size % world or (cfg.hidden_dim // world) % 4:
        raise ValueError(f"n_heads/hidden_dim/vocab_size not divisible by tp world {world}")


def shard_plan(cfg: Config, rank: int, world: int) -> Dict[str, Plan]:
    check_divisible(cfg, world)
    D, F, L, V, T = cfg.dim, cfg.hidden_dim, cfg.n_layers, cfg.vocab_size, cfg.seq_len
    hs2 = cfg.head_size // 2
    Dq, Fl, Vl = D // world, F // world, V // world
    full = lambda Lc, R, C: (Lc, R, C, 0, R, 0, C)
    rows = lambda Lc, R, C, r0, Rl: (Lc, R, C, r0, Rl, 0, C)
    cols = lambda Lc, R, C, c0, Cl: (Lc, R, C, 0, R, c0, Cl)
    return {
        "token_embedding_table": full(1, V, D),
        "rms_att_weight": full(1, L, D),
        "wq": rows(L, D, D, rank * Dq, Dq), "wk": rows(L, D, D, rank * Dq, Dq), "wv": rows(L, D, D, rank * Dq, Dq),
        "wo": cols(L, D, D, rank * Dq, Dq),
        "rms_ffn_weight": full(1, L, D),
        "w1": rows(L, F, D, rank * Fl, Fl), "w2": cols(L, D, F, rank * Fl, Fl), "w3": rows(L, F, D, rank * Fl, Fl),
        "rms_final_weight": full(1, 1, D),
        "freq_cis_real": full(1, T, hs2), "freq_cis_imag": full(1, T, hs2),
        "wcls": (1, V, D, 0, 0, 0, 0) if cfg.shared_weight else rows(1, V, D, rank * Vl, Vl),
    }


def shard_tensor(cfg: Config, name: str, full: np.ndarray, rank: int, world: int) -> np.ndarray:
    Lc, R, C, r0, Rl, c0, Cl = shard_plan(cfg, rank, world)[name]
    if Rl == 0 or Cl == 0:
        return np.zeros(0, dtype=np.float32)
    a = np.asarray(full, dtype=np.float32).reshape(Lc, R, C)
    return np.ascontiguousarray(a[:, r0:r0 + Rl, c0:c0 + Cl]).ravel()


def shard_all(cfg: Config, tensors: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    return {n: shard_tensor(cfg, n, tensors[n], rank, world) for n in TENSORS}
