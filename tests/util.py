"""Shared helpers for the parity tests (oracle on the CPU vs the CUDA path through the C ABI)."""
from __future__ import annotations

import functools
from typing import Dict

import numpy as np

from oracle import ref
from rama_b200 import checkpoint as ck


@functools.lru_cache(maxsize=4)
def model_tensors(name: str, seed: int = 1234, init: str = "unit", rms_jitter: float = 0.1):
    cfg = ck.CONFIGS[name]
    spec = ck.SynthSpec(seed=seed, init=init, rms_jitter=rms_jitter)
    return cfg, spec, ref.synth_tensors(cfg, spec)  # threaded C++ generator (== numpy, bitwise)


def rel_err(got: np.ndarray, want: np.ndarray) -> float:
    """max |got-want| / max(1, max|want|)  — the north star's 'max-abs/rel 1e-3' in one number."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want)) / max(1.0, float(np.max(np.abs(want)))))


LOGIT_TOL = 1e-3  # BASELINE.json north_star: logits within max-abs/rel 1e-3 of the CPU path


def rand(n: int, seed: int, scale: float = 1.0) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal(n) * scale).astype(np.float32)
