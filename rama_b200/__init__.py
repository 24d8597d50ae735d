"""rama_b200 — B200-native (sm_100a) implementation of rama's llama2 f32 decode path.

The product is the C-ABI library (include/rama_b200.h, rama_b200/csrc); this package holds the
ctypes binding, the host-side mirror of the reference's Device/forward/generate interface
(engine.py) and the llama2.c checkpoint helpers (checkpoint.py).  Importing the package does not
load the CUDA library; the first call does, and fails loudly if it is missing (no CPU fallback).
"""
from .checkpoint import CONFIGS, Config, SynthSpec  # noqa: F401

__version__ = "0.1.0"
