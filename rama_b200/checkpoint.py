"""llama2.c v0 ``.bin`` checkpoints: header, tensor order, synthetic weights, reader/writer.

Format (kept byte-for-byte): 7 little-endian i32 header, then f32 tensors in
``legacy_export`` order — reference engine/export/export.py:75-127 ⇔
engine/src/transformer/ram.rs:30-49; header parse engine/src/transformer/mod.rs:140-166
(``vocab > 0`` ⇒ classifier shared with the embedding).

Synthetic weights are produced by an integer-only recipe so that numpy (here), the C++
oracle (oracle/ref_cpu.cpp: ref_synth_fill) and the CUDA kernel (csrc/synth.cu) give
bit-identical tensors without shipping a 27 GB file:

    key = splitmix64(seed ^ tensor_id * 0xD1B54A32D192ED03)
    S   = sum of the eight 16-bit fields of splitmix64(key + 2i), splitmix64(key + 2i + 1)
    w_i = float32(2 S - 8*65535) * scale          (Irwin-Hall n=8, ≈ N(0, std²))

Only exact integer arithmetic, one exact int→f32 conversion and one IEEE multiply.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, asdict
from typing import Dict, Iterator, List, Tuple

import numpy as np

# file order of the 14 tensors
TENSORS = ["token_embedding_table", "rms_att_weight", "wq", "wk", "wv", "wo", "rms_ffn_weight",
           "w1", "w2", "w3", "rms_final_weight", "freq_cis_real", "freq_cis_imag", "wcls"]
T = {n: i for i, n in enumerate(TENSORS)}
MATRICES = ("token_embedding_table", "wq", "wk", "wv", "wo", "w1", "w2", "w3", "wcls")
NORMS = ("rms_att_weight", "rms_ffn_weight", "rms_final_weight")


@dataclass(frozen=True)
class Config:
    """engine/src/transformer/mod.rs:128-138"""
    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    vocab_size: int
    seq_len: int
    shared_weight: bool

    @property
    def head_size(self) -> int:
        return self.dim // self.n_heads

    def header(self) -> bytes:
        v = self.vocab_size if self.shared_weight else -self.vocab_size
        return struct.pack("<7i", self.dim, self.hidden_dim, self.n_layers, self.n_heads,
                           self.n_kv_heads, v, self.seq_len)

    @staticmethod
    def from_header(b: bytes) -> "Config":
        d, f, l, h, kv, v, t = struct.unpack("<7i", b[:28])
        return Config(d, f, l, h, kv, abs(v), t, v > 0)

    def as_i32(self) -> List[int]:
        return [self.dim, self.hidden_dim, self.n_layers, self.n_heads, self.n_kv_heads,
                self.vocab_size, self.seq_len, int(self.shared_weight)]

    def tensor_sizes(self) -> List[int]:
        D, F, L, V, S = self.dim, self.hidden_dim, self.n_layers, self.vocab_size, self.seq_len
        hs2 = self.head_size // 2
        return [V * D, L * D, L * D * D, L * D * D, L * D * D, L * D * D, L * D,
                L * F * D, L * D * F, L * F * D, D, S * hs2, S * hs2,
                0 if self.shared_weight else V * D]

    def file_bytes(self) -> int:
        return 28 + 4 * sum(self.tensor_sizes())

    def weight_bytes_per_token(self) -> int:
        """every matrix once (wq once), norms, classifier, one embedding row,
        one RoPE row."""
        D, F, L, V = self.dim, self.hidden_dim, self.n_layers, self.vocab_size
        return 4 * (L * (4 * D * D + 3 * D * F) + (2 * L + 1) * D + V * D + D + self.head_size)

    def kv_bytes(self, pos: int) -> int:
        """KV read 2·L·(pos+1)·D·4 + KV write 2·L·D·4."""
        return 2 * self.n_layers * (pos + 1) * self.dim * 4 + 2 * self.n_layers * self.dim * 4

    def avg_bytes_per_token(self, steps: int) -> float:
        return self.weight_bytes_per_token() + sum(self.kv_bytes(p) for p in range(steps)) / steps


CONFIGS: Dict[str, Config] = {
    # BASELINE.json configs[0..2]
    "stories15M": Config(288, 768, 6, 6, 6, 32000, 256, True),
    "stories110M": Config(768, 2048, 12, 12, 12, 32000, 1024, True),
    "llama2-7B": Config(4096, 11008, 32, 32, 32, 32000, 2048, False),
    # test-sized models (same code paths: shared / separate classifier, odd shapes)
    "tiny": Config(64, 176, 2, 4, 4, 512, 64, True),
    "tiny-sep": Config(96, 256, 3, 2, 2, 300, 48, False),
    # long context at test size (head_size 32): every pass count of the cluster attention kernel and the split-merge kernel
    "tiny-long": Config(64, 176, 2, 2, 2, 512, 2048, True),
    # per-SM weight slabs of 60-190 KB per kernel (what a rank of llama2-7B holds under TP = 4 / 8), 1 GPU
    "mid-4layer": Config(1536, 4096, 4, 12, 12, 32000, 1024, False),
    # 7B layer shapes with 2 layers: exercises the 7B kernel configuration cheaply
    "l7-2layer": Config(4096, 11008, 2, 32, 32, 32000, 2048, False),
}

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _sum16(h: np.ndarray) -> np.ndarray:
    m = np.uint64(0xFFFF)
    return ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m)
            + (h >> np.uint64(48))).astype(np.int64)


_STD_Z = math.sqrt(8.0 * (65536.0 ** 2 - 1.0) / 3.0)  # std of 2S - 8*65535


def synth_scale(std: float) -> np.float32:
    return np.float32(std / _STD_Z)


def synth_key(seed: int, tensor_id: int) -> int:
    k = (seed ^ ((tensor_id * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
    return int(_splitmix64(np.array([k], dtype=np.uint64))[0])


def synth_fill(n: int, seed: int, tensor_id: int, scale: np.float32, start: int = 0) -> np.ndarray:
    """Elements [start, start+n) of synthetic tensor `tensor_id`."""
    key = np.uint64(synth_key(seed, tensor_id))
    out = np.empty(n, dtype=np.float32)
    step = 1 << 22
    with np.errstate(over="ignore"):
        for o in range(0, n, step):
            m = min(step, n - o)
            i2 = (np.arange(start + o, start + o + m, dtype=np.uint64) * np.uint64(2)) + key
            s = _sum16(_splitmix64(i2)) + _sum16(_splitmix64(i2 + np.uint64(1)))
            out[o:o + m] = (2 * s - 8 * 65535).astype(np.float32) * scale
    return out


def rope_tables(seq_len: int, head_size: int, theta: float = 10000.0) -> Tuple[np.ndarray, np.ndarray]:
    """freqs_cos / freqs_sin as engine/export/model.py:41-47 (f32 throughout)."""
    idx = np.arange(0, head_size, 2, dtype=np.float32)[: head_size // 2]
    freqs = (np.float32(1.0) / np.power(np.float32(theta), idx / np.float32(head_size))).astype(np.float32)
    t = np.arange(seq_len, dtype=np.float32)
    ang = np.outer(t, freqs).astype(np.float32)
    return np.cos(ang).astype(np.float32).ravel(), np.sin(ang).astype(np.float32).ravel()


@dataclass(frozen=True)
class SynthSpec:
    """Value distribution of a synthetic checkpoint.

    init="unit": every matrix and the embedding ~N(0, 1/√D) — non-degenerate greedy decode.
    init="reference": model.py:231-247 — N(0, 0.02), wo/w3 N(0, 0.02/√(2L)).
    Norm weights are 1 + rms_jitter·N(0,1) (the bench spec uses 0; tests use >0 so that a
    mis-indexed norm weight is visible).
    """
    seed: int = 1234
    init: str = "unit"
    rms_jitter: float = 0.0

    def std(self, cfg: Config, name: str) -> float:
        if self.init == "unit":
            return cfg.dim ** -0.5
        if self.init == "reference":
            return 0.02 / math.sqrt(2 * cfg.n_layers) if name in ("wo", "w3") else 0.02
        raise ValueError(self.init)

    def plan(self, cfg: Config) -> List[Tuple[str, int, str, float, float]]:
        """[(name, n_elements, kind, scale, offset)] in file order.
        kind: 'synth' → offset + synth*scale ; 'rope_real'/'rope_imag' ; 'skip' (n = 0)."""
        out = []
        for name, n in zip(TENSORS, cfg.tensor_sizes()):
            if name in MATRICES:
                out.append((name, n, "synth" if n else "skip", float(synth_scale(self.std(cfg, name))), 0.0))
            elif name in NORMS:
                out.append((name, n, "synth", float(synth_scale(self.rms_jitter)), 1.0))
            elif name == "freq_cis_real":
                out.append((name, n, "rope_real", 0.0, 0.0))
            else:
                out.append((name, n, "rope_imag", 0.0, 0.0))
        return out


def synth_tensor(cfg: Config, spec: SynthSpec, name: str) -> np.ndarray:
    n = cfg.tensor_sizes()[T[name]]
    if name == "freq_cis_real":
        return rope_tables(cfg.seq_len, cfg.head_size)[0]
    if name == "freq_cis_imag":
        return rope_tables(cfg.seq_len, cfg.head_size)[1]
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    if name in NORMS:
        v = synth_fill(n, spec.seed, T[name], synth_scale(spec.rms_jitter))
        return (np.float32(1.0) + v).astype(np.float32)
    return synth_fill(n, spec.seed, T[name], synth_scale(spec.std(cfg, name)))


def synth_tensors(cfg: Config, spec: SynthSpec = SynthSpec()) -> Dict[str, np.ndarray]:
    return {name: synth_tensor(cfg, spec, name) for name in TENSORS}


def write_checkpoint(path: str, cfg: Config, tensors: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(cfg.header())
        for name, n in zip(TENSORS, cfg.tensor_sizes()):
            a = np.ascontiguousarray(tensors[name], dtype="<f4").ravel()
            if a.size != n:
                raise ValueError(f"{name}: {a.size} elements, expected {n}")
            a.tofile(f)


def read_checkpoint(path: str) -> Tuple[Config, Dict[str, np.ndarray]]:
    """mmap-backed views (no per-f32 reads — contrast engine/src/utils/read.rs:25-33)."""
    with open(path, "rb") as f:
        cfg = Config.from_header(f.read(28))
    mm = np.memmap(path, dtype="<f4", mode="r", offset=28)
    out, o = {}, 0
    for name, n in zip(TENSORS, cfg.tensor_sizes()):
        out[name] = mm[o:o + n]
        o += n
    if o > mm.size:
        raise ValueError("checkpoint truncated")
    return cfg, out


def iter_config_names() -> Iterator[str]:
    return iter(CONFIGS)


def config_dict(cfg: Config) -> dict:
    return asdict(cfg)
