"""Host-side mirror of the reference's engine interface for the decode path, over the C ABI.

Same names, argument meaning and error behaviour as the reference so that the parity tests read
like tests of the reference itself:

    reference (Rust)                                     here
    ---------------------------------------------------  -----------------------------------------
    device::gpu::GPU::new()            gpu.rs:213-234     GPU(device=0, tp=None)
    trait Device<T> (11 methods)       device.rs:3-24     GPU.array_add … GPU.to_cpu  (op level)
    Storage / View / MutView           mod.rs:16-126      DeviceBuffer / View
    TransformerWeights::from_weight    hbm.rs:55-90       GPU.load_host / load_file / load_synthetic
    RunState::from_state               hbm.rs:19-34       Session(gpu)
    forward(cfg,wv,rsv,token,pos,dev)  infer.rs:8-53      forward(session, token, pos)   (fused, graph)
                                                          forward_per_op(...)            (trait level)
    Device::sample                     cpu.rs:155-179     Session.sample(temperature, topp)
    generate(...)                      mod.rs:169-206     generate(session, prompt_tokens, steps, …)

Errors: the reference panics (`unwrap`); here every failure raises RamaError.
PyTorch is not needed by this module; device memory is owned by the library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import CConfig, CTp, RamaError, check, fp, ip
from .checkpoint import Config, SynthSpec, TENSORS, T, rope_tables

__all__ = ["GPU", "Session", "Batch", "Tokenizer", "generate_text", "DeviceBuffer", "View", "forward", "forward_per_op", "generate",
           "RamaError", "DeviceWeights", "DeviceRunState"]


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(fp)


def _cconfig(cfg: Config) -> CConfig:
    return CConfig(*cfg.as_i32())


class DeviceBuffer:
    """≙ `impl Storage for CudaSlice<f32>` (hbm.rs:6-10): an owned device allocation."""

    def __init__(self, gpu: "GPU", n: int, host: Optional[np.ndarray] = None):
        self.gpu, self.n = gpu, int(n)
        p = fp()
        check(_lib.lib().rama_dev_alloc(gpu.h, self.n, C.byref(p)))
        self.addr = C.cast(p, C.c_void_p).value or 0
        if host is not None:
            h = _f32(host).ravel()
            assert h.size == self.n
            check(_lib.lib().rama_dev_h2d(gpu.h, self.ptr(0), _fptr(h), self.n))

    def length(self) -> int:
        return self.n

    def ptr(self, off: int = 0):
        return C.cast(C.c_void_p(self.addr + 4 * off), fp)

    def to_host(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float32)
        check(_lib.lib().rama_dev_d2h(self.gpu.h, _fptr(out), self.ptr(0), self.n))
        return out

    def free(self):
        if self.addr:
            _lib.lib().rama_dev_free(self.gpu.h, self.ptr(0))
            self.addr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class View:
    """≙ View / MutView (mod.rs:16-98): (storage, ABSOLUTE element range); `slice` is relative
    to the storage, not to the parent view (mod.rs:44-51)."""

    def __init__(self, data: DeviceBuffer, start: int = 0, stop: Optional[int] = None):
        self.data = data
        self.start = start
        self.stop = data.n if stop is None else stop

    def slice(self, start: int = 0, stop: Optional[int] = None) -> "View":
        return View(self.data, start, self.data.n if stop is None else stop)

    mut_slice = slice

    def as_view(self) -> "View":
        return View(self.data, self.start, self.stop)

    def __len__(self):
        return self.stop - self.start

    def ptr(self):
        return self.data.ptr(self.start)


class GPU:
    """≙ device::gpu::GPU — the device handle (and, once loaded, its weights)."""

    def __init__(self, device: int = 0, tp: Optional[Tuple[int, int, bytes]] = None):
        L = _lib.lib()
        self.h = C.c_void_p()
        ctp = None
        if tp is not None and tp[1] > 1:
            ctp = CTp()
            ctp.rank, ctp.world = tp[0], tp[1]
            C.memmove(ctp.nccl_id, tp[2], 128)
        check(L.rama_ctx_create(device, C.byref(ctp) if ctp is not None else None, C.byref(self.h)))
        self.cfg: Optional[Config] = None
        self.rank, self.world = (tp[0], tp[1]) if tp is not None else (0, 1)

    @classmethod
    def multi(cls, n_gpus: int, devices: Optional[Sequence[int]] = None) -> "GPU":
        """One handle over n_gpus devices of this process (rama_ctx_create_multi): tensor parallelism for callers that,
        like the reference's engine binary and server, are a single process holding a single `GPU`."""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        arr = None
        if devices is not None:
            arr = (C.c_int32 * n_gpus)(*[int(d) for d in devices])
        check(_lib.lib().rama_ctx_create_multi(n_gpus, arr, C.byref(self.h)))
        self.cfg = None
        self.rank, self.world = 0, n_gpus
        return self

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(_lib.lib().rama_tp_unique_id(buf))
        return bytes(buf)

    def close(self):
        if self.h:
            _lib.lib().rama_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------
    def load_file(self, path: str) -> Config:
        check(_lib.lib().rama_ctx_load_file(self.h, path.encode()))
        return self._read_cfg()

    def load_host(self, cfg: Config, tensors: Dict[str, np.ndarray]) -> Config:
        keep = [_f32(tensors[n]).ravel() for n in TENSORS]
        arr = (fp * 14)()
        for i, a in enumerate(keep):
            arr[i] = _fptr(a) if a.size else fp()
        cc = _cconfig(cfg)
        check(_lib.lib().rama_ctx_load_host(self.h, C.byref(cc), arr))
        return self._read_cfg()

    def load_synthetic(self, cfg: Config, spec: SynthSpec = SynthSpec()) -> Config:
        scale = (C.c_float * 14)()
        offset = (C.c_float * 14)()
        for name, _n, kind, sc, off in spec.plan(cfg):
            if kind == "synth":
                scale[T[name]], offset[T[name]] = sc, off
        cos, sin = rope_tables(cfg.seq_len, cfg.head_size)
        cc = _cconfig(cfg)
        check(_lib.lib().rama_ctx_load_synthetic(self.h, C.byref(cc), spec.seed, scale, offset,
                                                 _fptr(cos), _fptr(sin)))
        return self._read_cfg()

    def _read_cfg(self) -> Config:
        cc = CConfig()
        check(_lib.lib().rama_ctx_config(self.h, C.byref(cc)))
        self.cfg = Config(cc.dim, cc.hidden_dim, cc.n_layers, cc.n_heads, cc.n_kv_heads, cc.vocab_size,
                          cc.seq_len, bool(cc.shared_weight))
        return self.cfg

    def weight_shard(self, name: str) -> np.ndarray:
        n = C.c_size_t()
        check(_lib.lib().rama_ctx_weight_to_host(self.h, T[name], None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        check(_lib.lib().rama_ctx_weight_to_host(self.h, T[name], _fptr(out), out.size, C.byref(n)))
        return out

    def weight_bytes(self) -> int:
        n = C.c_size_t()
        check(_lib.lib().rama_ctx_weight_bytes(self.h, C.byref(n)))
        return n.value

    def sync(self):
        check(_lib.lib().rama_ctx_sync(self.h))

    def mem_info(self) -> Tuple[int, int]:
        """(free, total) bytes of HBM on this device."""
        f, t = C.c_size_t(), C.c_size_t()
        check(_lib.lib().rama_ctx_mem_info(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    # ---- trait Device<T> (device.rs:3-24) ---------------------------------------------------
    def array_add(self, target: View, source: View, n: int):
        check(_lib.lib().rama_op_array_add(self.h, target.ptr(), source.ptr(), n))

    def array_mult(self, target: View, source: View, n: int):
        check(_lib.lib().rama_op_array_mult(self.h, target.ptr(), source.ptr(), n))

    def sinu(self, o: View, n: int):
        check(_lib.lib().rama_op_sinu(self.h, o.ptr(), n))

    def multi_head_attention(self, rsv: "DeviceRunState", cfg: Config, layer: int, pos: int,
                             keep_att: bool = True):
        cc = _cconfig(cfg)
        check(_lib.lib().rama_op_multi_head_attention(
            self.h, rsv.xb.ptr(), rsv.att.ptr() if keep_att else None, rsv.q.ptr(), rsv.key_cache.ptr(),
            rsv.value_cache.ptr(), C.byref(cc), layer, pos))

    def copy_from_slice(self, target: View, source: View, n: int):
        check(_lib.lib().rama_op_copy_from_slice(self.h, target.ptr(), source.ptr(), n))

    def rmsnorm(self, o: View, x: View, weight: View, n: int):
        check(_lib.lib().rama_op_rmsnorm(self.h, o.ptr(), x.ptr(), weight.ptr(), n))

    def apply_position(self, q: View, k: View, pos_real: View, pos_img: View, head_size: int):
        check(_lib.lib().rama_op_apply_position(self.h, q.ptr(), k.ptr(), pos_real.ptr(), pos_img.ptr(), head_size))

    def matmul(self, o: View, a: View, b: View, width: int, o_rows: int, o_cols: int):
        check(_lib.lib().rama_op_matmul(self.h, o.ptr(), a.ptr(), b.ptr(), width, o_rows, o_cols))

    def matmul_nt(self, o: View, a: View, b: View, M: int, N: int, K: int, variant: int = 0, flags: int = 0):
        """o[M][N] = a[M][K] · b[N][K]^T on the tensor cores (3xTF32); flags & 2 stores o^T."""
        check(_lib.lib().rama_op_matmul_nt(self.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, variant, flags))

    def softmax(self, x: View, n: int):
        check(_lib.lib().rama_op_softmax(self.h, x.ptr(), n))

    def sample(self, cfg: Config, rsv: "DeviceRunState", temperature: float, topp: float) -> int:
        nxt = C.c_int32()
        check(_lib.lib().rama_op_sample(self.h, rsv.logits.ptr(), cfg.vocab_size, temperature, topp, C.byref(nxt)))
        return nxt.value

    def to_cpu(self, rsv: "DeviceRunState") -> Dict[str, np.ndarray]:
        return {n: getattr(rsv, n).data.to_host() for n in _lib.STATE}


class DeviceWeights:
    """≙ TransformerWeights<CudaSlice<f32>> + TransformerWeightsView::from_gpu_ws (hbm.rs:55-121):
    one device allocation per tensor; `wcls` aliases the embedding when shared (state.rs:111-117)."""

    def __init__(self, gpu: GPU, cfg: Config, tensors: Dict[str, np.ndarray]):
        self.cfg = cfg
        self.wcls_exists = not cfg.shared_weight
        for name in TENSORS:
            a = _f32(tensors[name]).ravel()
            if name == "wcls" and not self.wcls_exists:
                a = np.ones(1, dtype=np.float32)  # ram.rs:46 stores a dummy vec![1.0]
            setattr(self, "_" + name, DeviceBuffer(gpu, a.size, a))
            setattr(self, name, View(getattr(self, "_" + name)))
        if not self.wcls_exists:
            self.wcls = View(self._token_embedding_table)


class DeviceRunState:
    """≙ RunState<CudaSlice<f32>>::from_state + RunStateView::from_rs (hbm.rs:19-34, state.rs:34-51)."""

    def __init__(self, gpu: GPU, cfg: Config):
        kv_dim = cfg.dim * cfg.n_kv_heads // cfg.n_heads
        sizes = dict(x=cfg.dim, xb=cfg.dim, xb2=cfg.dim, hb=cfg.hidden_dim, hb2=cfg.hidden_dim, q=cfg.dim,
                     k=cfg.dim, v=cfg.dim, att=cfg.n_heads * cfg.seq_len, logits=cfg.vocab_size,
                     key_cache=cfg.n_layers * cfg.seq_len * kv_dim, value_cache=cfg.n_layers * cfg.seq_len * kv_dim)
        for name, n in sizes.items():
            setattr(self, name, View(DeviceBuffer(gpu, n)))


def forward_per_op(cfg: Config, wv: DeviceWeights, rsv: DeviceRunState, token: int, pos: int, device: GPU):
    """Line-by-line mirror of the reference forward (infer.rs:8-53) over the op-level Device API.
    wq is issued once (the reference issues it twice, infer.rs:20-21; same result)."""
    dim, hidden_dim = cfg.dim, cfg.hidden_dim
    head_size = dim // cfg.n_heads
    device.copy_from_slice(rsv.x, wv.token_embedding_table.slice(token * dim, (token + 1) * dim), dim)
    pos_real = wv.freq_cis_real.slice(pos * (head_size // 2))
    pos_img = wv.freq_cis_imag.slice(pos * (head_size // 2))
    for layer in range(cfg.n_layers):
        device.rmsnorm(rsv.xb, rsv.x.as_view(), wv.rms_att_weight.slice(layer * dim), dim)
        device.matmul(rsv.q, wv.wq.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1)
        device.matmul(rsv.k, wv.wk.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1)
        device.matmul(rsv.v, wv.wv.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1)
        for h in range(cfg.n_heads):
            device.apply_position(rsv.q.mut_slice(h * head_size), rsv.k.mut_slice(h * head_size),
                                  pos_real, pos_img, head_size)
        lo = layer * cfg.seq_len * dim
        device.copy_from_slice(rsv.key_cache.mut_slice(lo + pos * dim, lo + (pos + 1) * dim), rsv.k.as_view(), dim)
        device.copy_from_slice(rsv.value_cache.mut_slice(lo + pos * dim, lo + (pos + 1) * dim), rsv.v.as_view(), dim)
        device.multi_head_attention(rsv, cfg, layer, pos)
        device.matmul(rsv.xb2, wv.wo.slice(layer * dim * dim), rsv.xb.as_view(), dim, dim, 1)
        device.array_add(rsv.x, rsv.xb2.as_view(), dim)
        device.rmsnorm(rsv.xb, rsv.x.as_view(), wv.rms_ffn_weight.slice(layer * dim), dim)
        device.matmul(rsv.hb, wv.w1.slice(layer * hidden_dim * dim), rsv.xb.as_view(), dim, hidden_dim, 1)
        device.matmul(rsv.hb2, wv.w3.slice(layer * hidden_dim * dim), rsv.xb.as_view(), dim, hidden_dim, 1)
        device.sinu(rsv.hb, hidden_dim)
        device.array_mult(rsv.hb, rsv.hb2.as_view(), hidden_dim)
        device.matmul(rsv.xb, wv.w2.slice(layer * dim * hidden_dim), rsv.hb.as_view(), hidden_dim, dim, 1)
        device.array_add(rsv.x, rsv.xb.as_view(), dim)
    device.copy_from_slice(rsv.xb, rsv.x.as_view(), dim)
    device.rmsnorm(rsv.x, rsv.xb.as_view(), wv.rms_final_weight, dim)
    device.matmul(rsv.logits, wv.wcls, rsv.x.as_view(), dim, cfg.vocab_size, 1)


class Session:
    """≙ RunState<Dev> for the fused path: KV cache, activations, stream and step graph in HBM."""

    def __init__(self, gpu: GPU):
        if gpu.cfg is None:
            raise RamaError(-5, "no weights loaded")
        self.gpu, self.cfg = gpu, gpu.cfg
        self.h = C.c_void_p()
        check(_lib.lib().rama_session_create(gpu.h, C.byref(self.h)))

    def close(self):
        if self.h:
            _lib.lib().rama_session_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(_lib.lib().rama_session_reset(self.h))

    def forward(self, token: int, pos: int):
        check(_lib.lib().rama_forward(self.h, token, pos))

    def sample(self, temperature: float, topp: float) -> int:
        nxt = C.c_int32()
        check(_lib.lib().rama_sample(self.h, temperature, topp, C.byref(nxt)))
        return nxt.value

    def sync(self):
        check(_lib.lib().rama_session_sync(self.h))

    def prefill(self, tokens: Sequence[int], pos0: int = 0, profile: bool = False):
        """Tensor-core prefill of `tokens` at positions [pos0, pos0+len): leaves the KV cache and the last
        position's logits as len(tokens) forward() calls would (mod.rs:187-192).  Returns
        (elapsed_ms, {kind: ms} or None, kernel launches)."""
        t = np.asarray(list(tokens), dtype=np.int32)
        ms = C.c_float()
        kinds = (C.c_float * _lib.PK_COUNT)()
        nl = C.c_int32()
        check(_lib.lib().rama_prefill(self.h, t.ctypes.data_as(ip), t.size, pos0, C.byref(ms),
                                      kinds if profile else None, C.byref(nl)))
        return ms.value, ({k: kinds[i] for i, k in enumerate(_lib.PREFILL_KINDS)} if profile else None), nl.value

    def set_prefill(self, min_rows: int):
        """generate(): prompts of at least min_rows rows (BOS included) use prefill; 0 = never."""
        check(_lib.lib().rama_session_set_prefill(self.h, min_rows))

    def generate(self, prompt: Sequence[int], steps: int, temperature: float = 0.0, topp: float = 0.9):
        """Device-resident loop. Returns (tokens[steps], elapsed_ms of the step loop)."""
        pr = np.asarray(list(prompt), dtype=np.int32)
        out = np.zeros(max(steps, 1), dtype=np.int32)
        ms = C.c_float()
        check(_lib.lib().rama_generate(self.h, pr.ctypes.data_as(ip) if pr.size else None, pr.size, steps,
                                       temperature, topp, out.ctypes.data_as(ip), C.byref(ms)))
        return out[:steps], ms.value

    def logits(self) -> np.ndarray:
        out = np.empty(self.cfg.vocab_size, dtype=np.float32)
        check(_lib.lib().rama_logits_to_host(self.h, _fptr(out), out.size))
        return out

    def set_debug(self, keep_att: bool):
        check(_lib.lib().rama_session_set_debug(self.h, int(keep_att)))

    def state(self, name: str) -> np.ndarray:
        i = _lib.STATE.index(name)
        n = C.c_size_t()
        check(_lib.lib().rama_state_to_host(self.h, i, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        check(_lib.lib().rama_state_to_host(self.h, i, _fptr(out), out.size, C.byref(n)))
        return out

    def to_cpu(self) -> Dict[str, np.ndarray]:
        """≙ Device::to_cpu (gpu.rs:196-209)."""
        return {n: self.state(n) for n in _lib.STATE}

    def launches_per_step(self) -> int:
        n = C.c_int()
        check(_lib.lib().rama_session_launches_per_step(self.h, C.byref(n)))
        return n.value

    def profile_step(self, token: int, pos: int) -> Dict[str, Tuple[float, int]]:
        ms = (C.c_float * _lib.K_COUNT)()
        ln = (C.c_int32 * _lib.K_COUNT)()
        check(_lib.lib().rama_profile_step(self.h, token, pos, ms, ln))
        return {k: (ms[i], ln[i]) for i, k in enumerate(_lib.KERNEL_KINDS)}


def step_timeline(session: "Session", token: int, pos: int, mode: int = 0, reps: int = 5):
    """In-graph per-kernel timeline of one decode step (rama_step_timeline).  Returns a list of dicts
    {kind, entry, ready, pro, end} (ns relative to the first stamp) in launch order."""
    cap = 5 * session.cfg.n_layers + 8
    st = (C.c_double * (4 * cap))()
    kd = (C.c_int32 * cap)()
    n = C.c_int32()
    check(_lib.lib().rama_step_timeline(session.h, token, pos, mode, reps, st, kd, cap, C.byref(n)))
    return [{"kind": _lib.KERNEL_KINDS[kd[i]], "entry": st[4 * i], "ready": st[4 * i + 1], "pro": st[4 * i + 2], "end": st[4 * i + 3]}
            for i in range(n.value)]


def summarize_timeline(tl) -> Dict[str, Dict[str, float]]:
    """Per kernel kind, microseconds per step: `chain` = dependency-resolved of this kernel → dependency-resolved of the next
    (the kernel's share of the step's critical path), split into `prologue` (ready → activations staged / peer partials
    arrived), `body` (→ CTA 0 done) and `tail` (→ next kernel's dependency resolved: the other CTAs, the grid drain and the
    dependent-launch hand-over); `early` = how long the kernel was resident before its dependency resolved."""
    out: Dict[str, Dict[str, float]] = {}
    for i, k in enumerate(tl):
        nxt = tl[i + 1]["ready"] if i + 1 < len(tl) else k["end"]
        a = out.setdefault(k["kind"], {"n": 0, "chain": 0.0, "prologue": 0.0, "body": 0.0, "tail": 0.0, "early": 0.0})
        a["n"] += 1
        a["chain"] += (nxt - k["ready"]) * 1e-3
        a["prologue"] += (k["pro"] - k["ready"]) * 1e-3
        a["body"] += (k["end"] - k["pro"]) * 1e-3
        a["tail"] += (nxt - k["end"]) * 1e-3
        a["early"] += (k["ready"] - k["entry"]) * 1e-3
    return out


class Tokenizer:
    """≙ tokenizer::bpe::Tokenizer (bpe.rs:9-97) + decode() (bpe.rs:102-116) over the C ABI (host code)."""

    def __init__(self, path: str, vocab_size: int):
        self.h = C.c_void_p()
        check(_lib.lib().rama_tokenizer_load(path.encode(), vocab_size, C.byref(self.h)))
        v, m = C.c_int32(), C.c_int32()
        check(_lib.lib().rama_tokenizer_info(self.h, C.byref(v), C.byref(m)))
        self.vocab_size, self.max_token_length = v.value, m.value

    def close(self):
        if self.h:
            _lib.lib().rama_tokenizer_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, text: str) -> List[int]:
        n = C.c_int32()
        b = text.encode("utf-8")
        check(_lib.lib().rama_tokenizer_encode(self.h, b, None, 0, C.byref(n)))
        out = (C.c_int32 * max(n.value, 1))()
        check(_lib.lib().rama_tokenizer_encode(self.h, b, out, n.value, C.byref(n)))
        return list(out[: n.value])

    def decode(self, token: int) -> bytes:
        n = C.c_int32()
        buf = C.create_string_buffer(256)
        check(_lib.lib().rama_tokenizer_decode(self.h, token, buf, 256, C.byref(n)))
        return buf.raw[: n.value]


def generate_text(session: "Session", tokenizer: Tokenizer, prompt: str, temperature: float, steps: int,
                  topp: float) -> str:
    """≙ generate(cfg, tokenizer, prompt, temperature, steps, topp, wv, rsv, device) -> String (mod.rs:169-206):
    encode the prompt (empty prompt → no tokens), run the loop, decode every `next` and concatenate."""
    prompt_tokens = tokenizer.encode(prompt) if len(prompt) > 0 else []
    toks = generate(session, prompt_tokens, steps, temperature, topp)
    return b"".join(tokenizer.decode(t) for t in toks).decode("utf-8", errors="replace")


class Batch:
    """Batched multi-sequence decode: one step for several Sessions at once (the server path — the
    reference runs one forward()/sample() loop per request, lib.rs:127-160)."""

    def __init__(self, gpu: GPU, max_seqs: int = 64):
        self.gpu = gpu
        self.h = C.c_void_p()
        check(_lib.lib().rama_batch_create(gpu.h, max_seqs, C.byref(self.h)))

    def close(self):
        if self.h:
            _lib.lib().rama_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _handles(sessions):
        arr = (C.c_void_p * len(sessions))()
        for i, s in enumerate(sessions):
            arr[i] = s.h
        return arr

    def forward(self, sessions: Sequence["Session"], tokens: Sequence[int], pos: Sequence[int]):
        """≙ forward(cfg, wv, rsv_i, token_i, pos_i, device) for every session i, in one pass."""
        n = len(sessions)
        t = np.asarray(list(tokens), dtype=np.int32)
        p = np.asarray(list(pos), dtype=np.int32)
        assert t.size == n and p.size == n
        check(_lib.lib().rama_forward_batch(self.h, self._handles(sessions), t.ctypes.data_as(ip), p.ctypes.data_as(ip), n))

    def sample(self, sessions: Sequence["Session"], temperature: float, topp: float) -> List[int]:
        n = len(sessions)
        out = np.zeros(n, dtype=np.int32)
        check(_lib.lib().rama_sample_batch(self.h, self._handles(sessions), n, temperature, topp, out.ctypes.data_as(ip)))
        return [int(x) for x in out]

    def generate(self, sessions: Sequence["Session"], prompts: Sequence[Sequence[int]], steps: int, temperature: float = 0.0,
                 topp: float = 0.9):
        """≙ generate() (mod.rs:169-206) for every session at once, token feedback on the device (rama_generate_batch).
        Returns (tokens[n][steps], elapsed_ms of the step loop)."""
        n = len(sessions)
        assert len(prompts) == n
        keep = [np.asarray(list(p), dtype=np.int32) for p in prompts]
        ptrs = (ip * n)(*[k.ctypes.data_as(ip) if k.size else ip() for k in keep])
        lens = np.asarray([k.size for k in keep], dtype=np.int32)
        out = np.zeros((n, max(steps, 1)), dtype=np.int32)
        ms = C.c_float()
        check(_lib.lib().rama_generate_batch(self.h, self._handles(sessions), n, ptrs, lens.ctypes.data_as(ip), steps, temperature,
                                             topp, out.ctypes.data_as(ip), C.byref(ms)))
        return out[:, :steps] if steps else out[:, :0], ms.value

    def sync(self):
        check(_lib.lib().rama_batch_sync(self.h))

    def launches_per_step(self) -> int:
        n = C.c_int32()
        check(_lib.lib().rama_batch_launches_per_step(self.h, C.byref(n)))
        return n.value


def forward(session: Session, token: int, pos: int):
    """≙ forward(cfg, wv, rsv, token, pos, device) (infer.rs:8)."""
    session.forward(token, pos)


def generate(session: Session, prompt_tokens: Sequence[int], steps: int, temperature: float, topp: float,
             host_loop: bool = False) -> List[int]:
    """≙ generate() (mod.rs:169-206) on token ids.  host_loop=True drives forward()+sample() from the
    host exactly like the reference loop (one token id each way per step); otherwise the whole loop
    stays on the device."""
    if not host_loop:
        return list(session.generate(prompt_tokens, steps, temperature, topp)[0])
    token, pos, out = 1, 0, []
    while pos < steps:
        session.forward(token, pos)
        if pos < len(prompt_tokens):
            nxt = int(prompt_tokens[pos])
        else:
            nxt = session.sample(temperature, topp)
        out.append(nxt)
        token = nxt
        pos += 1
    session.sync()
    return out
