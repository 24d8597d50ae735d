"""Prompt prefill (tcgen05 3xTF32 GEMMs + causal attention) vs the per-token path and the CPU oracle.

The reference has no prefill: generate() pushes the prompt through forward() token by token
(mod.rs:187-192).  So the contract is state equivalence: after prefill(tokens) the KV cache rows and the
last-position logits must equal what len(tokens) forward() calls leave behind (north-star tolerance 1e-3;
the cache is checked tighter), and generation must continue identically."""
import numpy as np
import pytest

from oracle import ref
from rama_b200 import checkpoint as ck
from rama_b200.engine import GPU, Session, generate
from util import LOGIT_TOL, model_tensors, rel_err

pytestmark = pytest.mark.gpu


def _pair(name, **kw):
    cfg, spec, tensors = model_tensors(name, **kw)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    return cfg, tensors, gpu, ref.Model(cfg, tensors)


def _kv(sess, cfg, n):
    k = sess.state("key_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
    v = sess.state("value_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
    return k, v


# f32 CUDA-core kernel / tensor-core kernel as (query warps, key shares) = (4,1) (4,2) (4,4) (2,2) (1,4)
@pytest.mark.parametrize("attn", ["cuda", "mma41", "mma42", "mma44", "mma22", "mma14"])
@pytest.mark.parametrize("name", ["tiny", "tiny-sep", "tiny-long"])   # head sizes 16, 48, 32
@pytest.mark.parametrize("n,pos0", [(1, 0), (7, 0), (33, 0), (None, 0), (20, 5), (1, 9)])
def test_prefill_equals_sequential_forward_and_oracle(name, n, pos0, attn, monkeypatch):
    monkeypatch.setenv("RAMA_PREFILL_ATTN", attn)
    cfg, tensors, gpu, om = _pair(name)
    if n is None and cfg.seq_len > 256:
        n = 200 - pos0   # (tiny-long: 2048 positions; 200 rows = four 64-key tiles with a ragged last one)
    n = cfg.seq_len - pos0 if n is None else n
    rng = np.random.default_rng(11)
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, pos0 + n - 1)]
    a, b, os_ = Session(gpu), Session(gpu), ref.State(om)
    for pos, t in enumerate(toks):
        a.forward(t, pos)
        ref.forward(om, os_, t, pos)
    for pos in range(pos0):  # the rows before pos0 come from the per-token path (chunked prompt)
        b.forward(toks[pos], pos)
    b.prefill(toks[pos0:], pos0)
    ka, va = _kv(a, cfg, pos0 + n)
    kb, vb = _kv(b, cfg, pos0 + n)
    assert rel_err(kb, ka) < 1e-4 and rel_err(vb, va) < 1e-4
    assert rel_err(kb, os_.key_cache.reshape(cfg.n_layers, cfg.seq_len, -1)[:, : pos0 + n]) < 1e-4
    assert rel_err(b.logits(), a.logits()) < LOGIT_TOL
    assert rel_err(b.logits(), os_.logits) < LOGIT_TOL
    # and decoding continues from the prefilled state exactly like from the stepped one
    if pos0 + n < cfg.seq_len:
        nxt = a.sample(0.0, 0.9)
        assert b.sample(0.0, 0.9) == nxt
        a.forward(nxt, pos0 + n); b.forward(nxt, pos0 + n)
        assert rel_err(b.logits(), a.logits()) < LOGIT_TOL
    a.close(); b.close(); gpu.close()


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
@pytest.mark.parametrize("temperature", [0.0, 0.8])
def test_generate_with_prefill_matches_stepped_generate(name, temperature):
    cfg, tensors, gpu, om = _pair(name)
    rng = np.random.default_rng(3)
    prompt = [int(t) for t in rng.integers(2, cfg.vocab_size, 20)]
    a, b = Session(gpu), Session(gpu)
    a.set_prefill(0)        # reference-shaped loop: every prompt token is a step
    b.set_prefill(4)
    ta = generate(a, prompt, cfg.seq_len, temperature, 0.9)
    tb = generate(b, prompt, cfg.seq_len, temperature, 0.9)
    assert ta[: len(prompt)] == prompt and tb == ta
    want, _, _, _ = ref.generate(om, ref.State(om), prompt, cfg.seq_len, temperature, 0.9)
    assert tb == [int(x) for x in want]
    a.close(); b.close(); gpu.close()


def test_prefill_errors():
    cfg, tensors, gpu, om = _pair("tiny")
    s = Session(gpu)
    from rama_b200.engine import RamaError
    with pytest.raises(RamaError):
        s.prefill([1] * (cfg.seq_len + 1), 0)          # past seq_len: the reference panics (infer.rs:32)
    with pytest.raises(RamaError):
        s.prefill([1, cfg.vocab_size], 0)               # token outside the vocabulary (infer.rs:13)
    with pytest.raises(RamaError):
        s.prefill([], 0)
    s.close(); gpu.close()


@pytest.mark.parametrize("attn", ["mma", "cuda"])
def test_prefill_512_at_7b_layer_shapes(attn, monkeypatch):
    monkeypatch.setenv("RAMA_PREFILL_ATTN", attn)   # mma: the default — at this geometry the 8-warp tensor-core kernel (4 query warps × 2 key shares)
    _prefill_512_at_7b_layer_shapes(attn)


def _prefill_512_at_7b_layer_shapes(attn):
    """BASELINE config 4 geometry (dim 4096, ffn 11008, 32 heads, 512 prompt tokens) on 2 layers: the tensor-core
    prefill against the CPU ORACLE run token by token over the same 512 rows (K = 4096 / 11008 contractions — where a
    tcgen05 accumulation drift would show), and against 512 per-token steps of the decode path."""
    cfg = ck.CONFIGS["l7-2layer"]
    spec = ck.SynthSpec()
    gpu = GPU(0)
    gpu.load_synthetic(cfg, spec)
    om = ref.Model(cfg, ref.synth_tensors(cfg, spec))  # same integer recipe as the device generator (bit-identical)
    os_ = ref.State(om)
    rng = np.random.default_rng(7)
    n = 512
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, n - 1)]
    a, b = Session(gpu), Session(gpu)
    for pos, t in enumerate(toks):
        a.forward(t, pos)
        ref.forward(om, os_, t, pos)
    ms, kinds, launches = b.prefill(toks, 0, profile=True)
    print("prefill-512 l7-2layer attn", attn, "ms", ms, kinds, launches)
    ka, va = _kv(a, cfg, n)
    kb, vb = _kv(b, cfg, n)
    ko = os_.key_cache.reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
    vo = os_.value_cache.reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
    assert rel_err(kb, ko) < 1e-4 and rel_err(vb, vo) < 1e-4          # every cache row of every layer vs the oracle
    assert rel_err(b.logits(), os_.logits) < LOGIT_TOL                # last-position logits vs the oracle
    assert rel_err(kb, ka) < 1e-4 and rel_err(vb, va) < 1e-4
    assert rel_err(b.logits(), a.logits()) < LOGIT_TOL
    want = int(np.flatnonzero(os_.logits == os_.logits.max())[-1])    # later index wins (cpu.rs:165)
    assert b.sample(0.0, 0.9) == want == a.sample(0.0, 0.9)
    a.close(); b.close(); gpu.close()


@pytest.mark.parametrize("attn", ["mma41", "mma42", "mma44", "mma22", "mma14", "cuda"])
def test_prefill_chunk_at_head_size_128_ragged(attn, monkeypatch):
    """Head size 128 (the 7B head) with a ragged chunk: 150 rows appended at position 70 — query blocks and key tiles both end
    mid-tile, and the first key tiles lie wholly below every query of the chunk."""
    monkeypatch.setenv("RAMA_PREFILL_ATTN", attn)
    cfg = ck.CONFIGS["mid-4layer"]
    gpu = GPU(0)
    gpu.load_synthetic(cfg, ck.SynthSpec())
    rng = np.random.default_rng(17)
    pos0, n = 70, 150
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, pos0 + n - 1)]
    a, b = Session(gpu), Session(gpu)
    for pos, t in enumerate(toks):
        a.forward(t, pos)
    for pos in range(pos0):
        b.forward(toks[pos], pos)
    b.prefill(toks[pos0:], pos0)
    ka, va = _kv(a, cfg, pos0 + n)
    kb, vb = _kv(b, cfg, pos0 + n)
    assert rel_err(kb, ka) < 1e-4 and rel_err(vb, va) < 1e-4
    assert rel_err(b.logits(), a.logits()) < LOGIT_TOL
    assert b.sample(0.0, 0.9) == a.sample(0.0, 0.9)
    a.close(); b.close(); gpu.close()


def test_prefill_longer_than_one_chunk_stories110m_shapes():
    """700 prompt rows = two chunks (512 + 188) at stories110M geometry (seq_len 1024): the second chunk attends
    to the first chunk's cache rows (pos0 = 512)."""
    cfg = ck.CONFIGS["stories110M"]
    gpu = GPU(0)
    gpu.load_synthetic(cfg, ck.SynthSpec())
    rng = np.random.default_rng(13)
    n = 700
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, n - 1)]
    a, b = Session(gpu), Session(gpu)
    for pos, t in enumerate(toks):
        a.forward(t, pos)
    b.prefill(toks, 0)
    ka, va = _kv(a, cfg, n)
    kb, vb = _kv(b, cfg, n)
    assert rel_err(kb, ka) < 1e-4 and rel_err(vb, va) < 1e-4
    assert rel_err(b.logits(), a.logits()) < LOGIT_TOL
    # generate() with a long prompt takes the prefill path and continues identically
    prompt = toks[1:601]
    a.set_prefill(0); b.set_prefill(16)
    ta = generate(a, prompt, 640, 0.0, 0.9)
    tb = generate(b, prompt, 640, 0.0, 0.9)
    assert ta == tb and ta[:600] == prompt
    a.close(); b.close(); gpu.close()
