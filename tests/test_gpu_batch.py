"""Batched multi-sequence decode (tensor-core GEMMs, split-K, per-sequence KV caches) vs the per-token
path and the CPU oracle.  The reference runs one forward()/sample() loop per request (lib.rs:127-160), so
the contract is: after rama_forward_batch every session is in the state its own forward(token, pos) call
would have left (logits within the 1e-3 north-star tolerance, cache rows tighter)."""
import numpy as np
import pytest

from oracle import ref
from rama_b200 import checkpoint as ck
from rama_b200.engine import GPU, Batch, RamaError, Session
from util import LOGIT_TOL, model_tensors, rel_err

pytestmark = pytest.mark.gpu


def _pair(name, **kw):
    cfg, spec, tensors = model_tensors(name, **kw)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    return cfg, tensors, gpu, ref.Model(cfg, tensors)


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
@pytest.mark.parametrize("B", [1, 5, 64])
def test_batched_steps_equal_per_session_forward(name, B):
    cfg, tensors, gpu, om = _pair(name)
    rng = np.random.default_rng(B)
    steps = 12
    start = [int(x) for x in rng.integers(0, 6, B)]  # sequences join the batch at different positions
    streams = [[1] + [int(t) for t in rng.integers(0, cfg.vocab_size, start[i] + steps)] for i in range(B)]
    solo = [Session(gpu) for _ in range(B)]
    bat = [Session(gpu) for _ in range(B)]
    states = [ref.State(om) for _ in range(min(B, 3))]
    for i in range(B):  # warm the caches up to each sequence's own start position through the batch-1 path
        for pos in range(start[i]):
            solo[i].forward(streams[i][pos], pos)
            bat[i].forward(streams[i][pos], pos)
            if i < len(states):
                ref.forward(om, states[i], streams[i][pos], pos)
        bat[i].sync(); solo[i].sync()
    batch = Batch(gpu, 64)
    for k in range(steps):
        toks = [streams[i][start[i] + k] for i in range(B)]
        pos = [start[i] + k for i in range(B)]
        batch.forward(bat, toks, pos)
        for i in range(B):
            solo[i].forward(toks[i], pos[i])
            if i < len(states):
                ref.forward(om, states[i], toks[i], pos[i])
        batch.sync()
        if k in (0, 5, steps - 1):
            for i in sorted({0, B // 2, B - 1}):
                assert rel_err(bat[i].logits(), solo[i].logits()) < LOGIT_TOL, (k, i)
            for i in range(len(states)):
                assert rel_err(bat[i].logits(), states[i].logits) < LOGIT_TOL, (k, i)
    nxt = batch.sample(bat, 0.0, 0.9)
    assert nxt == [s.sample(0.0, 0.9) for s in solo]
    nxt_t = batch.sample(bat, 0.8, 0.9)
    assert nxt_t == [s.sample(0.8, 0.9) for s in solo]
    for i in sorted({0, B - 1}):
        n = start[i] + steps
        ka = solo[i].state("key_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
        kb = bat[i].state("key_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
        va = solo[i].state("value_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
        vb = bat[i].state("value_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
        assert rel_err(kb, ka) < 1e-4 and rel_err(vb, va) < 1e-4
    batch.close()
    for s in solo + bat:
        s.close()
    gpu.close()


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
@pytest.mark.parametrize("temperature", [0.0, 0.8])
def test_generate_batch_equals_every_sequences_own_generate(name, temperature):
    """rama_generate_batch: the device-resident loop for several sequences at once — every row is the token stream the oracle's
    generate() (mod.rs:169-206) produces for that prompt, greedy and top-p, prompts of different lengths (one empty, one longer
    than the run), and the sessions are afterwards where their own loops would have left them."""
    cfg, tensors, gpu, om = _pair(name)
    rng = np.random.default_rng(17)
    steps = cfg.seq_len
    prompts = [[], [5], [int(t) for t in rng.integers(2, cfg.vocab_size, 7)], [int(t) for t in rng.integers(2, cfg.vocab_size, 19)],
               [int(t) for t in rng.integers(2, cfg.vocab_size, steps + 3)]]
    B = len(prompts)
    sess = [Session(gpu) for _ in range(B)]
    batch = Batch(gpu, 8)
    got, ms = batch.generate(sess, prompts, steps, temperature, 0.9)
    assert got.shape == (B, steps) and ms > 0
    for i, p in enumerate(prompts):
        want, _, _, _ = ref.generate(om, ref.State(om), p, steps, temperature, 0.9)
        assert [int(t) for t in got[i]] == [int(t) for t in want], (name, temperature, i)
    again, _ = batch.generate(sess, prompts, steps, temperature, 0.9)          # reusable, bit-reproducible
    assert (again == got).all()
    # the sessions hold their own caches: a per-session forward continues from there like after generate()
    st = ref.State(om)
    want0, _, _, _ = ref.generate(om, st, prompts[2], 10, 0.0, 0.9)
    g2, _ = batch.generate(sess[:3], prompts[:3], 10, 0.0, 0.9)
    assert [int(t) for t in g2[2]] == [int(t) for t in want0]
    sess[2].forward(int(want0[9]), 10); ref.forward(om, st, int(want0[9]), 10)
    assert rel_err(sess[2].logits(), st.logits) < LOGIT_TOL
    with pytest.raises(RamaError):
        batch.generate(sess, prompts, cfg.seq_len + 1)                         # the reference panics past seq_len
    batch.close()
    for s in sess:
        s.close()
    gpu.close()


def test_batch_errors():
    cfg, tensors, gpu, om = _pair("tiny")
    a, b = Session(gpu), Session(gpu)
    batch = Batch(gpu, 2)
    with pytest.raises(RamaError):
        batch.forward([a, a], [1, 1], [0, 0])                 # the same session twice
    with pytest.raises(RamaError):
        batch.forward([a, b], [1, 1], [0, cfg.seq_len])       # pos ≥ seq_len: the reference panics (infer.rs:32)
    with pytest.raises(RamaError):
        batch.forward([a, b], [1, cfg.vocab_size], [0, 0])    # token outside the vocabulary
    with pytest.raises(RamaError):
        Batch(gpu, 65)
    batch.close(); a.close(); b.close(); gpu.close()


def test_batch_64_at_7b_layer_shapes():
    """BASELINE config 5 geometry (dim 4096, ffn 11008, 32 heads, 64 concurrent sequences) on 2 layers: the batched
    tensor-core step against the CPU ORACLE (every sequence's own forward() stream, lib.rs:127-160) for six sequences
    spread over the batch, and against the per-token CUDA path."""
    cfg = ck.CONFIGS["l7-2layer"]
    spec = ck.SynthSpec()
    gpu = GPU(0)
    gpu.load_synthetic(cfg, spec)
    om = ref.Model(cfg, ref.synth_tensors(cfg, spec))  # bit-identical to the device generator
    B, steps = 64, 6
    rng = np.random.default_rng(1)
    streams = [[1] + [int(t) for t in rng.integers(0, cfg.vocab_size, steps)] for _ in range(B)]
    bat = [Session(gpu) for _ in range(B)]
    check = [0, 17, 63]
    oracle_check = [0, 9, 17, 31, 46, 63]
    solo = {i: Session(gpu) for i in check}
    ost = {i: ref.State(om) for i in oracle_check}
    batch = Batch(gpu, 64)
    for k in range(steps):
        batch.forward(bat, [streams[i][k] for i in range(B)], [k] * B)
        for i in check:
            solo[i].forward(streams[i][k], k)
        for i in oracle_check:
            ref.forward(om, ost[i], streams[i][k], k)
    batch.sync()
    for i in oracle_check:
        assert rel_err(bat[i].logits(), ost[i].logits) < LOGIT_TOL, i
        ko = ost[i].key_cache.reshape(cfg.n_layers, cfg.seq_len, -1)[:, :steps]
        vo = ost[i].value_cache.reshape(cfg.n_layers, cfg.seq_len, -1)[:, :steps]
        kb = bat[i].state("key_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :steps]
        vb = bat[i].state("value_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :steps]
        assert rel_err(kb, ko) < 1e-4 and rel_err(vb, vo) < 1e-4, i
    for i in check:
        assert rel_err(bat[i].logits(), solo[i].logits()) < LOGIT_TOL, i
    nxt = batch.sample(bat, 0.0, 0.9)
    for i in oracle_check:
        assert nxt[i] == int(np.flatnonzero(ost[i].logits == ost[i].logits.max())[-1]), i
    for i in check:
        assert nxt[i] == solo[i].sample(0.0, 0.9)
    print("launches per batched step", batch.launches_per_step())
    batch.close()
    for s in bat + list(solo.values()):
        s.close()
    gpu.close()
