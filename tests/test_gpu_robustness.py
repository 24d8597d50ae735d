"""Call-order and loader robustness of the C ABI (findings of the round-1 review): sequences of entry points a caller
may legally issue must not depend on hidden state left by an earlier call or on the relative speed of two streams."""
import os
import struct

import numpy as np
import pytest

from oracle import ref
from rama_b200 import checkpoint as ck
from rama_b200.engine import GPU, Batch, RamaError, Session, generate
from conftest import GOLDEN
from util import LOGIT_TOL, model_tensors, rel_err

pytestmark = pytest.mark.gpu


def _pair(name, **kw):
    cfg, spec, tensors = model_tensors(name, **kw)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    return cfg, tensors, gpu, ref.Model(cfg, tensors)


def test_loader_many_tiny_pieces_with_delayed_readers(monkeypatch, tmp_path):
    """The pread → pinned ring → DMA pipeline with 4 KB pieces (thousands of them, every ring slot reused hundreds of
    times) and reader threads delayed between claiming a piece and taking its slot: every tensor must still land
    bit-exactly where the plan says (a reader of piece i + 8 must never take the slot piece i is waiting for)."""
    monkeypatch.setenv("RAMA_LOAD_PIECE_KB", "4")
    monkeypatch.setenv("RAMA_LOAD_TEST_DELAY_US", "300")
    for name in ("ref_hs48", "ref_untied"):
        path = os.path.join(GOLDEN, name + ".bin")
        cfg, want = ck.read_checkpoint(path)
        gpu = GPU(0)
        gpu.load_file(path)
        for t in ck.TENSORS:
            if want[t].size:
                assert gpu.weight_shard(t).tobytes() == np.ascontiguousarray(want[t]).tobytes(), (name, t)
        gpu.close()
    # and a synthetic model large enough for ~15k pieces
    cfg = ck.CONFIGS["stories15M"]
    tensors = ref.synth_tensors(cfg, ck.SynthSpec(seed=5))
    path = str(tmp_path / "s15.bin")
    ck.write_checkpoint(path, cfg, tensors)
    gpu = GPU(0)
    gpu.load_file(path)
    for t in ("token_embedding_table", "wq", "wo", "w1", "w2", "w3", "rms_final_weight", "freq_cis_imag"):
        assert gpu.weight_shard(t).tobytes() == tensors[t].tobytes(), t
    gpu.close()


def test_header_with_int_min_vocab_is_rejected(tmp_path):
    path = str(tmp_path / "bad.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<7i", 64, 176, 2, 4, 4, -2 ** 31, 64))
        f.write(b"\0" * 4096)
    gpu = GPU(0)
    with pytest.raises(RamaError):
        gpu.load_file(path)
    gpu.close()


def test_reload_is_refused_while_sessions_or_batches_are_alive():
    cfg, tensors, gpu, om = _pair("tiny")
    s = Session(gpu)
    with pytest.raises(RamaError):
        gpu.load_host(cfg, tensors)           # the session's captured graphs hold the old weight pointers
    s.close()
    b = Batch(gpu, 4)
    with pytest.raises(RamaError):
        gpu.load_synthetic(cfg, ck.SynthSpec())
    b.close()
    gpu.load_host(cfg, tensors)               # nothing alive: allowed
    s = Session(gpu)
    want, _, _, _ = ref.generate(om, ref.State(om), [5, 6], 20, 0.0, 0.9)
    assert generate(s, [5, 6], 20, 0.0, 0.9) == list(want)
    s.close(); gpu.close()


@pytest.mark.parametrize("steps_kind", ["full_window", "short"])
def test_generate_then_prefill_then_sample_on_one_session(steps_kind):
    """rama_generate leaves no 'chained' state behind: a later prefill + sample (the continuation the header documents)
    and a later batched sample behave like on a fresh session.  With the sticky flag of round 1 the sampler wrote
    out_tokens[stale pos] — one element past the array after a full-window run — and returned a forced prompt token."""
    cfg, tensors, gpu, om = _pair("tiny-sep")
    s, fresh = Session(gpu), Session(gpu)
    steps = cfg.seq_len if steps_kind == "full_window" else 6
    prompt = [9, 8, 7, 6, 5, 4, 3, 2]          # longer than the short run: every sampled token there is a forced one
    generate(s, prompt, steps, 0.0, 0.9)
    rng = np.random.default_rng(2)
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, 17)]
    os_ = ref.State(om)
    for pos, t in enumerate(toks):
        ref.forward(om, os_, t, pos)
    for x in (s, fresh):
        x.prefill(toks, 0)
    assert rel_err(s.logits(), os_.logits) < LOGIT_TOL
    want = int(np.flatnonzero(os_.logits == os_.logits.max())[-1])
    assert s.sample(0.0, 0.9) == want == fresh.sample(0.0, 0.9)
    assert s.sample(0.8, 0.9) == fresh.sample(0.8, 0.9)
    # the same through the batched sampler (out_tokens is NULL there)
    generate(s, prompt, steps, 0.0, 0.9)
    batch = Batch(gpu, 2)
    batch.forward([s, fresh], [toks[0], toks[0]], [0, 0])
    assert batch.sample([s, fresh], 0.0, 0.9) == [fresh.sample(0.0, 0.9)] * 2
    batch.close(); s.close(); fresh.close(); gpu.close()


def test_batch_and_session_streams_are_ordered_without_explicit_syncs():
    """rama_forward_batch runs on the batch's stream, rama_forward / rama_sample / rama_logits_to_host on the session's:
    the header promises they compose.  No rama_batch_sync / rama_session_sync anywhere in this test."""
    cfg, tensors, gpu, om = _pair("tiny")
    B = 6
    rng = np.random.default_rng(4)
    streams = [[1] + [int(t) for t in rng.integers(0, cfg.vocab_size, 24)] for _ in range(B)]
    sess = [Session(gpu) for _ in range(B)]
    sts = [ref.State(om) for _ in range(B)]
    batch = Batch(gpu, 8)
    for rep in range(3):
        base = rep * 8
        # (1) asynchronous per-session forwards, immediately followed by a batched step on the same sessions
        for k in range(3):
            for i in range(B):
                sess[i].forward(streams[i][base + k], base + k)
                ref.forward(om, sts[i], streams[i][base + k], base + k)
        for k in range(3, 6):
            batch.forward(sess, [streams[i][base + k] for i in range(B)], [base + k] * B)
            for i in range(B):
                ref.forward(om, sts[i], streams[i][base + k], base + k)
        # (2) per-session reads right after the batched step, no batch.sync()
        for i in range(B):
            assert rel_err(sess[i].logits(), sts[i].logits) < LOGIT_TOL, (rep, i)
        batch.forward(sess, [streams[i][base + 6] for i in range(B)], [base + 6] * B)
        nxt = [sess[i].sample(0.0, 0.9) for i in range(B)]           # session stream, right behind the batch stream
        for i in range(B):
            ref.forward(om, sts[i], streams[i][base + 6], base + 6)
            assert nxt[i] == int(np.flatnonzero(sts[i].logits == sts[i].logits.max())[-1]), (rep, i)
        # (3) batched step → per-session forward on the same cache → batched sample
        for i in range(B):
            sess[i].forward(streams[i][base + 7], base + 7)
            ref.forward(om, sts[i], streams[i][base + 7], base + 7)
        got = batch.sample(sess, 0.0, 0.9)
        assert got == [int(np.flatnonzero(st.logits == st.logits.max())[-1]) for st in sts], rep
    for i in range(B):
        n = 24
        kb = sess[i].state("key_cache").reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]
        assert rel_err(kb, sts[i].key_cache.reshape(cfg.n_layers, cfg.seq_len, -1)[:, :n]) < 1e-4
    batch.close()
    for s in sess:
        s.close()
    gpu.close()


def test_pooled_session_does_not_carry_a_batched_step_error():
    """A step whose top-p candidate list is empty (topp = -1e9 ⇒ cutoff > every p; the reference panics, infer.rs:66)
    fails THAT step; the same sessions reused in a later batched step start clean (server: sessions return to a pool)."""
    cfg, tensors, gpu, om = _pair("tiny")
    a, b, fresh = Session(gpu), Session(gpu), Session(gpu)
    batch = Batch(gpu, 2)
    batch.forward([a, b], [1, 1], [0, 0])
    with pytest.raises(RamaError):
        batch.sample([a, b], 1.0, -1e9)
    batch.forward([a, b], [1, 7], [0, 0])
    fresh.forward(7, 0)
    want_b = fresh.sample(0.0, 0.9)
    fresh.forward(1, 0)
    assert batch.sample([a, b], 0.0, 0.9) == [fresh.sample(0.0, 0.9), want_b]
    batch.close(); a.close(); b.close(); fresh.close(); gpu.close()
