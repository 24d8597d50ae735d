"""Step- and sequence-level parity of the fused CUDA decode path vs the CPU oracle:
teacher-forced logits, full RunState dumps (≙ Device::to_cpu), 256 greedy tokens, and the
reference-shaped per-op forward.  Everything goes through the C ABI (rama_b200.engine → ctypes)."""
import os

import numpy as np
import pytest

from oracle import ref
from rama_b200 import _lib, checkpoint as ck
from rama_b200.engine import (GPU, DeviceRunState, DeviceWeights, Session, forward_per_op, generate)
from conftest import GOLDEN
from util import LOGIT_TOL, model_tensors, rel_err

pytestmark = pytest.mark.gpu

PROMPT = [10646, 2501, 263, 931]  # "once upon a time" under the reference tokenizer.bin (tests/test_tokenizer.py)


def _pair(name, **kw):
    cfg, spec, tensors = model_tensors(name, **kw)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    return cfg, tensors, gpu, ref.Model(cfg, tensors)


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
def test_teacher_forced_logits_and_state(name):
    cfg, tensors, gpu, om = _pair(name)
    os_, sess = ref.State(om), Session(gpu)
    sess.set_debug(True)
    rng = np.random.default_rng(5)
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, cfg.seq_len - 1)]
    for pos, tok in enumerate(toks):
        ref.forward(om, os_, tok, pos)
        sess.forward(tok, pos)
        assert rel_err(sess.logits(), os_.logits) < LOGIT_TOL, pos
        if pos in (0, 1, cfg.seq_len // 2, cfg.seq_len - 1):
            st = sess.to_cpu()  # ≙ Device::to_cpu: all 12 RunState buffers
            for k in _lib.STATE:
                want = getattr(os_, k)
                if k == "att":  # only [h][0..pos] is defined
                    H, T = cfg.n_heads, cfg.seq_len
                    assert rel_err(st[k].reshape(H, T)[:, : pos + 1], want.reshape(H, T)[:, : pos + 1]) < 1e-4
                else:
                    assert rel_err(st[k], want) < 1e-4, (k, pos)
    sess.close(); gpu.close()


@pytest.mark.parametrize("name", ["ref_shared", "ref_untied", "ref_hs48"])
def test_golden_checkpoints_through_file_loader(name):
    """rama_ctx_load_file (mmap → HBM) on the .bin the reference's exporter wrote, against the
    logits the reference's torch model produced (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    gpu = GPU(0)
    cfg = gpu.load_file(os.path.join(GOLDEN, name + ".bin"))
    assert cfg.shared_weight == (name != "ref_untied")
    sess = Session(gpu)
    for pos, tok in enumerate(g["tokens"]):
        sess.forward(int(tok), pos)
        got = sess.logits()
        assert rel_err(got, g["logits"][pos]) < 1e-4, pos
        assert int(np.argmax(got)) == int(np.argmax(g["logits"][pos]))
    sess.close(); gpu.close()


def test_per_op_forward_matches_oracle_and_fused():
    """The reference-shaped forward() over the 11 Device ops (infer.rs:8-53)."""
    cfg, tensors, gpu, om = _pair("tiny-sep")
    wv, rsv = DeviceWeights(gpu, cfg, tensors), DeviceRunState(gpu, cfg)
    os_, sess = ref.State(om), Session(gpu)
    toks = [1, 7, 250, 3, 3, 299, 0, 42]
    for pos, tok in enumerate(toks):
        ref.forward(om, os_, tok, pos)
        forward_per_op(cfg, wv, rsv, tok, pos, gpu)
        sess.forward(tok, pos)
        per_op = gpu.to_cpu(rsv)
        assert rel_err(per_op["logits"], os_.logits) < LOGIT_TOL
        assert rel_err(sess.logits(), per_op["logits"]) < 1e-4
        for k in ("x", "xb", "xb2", "hb", "hb2", "q", "k", "v", "key_cache", "value_cache"):
            assert rel_err(per_op[k], getattr(os_, k)) < 1e-4, (k, pos)
    nxt = gpu.sample(cfg, rsv, 0.0, 0.9)
    assert nxt == int(ref.lib().ref_sample(ref.fptr(os_.logits.copy()), cfg.vocab_size, 0.0, 0.9))
    sess.close(); gpu.close()


def _greedy_case(name, steps, prompt, seed=1234):
    cfg, tensors, gpu, om = _pair(name, seed=seed, rms_jitter=0.0)  # the bench data spec
    want, want_logits, gap, _ = ref.generate(om, ref.State(om), prompt, steps, 0.0, 0.9, want_logits=True)
    sess = Session(gpu)
    got_dev = generate(sess, prompt, steps, 0.0, 0.9)                      # device-resident loop
    sess.reset()
    got_host = generate(sess, prompt, steps, 0.0, 0.9, host_loop=True)     # forward()+sample() per token
    # teacher-forced logits on the oracle's own token stream
    sess.reset()
    worst, token = 0.0, 1
    for pos in range(steps):
        sess.forward(token, pos)
        if pos % 16 == 0 or pos == steps - 1:
            worst = max(worst, rel_err(sess.logits(), want_logits[pos]))
        token = int(want[pos])
    sess.close(); gpu.close()
    return cfg, want, got_dev, got_host, gap, worst


@pytest.mark.parametrize("step", ["default", "cluster", "kernels"])
def test_stories15M_256_greedy_tokens_identical(step, monkeypatch):
    """BASELINE.json configs[0]: stories15M, f32, greedy, 256 tokens — identical token sequence; with the layers as one
    cluster-scope kernel (RAMA_STEP=cluster) and as the chain of kernels."""
    if step != "default":
        monkeypatch.setenv("RAMA_STEP", step)
    cfg, want, got_dev, got_host, gap, worst = _greedy_case("stories15M", 256, PROMPT)
    assert gap > 1e-4, f"seed gives a top-2 gap of {gap}: argmax not stable under f32 reordering"
    assert list(want[: len(PROMPT)]) == PROMPT
    assert got_dev == list(want), f"first mismatch at {next(i for i in range(256) if got_dev[i] != want[i])}"
    assert got_host == list(want)
    assert worst < LOGIT_TOL
    assert len(set(got_dev)) > 100  # non-degenerate decode


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_stories15M_other_seeds(seed):
    """seeds 1..4 next to the default 1234 — identical greedy tokens wherever the oracle's own top-1/top-2 gap
    makes the argmax stable under f32 reordering (a seed below the gap threshold is refused, not compared)."""
    cfg, tensors, gpu, om = _pair("stories15M", seed=seed, rms_jitter=0.0)
    want, want_logits, gap, _ = ref.generate(om, ref.State(om), PROMPT, 256, 0.0, 0.9, want_logits=True)
    sess = Session(gpu)
    got = generate(sess, PROMPT, 256, 0.0, 0.9)
    worst = rel_err(sess.logits(), want_logits[255]) if got == list(want) else None
    sess.close(); gpu.close()
    if gap < 1e-4:
        pytest.skip(f"seed {seed}: oracle top-2 gap {gap:.2e} < 1e-4 — argmax not stable, seed refused")
    assert got == list(want), f"first mismatch at {next(i for i in range(256) if got[i] != want[i])} (gap {gap:.2e})"
    assert worst < LOGIT_TOL


def test_reference_init_secondary_dataset():
    """The reference's own initialisation (model.py:231-247: N(0,0.02), wo/w3 scaled by 1/sqrt(2L)) as the secondary
    dataset of SURVEY §8d: greedy decode degenerates there (few distinct tokens), so the check is teacher-forced logits on
    random token ids at every step plus the greedy stream."""
    cfg, spec, tensors = model_tensors("stories15M", init="reference", rms_jitter=0.0)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    om = ref.Model(cfg, tensors)
    os_, sess = ref.State(om), Session(gpu)
    rng = np.random.default_rng(7)
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, 127)]
    for pos, tok in enumerate(toks):
        ref.forward(om, os_, tok, pos)
        sess.forward(tok, pos)
        got, want = sess.logits(), os_.logits
        # logits of this init are ~1e-2 in magnitude: use a relative bound on their own scale, not max(1, ·)
        assert float(np.max(np.abs(got - want))) < 1e-3 * max(float(np.max(np.abs(want))), 1e-6) + 1e-6, pos
    sess.close(); gpu.close()


def test_stories110M_256_greedy_tokens_identical():
    """BASELINE.json configs[1]: stories110M, f32, greedy — the north star's 256 identical tokens (device loop and the
    host-driven forward()+sample() loop, mod.rs:187-204), logits every 16th position within 1e-3."""
    cfg, want, got_dev, got_host, gap, worst = _greedy_case("stories110M", 256, PROMPT)
    assert gap > 1e-4, f"seed gives a top-2 gap of {gap}: argmax not stable under f32 reordering"
    assert got_dev == list(want), f"first mismatch at {next(i for i in range(256) if got_dev[i] != want[i])}"
    assert got_host == list(want)
    assert worst < LOGIT_TOL
    assert len(set(got_dev)) > 100


def test_llama7B_layer_shapes_greedy_and_long_positions():
    """7B layer geometry (D 4096, F 11008, H 32, hs 128, T 2048, separate classifier) with 2 layers."""
    cfg, want, got_dev, got_host, gap, worst = _greedy_case("l7-2layer", 24, PROMPT)
    assert got_dev == list(want) and got_host == list(want)
    assert worst < LOGIT_TOL
    # long positions exercise every split of the flash-decode kernel (cache rows before `pos` are
    # zero on both sides, so the comparison stays exact in meaning)
    cfg, tensors, gpu, om = _pair("l7-2layer", rms_jitter=0.0)
    os_, sess = ref.State(om), Session(gpu)
    for pos, tok in [(0, 1), (63, 17), (64, 400), (1000, 31999), (2047, 5)]:
        ref.forward(om, os_, tok, pos)
        sess.forward(tok, pos)
        assert rel_err(sess.logits(), os_.logits) < LOGIT_TOL, pos
    sess.close(); gpu.close()


def test_long_context_attention_with_live_cache():
    """Teacher-forced decode far into the context window with a LIVE KV cache: the cluster attention kernel folds
    1, 2, 3, 4 passes per CTA (positions < 256, < 512, < 768, < 1024) and positions >= 1024 go back to the split-merge
    kernel; logits vs the oracle at the boundaries of every regime."""
    cfg, tensors, gpu, om = _pair("tiny-long")
    os_, sess = ref.State(om), Session(gpu)
    rng = np.random.default_rng(11)
    toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, 1100)]
    check = {0, 1, 31, 32, 33, 254, 255, 256, 257, 300, 511, 512, 513, 700, 767, 768, 1000, 1023, 1024, 1025, 1099}
    worst = 0.0
    for pos, tok in enumerate(toks[:1100]):
        ref.forward(om, os_, tok, pos)
        sess.forward(tok, pos)
        if pos in check:
            e = rel_err(sess.logits(), os_.logits)
            worst = max(worst, e)
            assert e < LOGIT_TOL, (pos, e)
    assert rel_err(sess.to_cpu()["key_cache"], os_.key_cache) < 1e-4
    sess.close(); gpu.close()


@pytest.mark.parametrize("env", [{"RAMA_ATTN": "split"}, {"RAMA_ATTN_WO": "1"}, {"RAMA_ATTN_WO": "2"}, {"RAMA_PDL": "0"},
                                 {"RAMA_GEMV_STAGED": "0"}, {"RAMA_GEMV_STAGE_KB": "208"}, {"RAMA_EMBED_KERNEL": "0"},
                                 {"RAMA_STEP": "cluster"}, {"RAMA_STEP": "cluster", "RAMA_STEP_CLUSTER": "8"},
                                 {"RAMA_STEP": "kernels"}])
def test_kernel_variants_kept_as_options_agree_with_the_oracle(env, monkeypatch):
    """Every measured alternative that stays selectable at run time is held to the same parity
    bar as the default path: teacher-forced logits and the greedy stream on the tiny models."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)   # read at rama_ctx_create / rama_session_create
    for name in ("tiny", "tiny-sep"):
        cfg, tensors, gpu, om = _pair(name)
        os_, sess = ref.State(om), Session(gpu)
        rng = np.random.default_rng(3)
        toks = [1] + [int(t) for t in rng.integers(0, cfg.vocab_size, cfg.seq_len - 1)]
        for pos, tok in enumerate(toks):
            ref.forward(om, os_, tok, pos)
            sess.forward(tok, pos)
            assert rel_err(sess.logits(), os_.logits) < LOGIT_TOL, (env, name, pos)
        want, _, _, _ = ref.generate(om, ref.State(om), [5, 6], cfg.seq_len, 0.0, 0.9)
        assert generate(sess, [5, 6], cfg.seq_len, 0.0, 0.9) == list(want), (env, name)
        sess.close(); gpu.close()


def test_generate_edge_cases_and_errors():
    cfg, tensors, gpu, om = _pair("tiny")
    sess = Session(gpu)
    assert generate(sess, [], 0, 0.0, 0.9) == []                       # zero steps
    full = generate(sess, [], cfg.seq_len, 0.0, 0.9)                    # exactly seq_len steps, empty prompt
    want, _, _, _ = ref.generate(om, ref.State(om), [], cfg.seq_len, 0.0, 0.9)
    assert full == list(want)
    long_prompt = list(range(3, 3 + cfg.seq_len + 5))                   # prompt longer than steps: all forced
    assert generate(sess, long_prompt, 10, 0.0, 0.9) == long_prompt[:10]
    with pytest.raises(_lib.RamaError):                                 # reference: slice panic past seq_len
        sess.generate([], cfg.seq_len + 1)
    with pytest.raises(_lib.RamaError):
        sess.forward(1, cfg.seq_len)
    with pytest.raises(_lib.RamaError):
        sess.forward(cfg.vocab_size, 0)
    with pytest.raises(_lib.RamaError):
        sess.generate([cfg.vocab_size + 3], 4)
    # the session still works after rejected calls, and runs are reproducible bit for bit
    a = generate(sess, [5, 6], 20, 0.0, 0.9)
    la = sess.logits()
    b = generate(sess, [5, 6], 20, 0.0, 0.9)
    assert a == b and la.tobytes() == sess.logits().tobytes()
    sess.close(); gpu.close()


def test_unsupported_configs_are_rejected():
    gpu = GPU(0)
    bad = ck.Config(64, 176, 2, 4, 2, 512, 64, True)  # n_kv_heads != n_heads (GQA): reference forward ignores it
    with pytest.raises(_lib.RamaError):
        gpu.load_host(bad, ck.synth_tensors(bad))
    with pytest.raises(_lib.RamaError):
        Session(gpu)  # nothing loaded
    gpu.close()


def test_two_sessions_share_one_context():
    """Server threading contract (lib.rs:133-153): many RunStates, one device + weights."""
    cfg, tensors, gpu, om = _pair("tiny")
    s1, s2 = Session(gpu), Session(gpu)
    a = generate(s1, [5, 6, 7], 30, 0.0, 0.9)
    b = generate(s2, [100, 101], 30, 0.0, 0.9)
    wa, _, _, _ = ref.generate(om, ref.State(om), [5, 6, 7], 30, 0.0, 0.9)
    wb, _, _, _ = ref.generate(om, ref.State(om), [100, 101], 30, 0.0, 0.9)
    assert a == list(wa) and b == list(wb)
    s1.close(); s2.close(); gpu.close()


def test_synthetic_loader_is_bit_identical_to_host_recipe():
    cfg = ck.CONFIGS["tiny-sep"]
    spec = ck.SynthSpec(seed=99, rms_jitter=0.1)
    gpu = GPU(0)
    gpu.load_synthetic(cfg, spec)
    host = ck.synth_tensors(cfg, spec)
    for name in ck.TENSORS:
        assert gpu.weight_shard(name).tobytes() == host[name].tobytes(), name
    gpu.close()
