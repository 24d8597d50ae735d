"""Pins the C++ oracle against outputs of the reference's own Python model + exporter
(fixtures made by tests/golden/make_golden.py from engine/export/{model,export}.py)."""
import os

import numpy as np
import pytest

from oracle import ref
from rama_b200 import checkpoint as ck
from conftest import GOLDEN

CASES = ["ref_shared", "ref_untied", "ref_hs48"]
SHARED = {"ref_shared", "ref_hs48"}   # classifier shares the embedding (vocab > 0 in the header)


@pytest.mark.parametrize("name", CASES)
def test_header_and_layout(name):
    cfg, tensors = ck.read_checkpoint(os.path.join(GOLDEN, name + ".bin"))
    assert cfg.shared_weight == (name in SHARED)
    assert cfg.file_bytes() == os.path.getsize(os.path.join(GOLDEN, name + ".bin"))
    # RoPE tables written by the reference exporter == our rope_tables() (model.py:41-47)
    cos, sin = ck.rope_tables(cfg.seq_len, cfg.head_size)
    # (f32 angle t·θ_i: one ulp of the angle is ~t·6e-8, so the bound grows with the window — 72 positions in ref_hs48)
    np.testing.assert_allclose(tensors["freq_cis_real"], cos, rtol=0, atol=5e-6)
    np.testing.assert_allclose(tensors["freq_cis_imag"], sin, rtol=0, atol=5e-6)
    m = ref.FileModel(os.path.join(GOLDEN, name + ".bin"))
    assert m.cfg == cfg


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("reduce_mode", [0, 1, 2])
def test_teacher_forced_logits_match_reference_model(name, reduce_mode):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    toks, want = g["tokens"], g["logits"]
    ref.lib().ref_set_reduce_mode(reduce_mode)
    try:
        m = ref.FileModel(os.path.join(GOLDEN, name + ".bin"))
        s = ref.State(m)
        for pos, tok in enumerate(toks):
            ref.forward(m, s, int(tok), pos)
            got = s.logits.copy()
            # torch (SDPA, different association) vs sequential f32 chains: observed ~1e-6
            np.testing.assert_allclose(got, want[pos], rtol=2e-4, atol=2e-5, err_msg=f"pos {pos}")
            assert int(np.argmax(got)) == int(np.argmax(want[pos]))
    finally:
        ref.lib().ref_set_reduce_mode(0)


@pytest.mark.parametrize("name", CASES)
def test_generate_forces_prompt_then_greedy(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    toks, want = g["tokens"], g["logits"]
    m = ref.FileModel(os.path.join(GOLDEN, name + ".bin"))
    s = ref.State(m)
    # prompt = golden tokens 1.. (token 0 is BOS which generate() supplies itself, mod.rs:182)
    prompt = [int(t) for t in toks[1:]]
    steps = len(toks)
    out, lg, gap, _ = ref.generate(m, s, prompt, steps, 0.0, 0.9, want_logits=True)
    assert list(out[: steps - 1]) == prompt          # forced (mod.rs:190-191)
    assert out[steps - 1] == int(np.argmax(want[steps - 1]))  # first free step = argmax
    np.testing.assert_allclose(lg, want, rtol=2e-4, atol=2e-5)
