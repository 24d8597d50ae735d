"""The C++ host side above the C ABI (include/rama_b200.hpp: Config / View / RunState / TransformerWeights / trait Device /
GPU / forward / generate with the reference's names) and the engine CLI built on it (≙ engine/src/main.rs).
CPU: both programs build and link against the product library.  GPU: the mirror's forward() — fused and per-op — reproduces
the logits the reference's own torch model produced for the golden checkpoints, and the CLI prints the oracle's greedy text."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

HOST = os.path.join(ROOT, "rama_b200", "host")
ENGINE = os.path.join(HOST, "engine")
MIRROR = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")
SERVICE = os.path.join(ROOT, "tests", "cpp", "service_test")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "rama_b200", "csrc"), "-s"])
    subprocess.check_call(["make", "-C", HOST, "-s"])


def test_host_programs_build_and_link():
    _build()
    assert os.access(ENGINE, os.X_OK) and os.access(MIRROR, os.X_OK) and os.access(SERVICE, os.X_OK)
    r = subprocess.run([ENGINE, "--help"], capture_output=True, text=True)   # no CUDA call before the arguments are valid
    assert r.returncode == 2 and "Usage: engine -m <MODEL> -t <TOKENIZER>" in r.stderr
    r = subprocess.run([ENGINE, "-m", "/nonexistent.bin", "-t", "/nonexistent.bin"], capture_output=True, text=True)
    assert r.returncode == 101 and "panicked" in r.stderr                     # File::open(path).unwrap()


def test_host_mirror_pure_host_logic():
    """Config::from_file, absolute-range views, from_file sizes + wcls alias, bounded channel, get_batch — no CUDA call."""
    _build()
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "host_units_test"), os.path.join(GOLDEN, "ref_shared.bin"),
                        os.path.join(GOLDEN, "ref_untied.bin")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "host units ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_shared", "ref_untied", "ref_hs48"])
@pytest.mark.parametrize("mode", ["fused", "per-op"])
def test_cpp_mirror_forward_matches_reference_goldens(tmp_path, name, mode):
    from oracle import ref
    from rama_b200 import checkpoint as ck
    _build()
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    tokens = g["tokens"].astype("<i4")
    (tmp_path / "tokens.i32").write_bytes(tokens.tobytes())
    out = tmp_path / "out.f32"
    r = subprocess.run([MIRROR, os.path.join(GOLDEN, name + ".bin"), str(tmp_path / "tokens.i32"), str(out), mode],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    n, V = g["logits"].shape
    raw = np.fromfile(out, dtype="<f4")
    logits, samples, kc0 = raw[: n * V].reshape(n, V), raw[n * V: n * V + n], raw[n * V + n:]
    scale = max(1.0, float(np.max(np.abs(g["logits"]))))
    assert float(np.max(np.abs(logits - g["logits"]))) / scale < 1e-4
    assert [int(s) for s in samples] == [int(np.argmax(row)) for row in g["logits"]]
    # layer-0 key cache through Device::to_cpu vs the oracle fed the same tokens
    cfg, tensors = ck.read_checkpoint(os.path.join(GOLDEN, name + ".bin"))
    om = ref.Model(cfg, tensors)
    st = ref.State(om)
    for pos, tok in enumerate(tokens):
        ref.forward(om, st, int(tok), pos)
    want = st.key_cache[: cfg.seq_len * cfg.dim]
    assert kc0.size == want.size
    assert float(np.max(np.abs(kc0 - want))) < 1e-4 * max(1.0, float(np.max(np.abs(want))))


@pytest.mark.gpu
def test_engine_cli_prints_the_oracles_greedy_text(tmp_path):
    """engine -m model.bin -t tokenizer.bin -s N -r 0 ≙ `cargo run --features gpu --bin engine -- ...` (main.rs:61-105)."""
    from oracle import ref
    from rama_b200 import checkpoint as ck
    _build()
    model = os.path.join(GOLDEN, "ref_untied.bin")
    cfg, tensors = ck.read_checkpoint(model)
    pieces = ["<unk>", "<s>", "</s>"] + [f"w{i:02d}." for i in range(3, cfg.vocab_size)]
    tok = tmp_path / "tokenizer.bin"
    with open(tok, "wb") as f:   # llama2.c tokenizer.bin layout (bpe.rs:27-43)
        f.write(struct.pack("<I", max(len(p) for p in pieces)))
        for p in pieces:
            b = p.encode()
            f.write(struct.pack("<fi", 0.0, len(b)) + b)
    steps = cfg.seq_len
    om = ref.Model(cfg, tensors)
    want, _, gap, _ = ref.generate(om, ref.State(om), [], steps, 0.0, 0.9)
    for extra in ([], ["--per-op"]):
        r = subprocess.run([ENGINE, "-m", model, "-t", str(tok), "-s", str(steps), "-r", "0.0"] + extra,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        text, tail = r.stdout.split("\n--------------------------------\n")
        assert tail.startswith("elapsed: ") and "avg tok/s:" in tail
        if all(int(t) >= 3 for t in want):
            got = [int(w[1:]) for w in text.split(".") if w]
            assert got == [int(t) for t in want], (extra, gap)
        else:   # a special piece in the stream: compare the text itself
            assert text == "".join("" if int(t) == 1 else pieces[int(t)] for t in want)


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4])
def test_engine_cli_on_several_gpus_of_one_process(tmp_path, gpus):
    """`engine --gpus N` / `RAMA_GPUS=N engine …`: the one `GPU` handle of main.rs:70-98 becomes a tensor-parallel context over
    N devices of the same process (rama_ctx_create_multi) and prints the oracle's greedy text."""
    import torch
    from oracle import ref
    from rama_b200 import checkpoint as ck
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    _build()
    model = os.path.join(GOLDEN, "ref_untied.bin")
    cfg, tensors = ck.read_checkpoint(model)
    if cfg.n_heads % gpus or cfg.vocab_size % gpus or cfg.hidden_dim % (4 * gpus):
        pytest.skip("golden model not divisible")
    pieces = ["<unk>", "<s>", "</s>"] + [f"w{i:02d}." for i in range(3, cfg.vocab_size)]
    tok = tmp_path / "tokenizer.bin"
    with open(tok, "wb") as f:
        f.write(struct.pack("<I", max(len(p) for p in pieces)))
        for p in pieces:
            b = p.encode()
            f.write(struct.pack("<fi", 0.0, len(b)) + b)
    steps = cfg.seq_len
    want, _, _, _ = ref.generate(ref.Model(cfg, tensors), ref.State(ref.Model(cfg, tensors)), [], steps, 0.0, 0.9)
    expect = "".join("" if int(t) == 1 else pieces[int(t)] for t in want)
    for how in ("flag", "env"):
        cmd = [ENGINE, "-m", model, "-t", str(tok), "-s", str(steps), "-r", "0.0"] + (["--gpus", str(gpus)] if how == "flag" else [])
        env = dict(os.environ, RAMA_GPUS=str(gpus)) if how == "env" else os.environ
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.split("\n--------------------------------\n")[0] == expect, how


def _letter_tokenizer(path, vocab_size):
    """llama2.c tokenizer.bin (bpe.rs:27-43) whose pieces are the specials, the lower-case letters, the space, a few merges and
    fillers: enough for Tokenizer::encode of plain prompts and a printable piece for every id the model may emit."""
    # (not "<unk>" / "</s>": decode() of such a piece panics in the reference, bpe.rs:105-110, and ends that request)
    pieces = ["[unk]", "<s>", "[/s]"] + [chr(ord("a") + i) for i in range(26)] + [" ", "th", "the", "on", "ce", "once", " a"]
    pieces += [f"[{i}]" for i in range(len(pieces), vocab_size)]
    scores = [0.0] * 3 + [-10.0] * 27 + [-1.0, -0.5, -2.0, -2.5, -0.8, -3.0] + [-20.0] * (vocab_size - 36)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", max(len(p) for p in pieces)))
        for p, sc in zip(pieces[:vocab_size], scores):
            b = p.encode()
            f.write(struct.pack("<fi", sc, len(b)) + b)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_shared", "ref_untied"])
@pytest.mark.parametrize("temperature", ["0.0", "0.8"])
def test_engine_service_streams_equal_single_request_generate(tmp_path, name, temperature):
    """EngineService (lib.rs) + get_batch (batcher.rs) in C++ over rama_forward_batch: 9 concurrent clients through 4 slots;
    every stream equals generate() for its prompt, and the steps really were shared."""
    from rama_b200 import checkpoint as ck
    _build()
    model = os.path.join(GOLDEN, name + ".bin")
    cfg, _ = ck.read_checkpoint(model)
    tok = tmp_path / "tokenizer.bin"
    _letter_tokenizer(tok, cfg.vocab_size)
    r = subprocess.run([SERVICE, model, str(tok), "9", "4", str(cfg.seq_len), temperature], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "service ok: 9 requests" in r.stdout
