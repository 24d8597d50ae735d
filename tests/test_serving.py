"""Request batching (rama_b200/serving.py): batcher.rs semantics on the host, and — on a GPU — continuous batching
through rama_forward_batch producing, per request, exactly the token stream generate() produces."""
import queue
import threading
import time

import numpy as np
import pytest

from rama_b200.serving import get_batch


def test_get_batch_fills_up_to_batch_size():
    q = queue.Queue()
    for i in range(5):
        q.put(f"p{i}")
    got = []
    get_batch(q, got, 3, 1.0)          # batcher.rs:33: stops at batch_size
    assert got == ["p0", "p1", "p2"]
    get_batch(q, got, 8, 0.05)         # batcher.rs:14: timeout ends the wait with what arrived
    assert got == ["p0", "p1", "p2", "p3", "p4"]


def test_get_batch_waits_for_late_arrivals_until_timeout():
    q = queue.Queue()
    threading.Timer(0.05, lambda: q.put("late")).start()
    got = []
    t0 = time.monotonic()
    get_batch(q, got, 2, 0.3)
    assert got == ["late"] and 0.25 <= time.monotonic() - t0 < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
@pytest.mark.parametrize("temperature", [0.0, 0.8])
def test_continuous_batching_equals_per_request_generate(name, temperature):
    from rama_b200.engine import GPU, Session, generate
    from rama_b200.serving import BatchedEngine
    from util import model_tensors
    cfg, spec, tensors = model_tensors(name)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    rng = np.random.default_rng(4)
    reqs = []
    for i in range(9):
        n_prompt = int(rng.integers(0, 12))
        prompt = [int(t) for t in rng.integers(2, cfg.vocab_size, n_prompt)]
        steps = int(rng.integers(1, cfg.seq_len + 1))
        reqs.append((prompt, steps))
    eng = BatchedEngine(gpu, max_batch=4, temperature=temperature, topp=0.9)   # 9 requests through 4 slots
    events = []
    ids = []
    for k, (prompt, steps) in enumerate(reqs):
        ids.append(eng.submit(prompt, steps, on_token=lambda rid, t: events.append((rid, t))))
        if k % 2 == 1:
            eng.step()                 # requests keep arriving while others are mid-generation
    eng.run_until_idle()
    ref_s = Session(gpu)
    ref_s.set_prefill(0)
    for rid, (prompt, steps) in zip(ids, reqs):
        want = generate(ref_s, prompt, steps, temperature, 0.9)
        got = eng.finished[rid].tokens
        assert got == want, (rid, len(prompt), steps)
        assert [t for r, t in events if r == rid] == want
    with pytest.raises(ValueError):
        eng.submit([5], cfg.seq_len + 1)
    ref_s.close(); eng.close(); gpu.close()
