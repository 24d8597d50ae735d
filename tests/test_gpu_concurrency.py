"""Threading contract of the boundary: the reference server keeps ONE device + weights in a static and
runs many request tasks concurrently, each with its own RunState (lib.rs:56,133-153).  Here: one GPU context, several
host threads, each driving its own Session through forward()/sample()/generate() at the same time — results must equal
the serial ones."""
import threading

import numpy as np
import pytest

from rama_b200.engine import GPU, Session, generate
from util import model_tensors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
def test_concurrent_sessions_on_one_context(name):
    cfg, spec, tensors = model_tensors(name)
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    rng = np.random.default_rng(2)
    jobs = [([int(t) for t in rng.integers(2, cfg.vocab_size, int(rng.integers(0, 6)))], float(tmp))
            for tmp in (0.0, 0.0, 0.8, 0.0, 0.9, 0.0, 0.0, 0.7)]
    serial = []
    s = Session(gpu)
    for prompt, temp in jobs:
        serial.append(generate(s, prompt, cfg.seq_len, temp, 0.9, host_loop=True))
    s.close()
    out = [None] * len(jobs)
    errs = []

    def work(i):
        try:
            sess = Session(gpu)                       # RunState::from_state per request task
            for rep in range(3):                      # three requests per thread, device loop and host loop
                out[i] = generate(sess, jobs[i][0], cfg.seq_len, jobs[i][1], 0.9, host_loop=(rep % 2 == 0))
                assert out[i] == serial[i], (i, rep)
            sess.close()
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert out == serial
    gpu.close()
