"""Device::sample parity (cpu.rs:155-179, infer.rs:55-85) — sampling runs on the device."""
import numpy as np
import pytest

from oracle import ref
from rama_b200 import _lib, checkpoint as ck
from rama_b200.engine import GPU, DeviceBuffer, View, Session, generate
from util import model_tensors, rand

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    g = GPU(0)
    yield g
    g.close()


class _RS:
    pass


def _dev_sample(gpu, logits, temperature, topp):
    rs = _RS()
    rs.logits = View(DeviceBuffer(gpu, logits.size, logits))
    cfg = ck.Config(4, 4, 1, 1, 1, logits.size, 1, True)
    nxt = gpu.sample(cfg, rs, temperature, topp)
    return nxt, rs.logits.data.to_host()


def _ref_sample(logits, temperature, topp):
    x = logits.copy()
    return int(ref.lib().ref_sample(ref.fptr(x), x.size, temperature, topp)), x


def test_greedy_ties_go_to_the_highest_index(gpu):
    x = np.zeros(32000, np.float32)
    x[[5, 777, 31999]] = 3.5
    assert _dev_sample(gpu, x, 0.0, 0.9)[0] == 31999 == _ref_sample(x, 0.0, 0.9)[0]
    x[31999] = 3.4999
    assert _dev_sample(gpu, x, 0.0, 0.9)[0] == 777 == _ref_sample(x, 0.0, 0.9)[0]
    c = np.full(100, -2.0, np.float32)  # all equal → last index
    assert _dev_sample(gpu, c, 0.0, 0.9)[0] == 99 == _ref_sample(c, 0.0, 0.9)[0]


@pytest.mark.parametrize("V", [2, 64, 300, 4097, 32000])
@pytest.mark.parametrize("temperature,topp,scale", [(1.0, 0.9, 1.0), (0.5, 0.9, 1.0), (2.0, 0.5, 1.0),
                                                     (0.8, 0.95, 6.0), (1.0, 0.1, 6.0), (1.0, 1.0, 1.0)])
def test_top_p_matches_oracle(gpu, V, temperature, topp, scale):
    agree = 0
    for seed in range(4):
        x = rand(V, 1000 + seed, scale)
        want, wprobs = _ref_sample(x, temperature, topp)
        if want < 0:  # no p above the cutoff (e.g. V=2, topp=0.1 ⇒ cutoff 0.9): the reference panics
            with pytest.raises(_lib.RamaError):
                _dev_sample(gpu, x, temperature, topp)
            agree += 1
            continue
        got, gprobs = _dev_sample(gpu, x, temperature, topp)
        # in place like the reference: logits now hold probabilities (the oracle's sequential f32 sum
        # over V terms carries ~1e-4 relative error at V=32000; rayon's sum in the reference differs again)
        np.testing.assert_allclose(gprobs, wprobs, rtol=3e-4 if V > 1000 else 2e-5, atol=1e-10)
        agree += int(got == want)
        if got != want:
            # the only legitimate difference: the f32 softmax sum is associated differently, which may
            # move a cumulative-sum boundary between two adjacent candidates — the picks must then be
            # neighbours in probability
            assert abs(float(wprobs[got]) - float(wprobs[want])) <= 1e-4 * float(wprobs.max())
    assert agree >= 3


def test_top_p_peaked_distribution_small_candidate_set(gpu):
    x = rand(32000, 7, 0.1)
    x[[11, 222, 3333]] = [9.0, 8.5, 8.0]
    want, _ = _ref_sample(x, 1.0, 0.9)
    assert _dev_sample(gpu, x, 1.0, 0.9)[0] == want


def test_empty_candidate_list_is_an_error_like_the_reference_panic(gpu):
    x = rand(64, 3)
    assert _ref_sample(x, 1.0, -1e9)[0] == -1       # oracle: reference would panic (infer.rs:66)
    with pytest.raises(_lib.RamaError):
        _dev_sample(gpu, x, 1.0, -1e9)


def test_generate_with_temperature_matches_oracle_stream():
    """T>0 is a deterministic function of the logits (constant RNG draw)."""
    cfg, spec, tensors = model_tensors("tiny")
    gpu = GPU(0)
    gpu.load_host(cfg, tensors)
    om = ref.Model(cfg, tensors)
    sess = Session(gpu)
    for temperature, topp in [(1.0, 0.9), (0.7, 0.8)]:
        want, _, _, _ = ref.generate(om, ref.State(om), [5, 6], 40, temperature, topp)
        got = generate(sess, [5, 6], 40, temperature, topp)
        got_host = generate(sess, [5, 6], 40, temperature, topp, host_loop=True)
        assert got == got_host
        # identical unless a cumulative-sum boundary flips (then the streams diverge): require a long common prefix
        common = next((i for i in range(40) if got[i] != want[i]), 40)
        assert common >= 20, (temperature, topp, common)
    sess.close(); gpu.close()
