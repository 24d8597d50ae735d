"""tcgen05 3xTF32 GEMM (rama_b200/csrc/gemm_tf32x3.cuh) vs float64 truth: f32-level accuracy is the
contract (the prefill / batched-decode paths must hold the 1e-3 logit tolerance of the decode path)."""
import numpy as np
import pytest

from rama_b200.engine import GPU, DeviceBuffer, View

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    g = GPU(0)
    yield g
    g.close()


def dev(gpu, a):
    return View(DeviceBuffer(gpu, a.size, a))


def run(gpu, M, N, K, variant, flags, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64).T
    o = dev(gpu, np.full(M * N, np.nan, np.float32))
    gpu.matmul_nt(o, dev(gpu, a), dev(gpu, b), M, N, K, variant, flags)
    got = o.data.to_host()
    got = got.reshape(N, M).T if flags & 2 else got.reshape(M, N)
    f32 = (a @ b.T).astype(np.float64)  # numpy f32 GEMM: the accuracy class we must stay in
    err = float(np.max(np.abs(got - want)))
    ref_err = float(np.max(np.abs(f32 - want)))
    return err, ref_err


SHAPES = [(128, 256, 64), (128, 256, 4096), (512, 4096, 4096), (512, 1024, 11008), (100, 300, 288),
          (1, 32, 16), (130, 520, 1000), (512, 768, 768), (512, 11008, 1024)]  # the last one: more tiles than SMs


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])  # 4, 5: CTA pair (tcgen05.mma.cta_group::2, 256×128 tile over two SMs)
@pytest.mark.parametrize("shape", SHAPES)
def test_matmul_nt_f32_accuracy(gpu, shape, variant):
    M, N, K = shape
    err, ref_err = run(gpu, M, N, K, variant, 0)
    # 3xTF32 drops the lo·lo term (2^-22 relative per product): allow a few times the f32 GEMM error
    assert err < max(8 * ref_err, 6e-6), (err, ref_err)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])  # 6-9: pre-split B operand (the batched-decode tiles of round 2)
@pytest.mark.parametrize("shape", [(4096, 64, 4096), (11008, 64, 4096), (4096, 64, 11008), (288, 3, 288), (1000, 17, 64)])
def test_matmul_nt_transposed_store(gpu, shape, variant):
    M, N, K = shape
    err, ref_err = run(gpu, M, N, K, variant, 2)
    assert err < max(8 * ref_err, 6e-6), (err, ref_err)


@pytest.mark.parametrize("variant", [4, 6, 8])
@pytest.mark.parametrize("ksplit", [2, 5])
def test_matmul_nt_transposed_split_k(gpu, variant, ksplit):
    """split-K partials summed in fixed order (flags bits 8..15 = split factor), uneven k-block counts per split"""
    err, ref_err = run(gpu, 4096, 64, 4096 + 96, variant, 2 | (ksplit << 8))
    assert err < max(8 * ref_err, 6e-6), (err, ref_err)


def test_matmul_nt_hi_round_mode(gpu):
    """flags bit 0: the B tile's hi half is rewritten rounded-to-nearest instead of relying on the tensor
    core ignoring the 13 low mantissa bits of the raw tile (the default).  Both must be f32-accurate."""
    err, ref_err = run(gpu, 256, 512, 2048, 0, 1)
    err0, _ = run(gpu, 256, 512, 2048, 0, 0)
    print("hi_round err", err, "raw-hi err", err0, "f32 gemm err", ref_err)
    assert err < max(8 * ref_err, 2e-6) and err0 < max(8 * ref_err, 2e-6)
