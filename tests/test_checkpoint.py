import os

import numpy as np
import pytest

from oracle import ref
from rama_b200 import checkpoint as ck


def test_byte_counts_match_survey():
    # SURVEY.md §8(d) / BASELINE.md §2
    assert ck.CONFIGS["stories15M"].weight_bytes_per_token() == 60_768_192
    assert ck.CONFIGS["stories110M"].weight_bytes_per_token() == 438_122_752
    assert ck.CONFIGS["llama2-7B"].weight_bytes_per_token() == 26_429_391_360
    assert ck.CONFIGS["stories15M"].avg_bytes_per_token(256) == 62_558_400
    assert ck.CONFIGS["llama2-7B"].avg_bytes_per_token(256) == 26_565_181_952
    assert ck.CONFIGS["llama2-7B"].file_bytes() == 26_954_711_068


def test_header_roundtrip_sign_convention():
    for cfg in ck.CONFIGS.values():
        assert ck.Config.from_header(cfg.header()) == cfg
    import struct
    assert struct.unpack("<7i", ck.CONFIGS["llama2-7B"].header())[5] == -32000  # separate wcls
    assert struct.unpack("<7i", ck.CONFIGS["stories15M"].header())[5] == 32000  # shared


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
def test_synth_numpy_equals_cpp_bitwise(name):
    cfg = ck.CONFIGS[name]
    spec = ck.SynthSpec(seed=7, rms_jitter=0.1)
    a, b = ck.synth_tensors(cfg, spec), ref.synth_tensors(cfg, spec)
    for k in ck.TENSORS:
        assert a[k].tobytes() == b[k].tobytes(), k
    assert abs(float(a["wq"].std()) - cfg.dim ** -0.5) < 0.02 * cfg.dim ** -0.5
    assert abs(float(a["rms_att_weight"].mean()) - 1.0) < 0.05


def test_synth_chunk_offsets_consistent():
    s = ck.synth_scale(0.5)
    full = ck.synth_fill(10_000, 3, 2, s)
    part = ck.synth_fill(1_000, 3, 2, s, start=4_321)
    assert part.tobytes() == full[4_321:5_321].tobytes()
    assert ck.synth_fill(100, 3, 2, s).tobytes() != ck.synth_fill(100, 3, 5, s).tobytes()
    assert ck.synth_fill(100, 3, 2, s).tobytes() != ck.synth_fill(100, 4, 2, s).tobytes()


def test_write_read_roundtrip(tmp_path):
    cfg = ck.CONFIGS["tiny-sep"]
    t = ck.synth_tensors(cfg, ck.SynthSpec(seed=5))
    p = str(tmp_path / "m.bin")
    ck.write_checkpoint(p, cfg, t)
    assert os.path.getsize(p) == cfg.file_bytes()
    cfg2, t2 = ck.read_checkpoint(p)
    assert cfg2 == cfg
    for k in ck.TENSORS:
        assert np.array_equal(t[k], t2[k]), k
    # oracle's mmap loader sees the same tensors and aliases wcls only when shared
    m = ref.FileModel(p)
    assert m.cfg == cfg
