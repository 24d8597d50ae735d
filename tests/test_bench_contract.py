"""bench.py contract (driver side): the reference arm runs on the CPU and prints ONE JSON line with the agreed keys; the CUDA arm
refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _json_lines(out: str):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--model", "stories15M", "--steps", "1", "--warmup", "0",
                        "--cpu-tokens", "24"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert d["value"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "stories15M" in d["config"]["workload"]


def test_reference_arm_under_torchrun_uses_all_cores_and_stays_bounded():
    """torchrun exports OMP_NUM_THREADS=1 when nproc > 1 (round 1: the oracle then ran single-threaded and the arm timed
    out at N = 2/4/8).  Rank 0 must still use every host core, size its sample from the time budget and print the same
    `config` keys as the CUDA arm."""
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--model", "stories15M", "--steps", "3",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    (d,) = _json_lines(r.stdout)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2 and d["steps"] == 3 and d["warmup"] == 1
    assert sorted(d["config"]) == ["l2", "parallelism", "seed", "workload"] and d["config"]["parallelism"] == "tp2"
    assert d["ms_per_step"] * (d["steps"] + d["warmup"]) < 60e3


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--model", "stories15M"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and _json_lines(r.stdout) == []


def test_cuda_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return  # (on the GPU box this arm is what the driver runs)
    r = subprocess.run([sys.executable, BENCH, "--model", "stories15M", "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and _json_lines(r.stdout) == []
    assert "no CPU path" in (r.stderr + r.stdout)
