"""Tensor-parallel arithmetic on the CPU with world_size-2 gloo (no GPU): shards made by the
production shard plan, per-rank partial results from the oracle's ops, all-reduce after wo and w2,
vocabulary-split classifier — must reproduce the unsharded oracle forward."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref
from rama_b200 import checkpoint as ck
from rama_b200.sharding import shard_all, shard_plan, shard_tensor


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _mm(w, x, rows, width):
    o = np.empty(rows, np.float32)
    ref.lib().ref_matmul(ref.fptr(o), ref.fptr(np.ascontiguousarray(w)), ref.fptr(np.ascontiguousarray(x)), width, rows, 1)
    return o


def _allreduce(a):
    t = torch.from_numpy(a.copy()); dist.all_reduce(t); return t.numpy()


def _tp_forward(cfg, sh, st, token, pos, rank, world):
    """One decode step on this rank's shard (mirrors enqueue_step() in csrc/session.cu)."""
    import ctypes as C
    D, F, H, T = cfg.dim, cfg.hidden_dim, cfg.n_heads, cfg.seq_len
    hs, Dq, Fl, Vl, Hl = cfg.head_size, D // world, F // world, cfg.vocab_size // world, H // world
    L = ref.lib()
    x = sh["token_embedding_table"].reshape(-1, D)[token].copy()
    pr = sh["freq_cis_real"][pos * (hs // 2):(pos + 1) * (hs // 2)].copy()
    pi = sh["freq_cis_imag"][pos * (hs // 2):(pos + 1) * (hs // 2)].copy()
    lcfg = ref.CConfig(Dq, Fl, cfg.n_layers, Hl, Hl, cfg.vocab_size, T, int(cfg.shared_weight))  # local heads
    for l in range(cfg.n_layers):
        xb = np.empty(D, np.float32)
        L.ref_rmsnorm(ref.fptr(xb), ref.fptr(x), ref.fptr(sh["rms_att_weight"][l * D:(l + 1) * D].copy()), D)
        q = _mm(sh["wq"].reshape(-1, Dq, D)[l], xb, Dq, D)
        k = _mm(sh["wk"].reshape(-1, Dq, D)[l], xb, Dq, D)
        v = _mm(sh["wv"].reshape(-1, Dq, D)[l], xb, Dq, D)
        for h in range(Hl):
            qh, kh = q[h * hs:(h + 1) * hs], k[h * hs:(h + 1) * hs]
            L.ref_apply_position(qh.ctypes.data_as(C.POINTER(C.c_float)), kh.ctypes.data_as(C.POINTER(C.c_float)),
                                 ref.fptr(pr), ref.fptr(pi), hs)
        st["kc"][l, pos], st["vc"][l, pos] = k, v
        att, xo = np.zeros(Hl * T, np.float32), np.zeros(Dq, np.float32)
        L.ref_multi_head_attention(ref.fptr(xo), ref.fptr(att), ref.fptr(q), ref.fptr(st["kc"].reshape(-1)),
                                   ref.fptr(st["vc"].reshape(-1)), C.byref(lcfg), l, pos)
        x = x + _allreduce(_mm(sh["wo"].reshape(-1, D, Dq)[l], xo, D, Dq))          # row-parallel + all-reduce
        L.ref_rmsnorm(ref.fptr(xb), ref.fptr(x), ref.fptr(sh["rms_ffn_weight"][l * D:(l + 1) * D].copy()), D)
        h1 = _mm(sh["w1"].reshape(-1, Fl, D)[l], xb, Fl, D)
        h3 = _mm(sh["w3"].reshape(-1, Fl, D)[l], xb, Fl, D)
        L.ref_sinu(ref.fptr(h1), Fl)
        x = x + _allreduce(_mm(sh["w2"].reshape(-1, D, Fl)[l], h1 * h3, D, Fl))
    xn = np.empty(D, np.float32)
    L.ref_rmsnorm(ref.fptr(xn), ref.fptr(x), ref.fptr(sh["rms_final_weight"].copy()), D)
    wcls = sh["wcls"].reshape(Vl, D) if not cfg.shared_weight else sh["token_embedding_table"].reshape(-1, D)[rank * Vl:(rank + 1) * Vl]
    mine = torch.from_numpy(_mm(wcls, xn, Vl, D))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts).numpy()


def _worker(rank, world, port, name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref.lib().ref_set_threads(1)
        cfg = ck.CONFIGS[name]
        tensors = ck.synth_tensors(cfg, ck.SynthSpec(seed=21, rms_jitter=0.1))
        sh = shard_all(cfg, tensors, rank, world)
        Dq = cfg.dim // world
        st = {"kc": np.zeros((cfg.n_layers, cfg.seq_len, Dq), np.float32), "vc": np.zeros((cfg.n_layers, cfg.seq_len, Dq), np.float32)}
        om = ref.Model(cfg, tensors); os_ = ref.State(om)
        worst = 0.0
        for pos, tok in enumerate([1, 9, 200, 3, 77, 5]):
            ref.forward(om, os_, tok, pos)
            got = _tp_forward(cfg, sh, st, tok, pos, rank, world)
            worst = max(worst, float(np.max(np.abs(got - os_.logits))))
            assert int(np.argmax(got)) == int(np.argmax(os_.logits))
        q.put((rank, worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny", "tiny-sep"])
def test_tp2_gloo_matches_unsharded_oracle(name):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert sorted(r for r, _ in res) == [0, 1]
    assert max(w for _, w in res) < 1e-4  # only the association of the k-sum across ranks differs


def test_shard_plan_covers_every_element_once():
    cfg = ck.CONFIGS["tiny-sep"]
    tensors = ck.synth_tensors(cfg, ck.SynthSpec(seed=2))
    for world in (1, 2):
        for name in ("wq", "wo", "w1", "w2", "wcls"):
            Lc, R, C = shard_plan(cfg, 0, world)[name][:3]
            full = tensors[name].reshape(Lc, R, C)
            parts = [shard_tensor(cfg, name, tensors[name], r, world) for r in range(world)]
            p0 = shard_plan(cfg, 0, world)[name]
            axis = 2 if p0[6] != C else 1
            re = np.concatenate([p.reshape(Lc, shard_plan(cfg, r, world)[name][4], shard_plan(cfg, r, world)[name][6])
                                 for r, p in enumerate(parts)], axis=axis)
            assert np.array_equal(re, full), (name, world)
    with pytest.raises(ValueError):
        shard_plan(cfg, 0, 3)
