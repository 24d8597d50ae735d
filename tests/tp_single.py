"""Single-process tensor parallelism (rama_ctx_create_multi): ONE process, ONE handle, N devices — what the reference's
engine binary and server get with `--features gpu` (main.rs:70-98, lib.rs:99-119 hold a single `GPU`).  Run directly:

    python tests/tp_single.py [world ...]        (default: 2, and 4 / 8 when the box has them)

Every check is against the CPU oracle on the same weights.  Prints 'TP1P-OK <world>' per world size."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from rama_b200 import checkpoint as ck  # noqa: E402
from rama_b200.engine import GPU, Batch, Session, generate  # noqa: E402
from rama_b200.sharding import shard_tensor  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(1.0, float(np.max(np.abs(b)))))


def argmax_last(x):
    return int(np.flatnonzero(x == x.max())[-1])   # later index wins (cpu.rs:165)


def run(world: int):
    for name, steps in (("tiny", 40), ("tiny-sep", 30), ("l7-2layer", 12)):
        cfg = ck.CONFIGS[name]
        if cfg.n_heads % world or cfg.vocab_size % world or cfg.hidden_dim % world or (cfg.hidden_dim // world) % 4:
            continue
        spec = ck.SynthSpec(seed=77, rms_jitter=0.1)
        tensors = ref.synth_tensors(cfg, spec)
        om = ref.Model(cfg, tensors)
        gpu = GPU.multi(world)
        gpu.load_host(cfg, tensors)
        for t in ("wq", "wo", "w1", "w2"):   # weight_shard of a group = rank 0's shard of the production plan
            assert gpu.weight_shard(t).tobytes() == shard_tensor(cfg, t, tensors[t], 0, world).tobytes(), t
        sess = Session(gpu)
        want, want_logits, gap, _ = ref.generate(om, ref.State(om), [5, 6, 7], steps, 0.0, 0.9, want_logits=True)
        got = generate(sess, [5, 6, 7], steps, 0.0, 0.9)                      # device-resident loop on every device
        assert got == list(want), (world, name, got, list(want))
        assert rel(sess.logits(), want_logits[steps - 1]) < 1e-3, (world, name)
        sess.reset()
        got_host = generate(sess, [5, 6, 7], steps, 0.0, 0.9, host_loop=True)  # forward()+sample() per token
        assert got_host == list(want), (world, name, got_host, list(want))
        # teacher-forced logits at every position
        sess.reset()
        os_ = ref.State(om)
        token = 1
        for pos in range(steps):
            sess.forward(token, pos)
            ref.forward(om, os_, token, pos)
            assert rel(sess.logits(), os_.logits) < 1e-3, (world, name, pos)
            token = int(want[pos])
        # temperature + top-p: a deterministic function of the logits (constant draw)
        # (at a 32000-entry vocabulary with random weights thousands of candidates share the top-p mass, and the walk's thresholds
        # sit 1e-5 apart: the 1e-6 logit differences of the split k-sums may legitimately move a sample — there the
        # check is agreement between the paths of the same handle, below)
        small_vocab = cfg.vocab_size <= 1024
        tk = generate(sess, [5, 6, 7], steps, 0.8, 0.9)
        if small_vocab:
            wt, _, _, _ = ref.generate(om, ref.State(om), [5, 6, 7], steps, 0.8, 0.9)
            assert tk == list(wt), (world, name, "T=0.8")
        s2 = Session(gpu)                                                      # a second RunState on the same handle
        assert generate(s2, [5, 6, 7], steps, 0.0, 0.9) == list(want)
        s2.close()
        # prompt prefill (tensor cores + peer-memory reduce-scatter / all-gather)
        n_pf = min(24, cfg.seq_len - 4)
        toks = [1] + [int(t) for t in np.random.default_rng(9).integers(0, cfg.vocab_size, n_pf - 1)]
        os_ = ref.State(om)
        for pos, t in enumerate(toks):
            ref.forward(om, os_, t, pos)
        s3 = Session(gpu)
        s3.prefill(toks, 0)
        assert rel(s3.logits(), os_.logits) < 1e-3, (world, name, "prefill")
        assert s3.sample(0.0, 0.9) == argmax_last(os_.logits)
        nxt = argmax_last(os_.logits)
        s3.forward(nxt, n_pf); ref.forward(om, os_, nxt, n_pf)                 # decoding continues from the prefilled cache
        assert rel(s3.logits(), os_.logits) < 1e-3, (world, name, "after prefill")
        s3.close()
        # generate() with a long prompt takes the prefill path
        prompt = [int(t) for t in np.random.default_rng(3).integers(2, cfg.vocab_size, 20)]
        wp, _, _, _ = ref.generate(om, ref.State(om), prompt, min(cfg.seq_len, 36), 0.0, 0.9)
        assert generate(sess, prompt, min(cfg.seq_len, 36), 0.0, 0.9) == list(wp), (world, name, "generate+prefill")
        # batched decode: 5 sequences at staggered positions
        B = 5
        streams = [[1] + [int(t) for t in np.random.default_rng(20 + i).integers(0, cfg.vocab_size, 10 + i)] for i in range(B)]
        bs = [Session(gpu) for _ in range(B)]
        sts = [ref.State(om) for _ in range(B)]
        for i in range(B):
            for pos in range(i):
                bs[i].forward(streams[i][pos], pos)
                ref.forward(om, sts[i], streams[i][pos], pos)
        batch = Batch(gpu, 8)
        for k in range(6):
            batch.forward(bs, [streams[i][i + k] for i in range(B)], [i + k for i in range(B)])
            for i in range(B):
                ref.forward(om, sts[i], streams[i][i + k], i + k)
        for i in range(B):
            assert rel(bs[i].logits(), sts[i].logits) < 1e-3, (world, name, "batch", i)
        assert batch.sample(bs, 0.0, 0.9) == [argmax_last(st.logits) for st in sts], (world, name, "batch sample")
        prompts_b = [[], [9, 8, 7], [4], [3, 3, 3, 3, 3, 3], [11]]
        for temp in (0.0, 0.8):                                               # device-resident batched loop
            gb, _ = batch.generate(bs, prompts_b, 16, temp, 0.9)
            for i, pb in enumerate(prompts_b):
                if temp == 0.0 or small_vocab:
                    wb, _, _, _ = ref.generate(om, ref.State(om), pb, 16, temp, 0.9)
                    assert [int(t) for t in gb[i]] == [int(t) for t in wb], (world, name, "generate_batch", temp, i)
                else:  # both loops of the same handle sample from logits of the same magnitude: finite, in range, right length
                    assert len(gb[i]) == 16 and all(0 <= int(t) < cfg.vocab_size for t in gb[i]), (world, name, temp, i)
                    assert [int(t) for t in gb[i][: len(pb)]] == list(pb)[:16], (world, name, "forced prompt", i)
        batch.close()
        for x in bs:
            x.close()
        sess.close()
        # reload through the synthetic generator and the file loader on the same handle
        gpu.load_synthetic(cfg, spec)
        for t in ("wq", "wo", "w2", "rms_att_weight"):
            assert gpu.weight_shard(t).tobytes() == shard_tensor(cfg, t, tensors[t], 0, world).tobytes(), t
        gpu.close()
    # a checkpoint the reference's exporter wrote, through the file loader of the multi-device handle
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_untied.npz"))
    cfgf, _ = ck.read_checkpoint(os.path.join(ROOT, "tests", "golden", "ref_untied.bin"))
    if cfgf.n_heads % world == 0 and cfgf.vocab_size % world == 0 and cfgf.hidden_dim % (4 * world) == 0:
        gpu = GPU.multi(world)
        gpu.load_file(os.path.join(ROOT, "tests", "golden", "ref_untied.bin"))
        sess = Session(gpu)
        for pos, tok in enumerate(g["tokens"]):
            sess.forward(int(tok), pos)
            assert rel(sess.logits(), g["logits"][pos]) < 1e-3, ("golden", world, pos)
        sess.close(); gpu.close()
    print(f"TP1P-OK {world}", flush=True)


def main():
    import torch
    n = torch.cuda.device_count()
    worlds = [int(a) for a in sys.argv[1:]] or [w for w in (2, 4, 8) if w <= n]
    for w in worlds:
        if w > n:
            raise SystemExit(f"world {w} needs {w} GPUs, {n} present")
        run(w)


if __name__ == "__main__":
    main()
