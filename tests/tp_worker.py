"""torchrun worker for the multi-GPU parity test (one process per GPU, NCCL inside the library).
Rank 0 prints 'TP-OK' on success; any rank raising makes torchrun exit non-zero."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from rama_b200 import checkpoint as ck  # noqa: E402
from rama_b200.engine import GPU, Batch, Session, generate  # noqa: E402
from rama_b200.sharding import shard_tensor  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only to hand out the NCCL id; the data path is the library's NCCL
    cases = [(m, n, st) for m in ("p2p", "nccl", "p2p+persistent", "p2p+reduce0", "p2p+reduce1", "p2p+reduce2")
             for n, st in (("tiny", 40), ("tiny-sep", 30), ("l7-2layer", 12))]
    for mode, name, steps in cases:
        # read at rama_ctx_create: fused peer-memory exchange vs NCCL collectives; one kernel per op group vs the
        # persistent cooperative step kernel
        os.environ["RAMA_TP_COMM"] = mode.split("+")[0]
        os.environ["RAMA_STEP"] = "persistent" if mode.endswith("persistent") else "kernels"
        # how the norm prologues reduce the peer partials: every CTA reads all / cluster-shared / two-phase via a local LL buffer
        os.environ["RAMA_TP_REDUCE"] = mode[-1] if "reduce" in mode else "-1"
        cfg = ck.CONFIGS[name]
        if cfg.n_heads % world:
            continue
        spec = ck.SynthSpec(seed=77, rms_jitter=0.1)
        tensors = ref.synth_tensors(cfg, spec)
        ids = [GPU.unique_id() if rank == 0 else None]  # one ncclUniqueId per communicator
        dist.broadcast_object_list(ids, src=0)
        gpu = GPU(local, (rank, world, ids[0]))
        gpu.load_host(cfg, tensors)
        for t in ("wq", "wo", "w1", "w2", "wcls", "token_embedding_table"):  # shards = production plan
            assert gpu.weight_shard(t).tobytes() == shard_tensor(cfg, t, tensors[t], rank, world).tobytes(), t
        sess = Session(gpu)
        om = ref.Model(cfg, tensors)
        want, want_logits, gap, _ = ref.generate(om, ref.State(om), [5, 6, 7], steps, 0.0, 0.9, want_logits=True)
        got = generate(sess, [5, 6, 7], steps, 0.0, 0.9)
        assert got == list(want), (mode, name, rank, got, list(want))
        sess.reset()
        got_host = generate(sess, [5, 6, 7], steps, 0.0, 0.9, host_loop=True)
        assert got_host == list(want), (mode, name, rank, got_host, list(want))
        lg = sess.logits()  # all-gathered across ranks
        err = float(np.max(np.abs(lg - want_logits[steps - 1])) / max(1.0, float(np.max(np.abs(want_logits[steps - 1])))))
        assert err < 1e-3, (mode, name, rank, err)
        # temperature sampling needs the gathered logits on every rank; all ranks must agree
        tk = generate(sess, [5, 6, 7], steps, 0.8, 0.9)
        tk_all = [None] * world
        dist.all_gather_object(tk_all, tk)
        assert all(t == tk_all[0] for t in tk_all), (mode, name)
        # a second session on the same context (server model: one RunState per request)
        s2 = Session(gpu)
        assert generate(s2, [5, 6, 7], steps, 0.0, 0.9) == list(want)
        s2.close()
        sess.close()
        if mode == "p2p":
            # prompt prefill under TP (NCCL all-reduce of the wo / w2 outputs): last logits and continued decoding
            n_pf = min(24, cfg.seq_len - 4)
            toks = [1] + [int(t) for t in np.random.default_rng(9).integers(0, cfg.vocab_size, n_pf - 1)]
            os_ = ref.State(om)
            for pos, t in enumerate(toks):
                ref.forward(om, os_, t, pos)
            s3 = Session(gpu)
            s3.prefill(toks, 0)
            lg3 = s3.logits()
            e3 = float(np.max(np.abs(lg3 - os_.logits)) / max(1.0, float(np.max(np.abs(os_.logits)))))
            assert e3 < 1e-3, ("prefill", name, rank, e3)
            nxt = s3.sample(0.0, 0.9)
            assert nxt == int(np.flatnonzero(os_.logits == os_.logits.max())[-1]), ("prefill sample", name, rank)
            s3.close()
            # batched decode under TP: 3 sequences at staggered positions vs the oracle
            B = 3
            streams = [[1] + [int(t) for t in np.random.default_rng(20 + i).integers(0, cfg.vocab_size, 8 + i)] for i in range(B)]
            bs = [Session(gpu) for _ in range(B)]
            sts = [ref.State(om) for _ in range(B)]
            for i in range(B):  # sequence i has already decoded i tokens through the batch-1 path
                for pos in range(i):
                    bs[i].forward(streams[i][pos], pos)
                    ref.forward(om, sts[i], streams[i][pos], pos)
                bs[i].sync()
            batch = Batch(gpu, 8)
            for k in range(6):
                batch.forward(bs, [streams[i][i + k] for i in range(B)], [i + k for i in range(B)])
                for i in range(B):
                    ref.forward(om, sts[i], streams[i][i + k], i + k)
            batch.sync()
            for i in range(B):
                lb = bs[i].logits()
                eb = float(np.max(np.abs(lb - sts[i].logits)) / max(1.0, float(np.max(np.abs(sts[i].logits)))))
                assert eb < 1e-3, ("batch", name, rank, i, eb)
            nb = batch.sample(bs, 0.0, 0.9)
            assert nb == [int(np.flatnonzero(st.logits == st.logits.max())[-1]) for st in sts], ("batch sample", name, rank)
            # device-resident batched loop under TP
            prompts_b = [[], [9, 8, 7], [4]]
            gb, _ = batch.generate(bs, prompts_b, 16, 0.0, 0.9)
            for i, pb in enumerate(prompts_b):
                wb, _, _, _ = ref.generate(om, ref.State(om), pb, 16, 0.0, 0.9)
                assert [int(t) for t in gb[i]] == [int(t) for t in wb], ("generate_batch", name, rank, i)
            batch.close()
            for x in bs:
                x.close()
        # synthetic loader generates each rank's shard in place: same bits as the host recipe
        gpu.load_synthetic(cfg, spec)
        for t in ("wq", "wo", "w2", "rms_att_weight"):
            assert gpu.weight_shard(t).tobytes() == shard_tensor(cfg, t, tensors[t], rank, world).tobytes(), t
        dist.barrier()
        gpu.close()
        dist.barrier()
    if rank == 0:
        print("TP-OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
