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
from rama_b200.engine import GPU, Session, generate  # noqa: E402
from rama_b200.sharding import shard_tensor  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only to hand out the NCCL id; the data path is the library's NCCL
    cases = [(m, n, st) for m in ("p2p", "nccl") for n, st in (("tiny", 40), ("tiny-sep", 30), ("l7-2layer", 12))]
    for mode, name, steps in cases:
        os.environ["RAMA_TP_COMM"] = mode  # read at rama_ctx_create: fused peer-memory exchange vs NCCL collectives
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
        assert got_host == list(want), (name, rank)
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
