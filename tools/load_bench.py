#!/usr/bin/env python
"""Checkpoint loader throughput: writes a synthetic llama2.c v0 .bin (default: the 2-layer 7B geometry, 2.7 GB) to a
scratch directory, loads it with rama_ctx_load_file (pread → pinned ring → HBM) twice (first: page cache as the write
left it; second: warm) and checks the weights arrived bit-exact.  load_bench.py [model] [dir]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rama_b200 import _lib, checkpoint as ck
from rama_b200.engine import GPU

model = sys.argv[1] if len(sys.argv) > 1 else "l7-2layer"
d = sys.argv[2] if len(sys.argv) > 2 else "/tmp"
cfg = ck.CONFIGS[model]
spec = ck.SynthSpec(seed=5)
path = os.path.join(d, f"rama_{model}.bin")
t0 = time.time()
tensors = ck.synth_tensors(cfg, spec)
ck.write_checkpoint(path, cfg, tensors)
print(f"wrote {os.path.getsize(path) / 1e9:.2f} GB in {time.time() - t0:.1f} s", flush=True)
out = []
for attempt in range(2):
    gpu = GPU(0)
    t0 = time.time()
    gpu.load_file(path)
    dt = time.time() - t0
    g = C.c_double()
    _lib.check(_lib.lib().rama_last_load_gbps(C.byref(g)))
    ok = all(gpu.weight_shard(n).tobytes() == np.ascontiguousarray(tensors[n], np.float32).tobytes()
             for n in ("wq", "wo", "w2", "token_embedding_table", "rms_final_weight"))
    out.append({"attempt": attempt, "wall_s": round(dt, 3), "pipeline_gbps": round(g.value, 2), "bit_exact": ok})
    print(json.dumps(out[-1]), flush=True)
    gpu.close()
os.remove(path)
