python -m pytest tests/test_gpu_prefill.py tests/test_gpu_batch.py tests/test_serving.py -m gpu -x -q 2>&1 | tail -2
for p in 1 0 1 0; do
RAMA_PREFILL_PDL=$p python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from rama_b200 import checkpoint as ck
from rama_b200.engine import GPU, Session
cfg = ck.CONFIGS["llama2-7B"]
gpu = GPU(0); gpu.load_synthetic(cfg, ck.SynthSpec())
s = Session(gpu)
toks = [1] + [(7919 * i + 13) % cfg.vocab_size for i in range(1, 512)]
for _ in range(3): s.prefill(toks, 0)
ms = [s.prefill(toks, 0)[0] for _ in range(8)]
print("prefill pdl", os.environ["RAMA_PREFILL_PDL"], round(sum(ms)/len(ms), 3), "ms", [round(m,2) for m in ms])
PY
done
