"""Generates tests/golden/*.bin + *.npz from the REFERENCE's own Python code.

Runs only in the build container (needs /root/reference and torch CPU); the outputs are
committed so that the tests never read /root/reference.  What it pins:

  engine/export/model.py   Transformer (llama2.c training-side model: same network)
  engine/export/export.py  legacy_export → the v0 .bin layout the Rust engine reads (ram.rs:30-49)

For each case: build a small Transformer, randomise every parameter (incl. RMSNorm weights),
export it with the reference's legacy_export, run the reference's torch forward on a fixed
token sequence and save the logits of every position.  tests/test_oracle_golden.py then checks
that the C++ restatement of the Rust engine, fed the exported .bin token by token, reproduces
those logits (f32 tolerance) and the greedy continuation.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/engine/export"
sys.path.insert(0, REF)
from model import ModelArgs, Transformer  # noqa: E402  (reference code)
from export import legacy_export  # noqa: E402  (reference code)

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (ModelArgs kwargs, shared classifier?, seed, n_tokens)
    "ref_shared": (dict(dim=32, n_layers=2, n_heads=2, vocab_size=96, hidden_dim=80, multiple_of=8,
                        max_seq_len=24), True, 11, 24),
    "ref_untied": (dict(dim=48, n_layers=3, n_heads=4, vocab_size=64, hidden_dim=None, multiple_of=16,
                        max_seq_len=20), False, 12, 20),
    # stories15M's head geometry (head_size 48: 12 of 32 lanes per K/V row) and a context that spans three 32-timestep
    # attention chunks — pins RoPE at hs/2 = 24 and the chunked flash-decode kernels against the reference's torch attention
    "ref_hs48": (dict(dim=96, n_layers=2, n_heads=2, vocab_size=128, hidden_dim=128, multiple_of=8,
                      max_seq_len=72), True, 13, 72),
}


def make(name, kw, shared, seed, n_tok):
    torch.manual_seed(seed)
    m = Transformer(ModelArgs(**kw))
    if not shared:
        m.output.weight = torch.nn.Parameter(m.output.weight.detach().clone())  # break weight tying
    with torch.no_grad():
        for pn, p in m.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, kw["dim"] ** -0.5)
            else:
                p.copy_(1.0 + 0.2 * torch.randn_like(p))
    m.eval()
    g = torch.Generator().manual_seed(seed + 100)
    toks = torch.randint(3, kw["vocab_size"], (1, n_tok), generator=g)
    toks[0, 0] = 1  # BOS, as mod.rs:182
    with torch.no_grad():
        logits = m(toks, targets=toks)[0].float().numpy()  # targets given ⇒ logits for every position
    path = os.path.join(HERE, name + ".bin")
    legacy_export(m, path)
    np.savez(os.path.join(HERE, name + ".npz"), tokens=toks[0].numpy().astype(np.int32), logits=logits)
    print(name, "bin bytes", os.path.getsize(path), "logits", logits.shape)


if __name__ == "__main__":
    # python make_golden.py [case ...]   (no argument: every case; committed fixtures are only rewritten on purpose)
    for name, (kw, shared, seed, n) in CASES.items():
        if len(sys.argv) == 1 or name in sys.argv[1:]:
            make(name, kw, shared, seed, n)
