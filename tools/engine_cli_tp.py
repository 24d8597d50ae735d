#!/usr/bin/env python
"""The C++ engine CLI (≙ engine/src/main.rs with --features gpu) on N GPUs of ONE process — no torchrun — against the CPU oracle.

    python tools/engine_cli_tp.py --gpus 8 [--model llama2-7B] [--steps 64] [--dir /dev/shm]

Writes the synthetic checkpoint (llama2.c v0 .bin, the bench's seed) and a tokenizer whose piece for id i is "w<i>." so the
printed text can be read back as token ids, runs `rama_b200/host/engine -m … -t … -s steps -r 0 --gpus N`, decodes the same
number of greedy tokens with the oracle on the same file and compares.  Prints ENGINE-TP-OK on success."""
import argparse
import os
import struct
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402  (a tool, not the product: the oracle is the checker here)
from rama_b200 import checkpoint as ck  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--model", default="llama2-7B")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--dir", default="/dev/shm")
    args = ap.parse_args()
    cfg = ck.CONFIGS[args.model]
    spec = ck.SynthSpec()
    ref.lib().ref_set_threads(len(os.sched_getaffinity(0)))
    path = os.path.join(args.dir, f"rama_{args.model}.bin")
    tok = os.path.join(args.dir, "rama_tok.bin")
    t0 = time.time()
    tensors = ref.synth_tensors(cfg, spec)
    ck.write_checkpoint(path, cfg, tensors)
    del tensors
    print(f"checkpoint {path}: {os.path.getsize(path) / 1e9:.2f} GB in {time.time() - t0:.1f} s", flush=True)
    pieces = ["[unk]", "<s>", "[/s]"] + [f"w{i}." for i in range(3, cfg.vocab_size)]
    with open(tok, "wb") as f:   # llama2.c tokenizer.bin layout (bpe.rs:27-43)
        f.write(struct.pack("<I", max(len(p) for p in pieces)))
        for p in pieces:
            b = p.encode()
            f.write(struct.pack("<fi", 0.0, len(b)) + b)
    try:
        t0 = time.time()
        r = subprocess.run([os.path.join(ROOT, "rama_b200", "host", "engine"), "-m", path, "-t", tok, "-s", str(args.steps), "-r", "0.0",
                            "--gpus", str(args.gpus)], capture_output=True, text=True, timeout=900)
        print(f"engine --gpus {args.gpus}: rc {r.returncode} in {time.time() - t0:.1f} s (load included)", flush=True)
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            raise SystemExit(1)
        text, tail = r.stdout.split("\n--------------------------------\n")
        print(tail.strip())
        got = [int(w[1:]) for w in text.split(".") if w.startswith("w")]
        _, mm = ck.read_checkpoint(path)          # mmap views of the same file
        om = ref.Model(cfg, mm)
        want, _, gap, el = ref.generate(om, ref.State(om), [], args.steps, 0.0, 0.9)
        want = [int(t) for t in want]
        print(f"oracle: {args.steps} tokens in {el:.1f} s, min top-2 gap {gap:.2e}")
        if got != [t for t in want if t >= 3]:
            print("MISMATCH", got[:16], want[:16])
            raise SystemExit(1)
        print(f"ENGINE-TP-OK gpus={args.gpus} model={args.model} tokens={len(got)}")
    finally:
        for p in (path, tok):
            try:
                os.remove(p)
            except OSError:
                pass


if __name__ == "__main__":
    main()
