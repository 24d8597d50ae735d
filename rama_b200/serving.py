"""Host-side request batching over the batched decode path — the server side of the hot path.

Reference: `engine::EngineService::handler` (engine/src/lib.rs:127-160) spawns one task per request, each with
its own RunState, looping `forward` + `sample` (generate_stream, mod.rs:208-248): n live requests stream the
weights n times per token.  `server/src/batcher.rs:8-38` (`get_batch(receiver, prompts, batch_size, wait_time)`)
was meant to group requests but is dead code.  Here the same pieces are wired to `rama_forward_batch`:

* `get_batch` — batcher.rs semantics on a `queue.Queue`: up to `batch_size` prompts or until `wait_time` elapsed;
* `BatchedEngine` — one Session (≙ RunState) per live request, prompts prefilled on admission
  (`rama_prefill`), then ALL live requests advance together, one `rama_forward_batch` + `rama_sample_batch`
  per token (continuous batching: requests join and leave between steps).  Per request the token stream is the
  one `generate()` (mod.rs:169-206) produces: the forced prompt tokens, then the sampled ones, `steps` in total.

Pure host logic; every device call goes through the C ABI (rama_b200.engine).
"""
from __future__ import annotations

import queue
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

from .engine import GPU, Batch, Session, Tokenizer

__all__ = ["get_batch", "Request", "BatchedEngine"]


def get_batch(receiver: "queue.Queue", prompts: List, batch_size: int, wait_time: float) -> None:
    """≙ batcher.rs:8-38: append received items to `prompts` until it holds `batch_size` of them or `wait_time`
    seconds have passed (tokio::time::timeout around the receive loop)."""
    deadline = time.monotonic() + wait_time
    while len(prompts) < batch_size:
        left = deadline - time.monotonic()
        if left <= 0:
            return
        try:
            prompts.append(receiver.get(timeout=left))
        except queue.Empty:
            return


@dataclass
class Request:
    rid: int
    prompt_tokens: List[int]
    steps: int
    on_token: Optional[Callable[[int, int], None]] = None   # (request id, token) as generate_stream's sender
    tokens: List[int] = field(default_factory=list)          # `next` of every step so far (mod.rs:187-203)
    session: Optional[Session] = None
    pos: int = 0                                             # next position to feed
    cur: int = 1                                             # token to feed at `pos`
    done: bool = False

    def text(self, tok: Tokenizer) -> str:
        return b"".join(tok.decode(t) for t in self.tokens).decode("utf-8", errors="replace")


class BatchedEngine:
    def __init__(self, gpu: GPU, max_batch: int = 64, temperature: float = 0.0, topp: float = 0.9,
                 prefill_min_rows: int = 2):
        self.gpu, self.cfg = gpu, gpu.cfg
        self.max_batch = max_batch
        self.temperature, self.topp = temperature, topp
        self.prefill_min_rows = prefill_min_rows
        self.batch = Batch(gpu, max_batch)
        self.pool: List[Session] = []          # idle sessions (KV rows are rewritten before they are read again)
        self.live: List[Request] = []
        self.waiting: List[Request] = []
        self.finished: Dict[int, Request] = {}
        self._next_id = 0

    # ---- admission ---------------------------------------------------------------------------------------
    def submit(self, prompt_tokens: Sequence[int], steps: int, on_token=None) -> int:
        """Queue a request; it joins the batch at the next step boundary.  steps as in generate(): total number
        of positions (prompt included); steps > seq_len is the reference's out-of-bounds panic."""
        if steps > self.cfg.seq_len:
            raise ValueError(f"steps {steps} exceeds seq_len {self.cfg.seq_len}")
        r = Request(self._next_id, [int(t) for t in prompt_tokens], int(steps), on_token)
        self._next_id += 1
        self.waiting.append(r)
        return r.rid

    def _emit(self, r: Request, token: int):
        r.tokens.append(token)
        if r.on_token:
            r.on_token(r.rid, token)

    def _admit(self):
        while self.waiting and len(self.live) < self.max_batch:
            r = self.waiting.pop(0)
            if r.steps <= 0:
                r.done = True
                self.finished[r.rid] = r
                continue
            r.session = self.pool.pop() if self.pool else Session(self.gpu)
            rows = [1] + r.prompt_tokens               # BOS first (mod.rs:182), then the forced prompt
            n_forced = min(len(r.prompt_tokens), r.steps)
            # positions 0..n_forced-1 emit the prompt tokens themselves (mod.rs:189-191)
            for t in r.prompt_tokens[:n_forced]:
                self._emit(r, t)
            if n_forced >= r.steps:                    # the prompt alone exhausts the step budget
                # the reference still runs forward() for those positions; nothing observable depends on it
                r.done = True
                self.pool.append(r.session)
                r.session = None
                self.finished[r.rid] = r
                continue
            feed = rows[: n_forced + 1]                # tokens fed at positions 0..n_forced
            if len(feed) - 1 >= self.prefill_min_rows:
                # all but the last fed row through prefill; the last one joins the batched step so that its
                # logits come from the same path as everybody else's
                r.session.prefill(feed[:-1], 0)
            else:
                for p, t in enumerate(feed[:-1]):
                    r.session.forward(t, p)
                r.session.sync()
            r.pos, r.cur = len(feed) - 1, feed[-1]
            self.live.append(r)

    # ---- one token for every live request -------------------------------------------------------------------
    def step(self) -> int:
        self._admit()
        if not self.live:
            return 0
        sessions = [r.session for r in self.live]
        self.batch.forward(sessions, [r.cur for r in self.live], [r.pos for r in self.live])
        nxt = self.batch.sample(sessions, self.temperature, self.topp)
        still = []
        for r, t in zip(self.live, nxt):
            self._emit(r, t)
            r.pos += 1
            r.cur = t
            if len(r.tokens) >= r.steps:
                r.done = True
                self.pool.append(r.session)
                r.session = None
                self.finished[r.rid] = r
            else:
                still.append(r)
        n = len(self.live)
        self.live = still
        return n

    def run_until_idle(self):
        while self.waiting or self.live:
            self.step()

    def close(self):
        self.batch.close()
        for r in self.live:
            if r.session:
                r.session.close()
        for s in self.pool:
            s.close()
        self.live, self.pool = [], []
