#!/usr/bin/env python
"""Exhaustive interleaving check of the cross-GPU flag protocols (host model, no GPU needed).

The kernels in csrc/cuda/collectives.cu and the staged pull of csrc/cuda/gemm_tcgen05.cu synchronise W GPUs with nothing but
plain stores / loads / atomics on peer memory: a release store of an epoch into the peer's signal slot, an acquire spin on the
own slot, staging slots that alternate with the epoch's parity, arrival counters.  Their comments argue informally why a slot
is never overwritten before it was read and why nobody waits forever.  This tool turns each kernel's synchronisation skeleton
into W straight-line programs over a shared memory and explores EVERY interleaving of their steps (sequentially consistent
steps; the release / acquire pairs of the kernels are what makes the real execution equivalent to one of them), checking

  * safety   - every load observes exactly the value the algorithm needs (the peer's data of THIS call: not a stale one of an
               earlier call, not one a fast peer already wrote for a later call),
  * liveness - no reachable state in which an unfinished rank can never step again (a lost / overwritten flag = deadlock).

Each protocol comes with a NEGATIVE CONTROL - the same program with the protecting mechanism removed (one staging slot instead
of two, no exit barrier, one stage buffer) - that the checker must reject; a checker that cannot find those bugs proves nothing.

    python tools/protocol_model.py                 # all protocols, W = 2 and 3, prints a table (profiles/r2/protocol_model.md)
    python tools/protocol_model.py --world 4 --calls 2 --only oneshot

The GPU-side counterparts are tests/mp/protocol_worker.py (random per-CTA delays before every signal, fault injection) and
tools/sanitize_collectives.sh (compute-sanitizer racecheck / synccheck).  Reference analogue: ps-lite's PS_DROP_MSG fault
injection and its resender (PS/src/van.cc:297-299, resender.h) test the message protocol the same way - by construction of
adverse schedules.
"""
from __future__ import annotations

import argparse
import sys
import time
from collections import deque

# ----------------------------------------------------------------------------- programs
# An op is a tuple:
#   ("st", loc, value)          store
#   ("add", loc, delta)         atomic add (arrival counters)
#   ("wge", loc, value)         spin until mem[loc] >= value          (flag waits: monotonic epochs)
#   ("weq", loc, value)         spin until mem[loc] == value          (flag-in-data: the slot must hold exactly this packet)
#   ("ld", loc, value, what)    load and require mem[loc] == value    (safety assertion)
# Locations are hashable names; absent = 0.


_DROP = None          # (src, dst): rank src never signals rank dst (symm.cuh: SymmDev::drop_peer, the fault injection hook)


def _barrier(me: int, W: int, chan: str, val: int) -> list:
    """symm.cuh:block_barrier - one release store per peer, one acquire spin per peer."""
    return [("st", ("sig", chan, p, me), val) for p in range(W) if _DROP != (me, p)] + \
           [("wge", ("sig", chan, me, p), val) for p in range(W)]


def prog_twoshot(me: int, W: int, K: int, exit_barrier: bool = True) -> list:
    """kallreduce / kallreduce_mc / kps_sync: entry barrier, rank r reduces slice r of everybody's buffer and writes the result
    into slice r of everybody's buffer (in place), exit barrier.  Between calls the owner produces new input in place."""
    ops = []
    for k in range(K):
        ops += [("st", ("buf", me, c), ("in", me, k, c)) for c in range(W)]          # the producer (backward) fills my bucket
        ops += _barrier(me, W, "ar", 2 * k + 1)
        ops += [("ld", ("buf", p, me), ("in", p, k, me), f"call {k}: slice {me} of rank {p}'s input") for p in range(W)]
        ops += [("st", ("buf", p, me), ("out", k, me)) for p in range(W)]
        if exit_barrier:
            ops += _barrier(me, W, "ar", 2 * k + 2)
        ops += [("ld", ("buf", me, c), ("out", k, c), f"call {k}: reduced slice {c} in my buffer") for c in range(W)]   # the consumer (optimizer)
    return ops


def prog_oneshot(me: int, W: int, K: int, slots: int = 2) -> list:
    """kallreduce_oneshot: push my input into staging slot [epoch parity][me] of every peer, ONE barrier, reduce my own copies.
    No exit barrier - the parity alternation is the protection."""
    ops = []
    for k in range(K):
        par = k % slots
        ops += [("st", ("stage", p, par, me), ("in", me, k)) for p in range(W)]
        ops += _barrier(me, W, "os", k + 1)
        ops += [("ld", ("stage", me, par, p), ("in", p, k), f"call {k}: rank {p}'s staged input") for p in range(W)]
    return ops


def prog_ll(me: int, W: int, K: int, slots: int = 2) -> list:
    """kallreduce_ll: (payload, epoch) packets, no barrier at all; the receiver polls its slot until the epoch matches."""
    ops = []
    for k in range(K):
        par = k % slots
        ops += [("st", ("ll", p, par, me), (k + 1, me)) for p in range(W)]
        ops += [("weq", ("ll", me, par, p), (k + 1, p)) for p in range(W)]
    return ops


def prog_staged_pull(me: int, W: int, K: int, buffers: int = 2, chunks: int = 2) -> list:
    """gemm_tcgen05.cu staged_push / staged_wait_rows: launch i multicasts my shard chunk by chunk into stage buffer i % buffers
    of every rank and bumps the chunk's arrival counter everywhere (multimem.red.release); the TMA producer consumes a chunk of
    ANY rank's shard once its counter reached the launch's sequence number.  No entry barrier: two alternating buffers."""
    ops = []
    for i in range(K):
        b = i % buffers
        for c in range(chunks):
            ops += [("st", ("stg", p, b, me, c), ("w", me, i, c)) for p in range(W)]
            ops += [("add", ("cnt", p, me, c), 1) for p in range(W)]
        for q in range(W):
            for c in range(chunks):
                ops.append(("wge", ("cnt", me, q, c), i + 1))
                ops.append(("ld", ("stg", me, b, q, c), ("w", q, i, c), f"launch {i}: chunk {c} of rank {q}'s shard"))
    return ops


def progs_ps_async(W: int, K: int, wait_ack: bool = True) -> list:
    """kps_async_worker / kps_async_server (rank 0 = server, ranks 1.. = workers): worker w leaves push #k in its own heap and
    raises its flag on the server; the server takes the raised flags in ARRIVAL order (any order - modelled as one server
    thread per worker: every order the kernel's polling loop can produce is an interleaving of those), reads THAT worker's
    gradient over the fabric, updates, writes the new weights into that worker's heap and answers on the second channel; the
    worker produces its next gradient only after the answer."""
    progs = []
    for w in range(1, W):
        srv, wrk = [], []
        for k in range(K):
            wrk.append(("st", ("grad", w), ("g", w, k)))
            wrk.append(("st", ("flag", w), k + 1))
            if wait_ack:
                wrk.append(("wge", ("ack", w), k + 1))
                wrk.append(("ld", ("wts", w), ("w", w, k), f"weights answering my push {k}"))
            srv.append(("wge", ("flag", w), k + 1))
            srv.append(("ld", ("grad", w), ("g", w, k), f"gradient of push {k} in worker {w}'s heap"))
            srv.append(("st", ("wts", w), ("w", w, k)))
            srv.append(("st", ("ack", w), k + 1))
        progs += [(f"worker {w}", wrk), (f"server (thread of worker {w})", srv)]
    return progs


def _per_rank(prog):
    return lambda W, K, **kw: [(f"rank {r}", prog(r, W, K, **kw)) for r in range(W)]


PROTOCOLS = {
    # name: (builder of all programs, kwargs of the real kernel, kwargs of the negative control, what the control removes)
    "twoshot": (_per_rank(prog_twoshot), {}, {"exit_barrier": False}, "no exit barrier"),
    "oneshot": (_per_rank(prog_oneshot), {"slots": 2}, {"slots": 1}, "one staging slot (no parity)"),
    "ll": (_per_rank(prog_ll), {"slots": 2}, {"slots": 1}, "one packet slot (no parity)"),
    "staged_pull": (_per_rank(prog_staged_pull), {"buffers": 2}, {"buffers": 1}, "one stage buffer"),
    "ps_async": (progs_ps_async, {}, {"wait_ack": False}, "worker does not wait for the answer"),
}


# ----------------------------------------------------------------------------- explorer
class Verdict:
    def __init__(self):
        self.states = 0
        self.transitions = 0
        self.violation = None          # text, or None
        self.trace = None

    @property
    def ok(self) -> bool:
        return self.violation is None


def explore(named_programs: list, max_states: int = 20_000_000) -> Verdict:
    """Breadth-first search over (program counters, memory) of a list of (name, program).  Memory is kept as a tuple over the
    sorted set of locations the programs mention."""
    names = [n for n, _ in named_programs]
    programs = [p for _, p in named_programs]
    locs = sorted({op[1] for prog in programs for op in prog}, key=repr)
    index = {loc: i for i, loc in enumerate(locs)}
    progs = [[(op[0], index[op[1]]) + tuple(op[2:]) for op in prog] for prog in programs]
    n = len(progs)
    lens = [len(p) for p in progs]
    v = Verdict()
    start = (tuple([0] * n), tuple([0] * len(locs)))
    seen = {start: None}
    queue = deque([start])
    while queue:
        state = queue.popleft()
        pcs, mem = state
        v.states += 1
        moved = False
        for r in range(n):
            pc = pcs[r]
            if pc == lens[r]:
                continue
            op = progs[r][pc]
            kind, loc = op[0], op[1]
            cur = mem[loc]
            new_mem = mem
            if kind == "st":
                if cur != op[2]:
                    new_mem = mem[:loc] + (op[2],) + mem[loc + 1:]
            elif kind == "add":
                new_mem = mem[:loc] + (cur + op[2],) + mem[loc + 1:]
            elif kind == "wge":
                if not (isinstance(cur, int) and cur >= op[2]):
                    continue
            elif kind == "weq":
                if cur != op[2]:
                    continue
            elif kind == "ld":
                if cur != op[2]:
                    v.violation = f"{names[r]} reads {locs[loc]} for [{op[3]}]: expected {op[2]}, found {cur}"
                    v.trace = _trace(seen, state, programs, names) + [f"{names[r]}: {programs[r][pc]}"]
                    return v
            moved = True
            v.transitions += 1
            nxt = (pcs[:r] + (pc + 1,) + pcs[r + 1:], new_mem)
            if nxt not in seen:
                seen[nxt] = (state, r)
                queue.append(nxt)
                if len(seen) > max_states:
                    raise RuntimeError(f"more than {max_states} states")
        if not moved and any(pcs[r] < lens[r] for r in range(n)):
            stuck = [f"{names[r]} at {programs[r][pcs[r]]} (memory holds {mem[progs[r][pcs[r]][1]]})" for r in range(n) if pcs[r] < lens[r]]
            v.violation = "deadlock: " + "; ".join(stuck)
            v.trace = _trace(seen, state, programs, names)
            return v
    return v


def _trace(seen: dict, state, programs: list, names: list) -> list:
    steps = []
    while seen[state] is not None:
        prev, r = seen[state]
        steps.append(f"{names[r]}: {programs[r][prev[0][r]]}")
        state = prev
    return steps[::-1]


def check(name: str, W: int, K: int, control: bool = False, drop=None, **extra) -> Verdict:
    """control: the variant without the protecting mechanism; drop = (src, dst): fault injection - src never signals dst, which
    every barrier-based protocol must turn into a wait that never completes at dst (on the device: the spin timeout fires and
    sets the error word, tests/mp/protocol_worker.py) and never into a wrong value."""
    global _DROP
    build, real, neg, _ = PROTOCOLS[name]
    kw = neg if control else real
    if kw is None:
        raise ValueError(f"{name} has no negative control")
    kw = dict(kw, **extra)
    if name == "staged_pull" and W >= 4:
        kw.setdefault("chunks", 1)        # one chunk per shard keeps three launches of four ranks enumerable
    _DROP = drop
    try:
        return explore(build(W, K, **kw))
    finally:
        _DROP = None


def default_calls(name: str, W: int) -> int:
    """Back-to-back calls per rank: >= 3 wherever slot reuse (call k + 2 against call k) is the thing under test."""
    if W == 2:
        return 4
    if W == 3:
        return 2 if name == "twoshot" else 3
    return 2 if name == "twoshot" else 3


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--world", type=int, nargs="*", default=[2, 3])
    ap.add_argument("--calls", type=int, default=0, help="back-to-back calls per rank (default: see default_calls)")
    ap.add_argument("--only", nargs="*", default=list(PROTOCOLS))
    ap.add_argument("--markdown", action="store_true")
    ap.add_argument("--trace", action="store_true", help="print the counterexample of every rejected variant")
    a = ap.parse_args(argv)
    rows, bad = [], 0
    for name in a.only:
        for W in a.world:
            K = a.calls or default_calls(name, W)
            t0 = time.time()
            v = check(name, W, K)
            rows.append((name, W, K, "protocol", v.states, v.transitions, "safe + live" if v.ok else "VIOLATION: " + v.violation, time.time() - t0))
            bad += not v.ok
            if PROTOCOLS[name][2] is not None:
                t0 = time.time()
                c = check(name, W, K, control=True)
                rows.append((name, W, K, "control: " + PROTOCOLS[name][3], c.states, c.transitions,
                             "rejected (" + ("deadlock" if c.violation.startswith("deadlock") else "wrong value read") +
                             f", {len(c.trace)}-step counterexample)" if not c.ok else "NOT REJECTED", time.time() - t0))
                bad += c.ok
                if a.trace and not c.ok:
                    print(f"--- {name} W={W}, {PROTOCOLS[name][3]}: {c.violation}\n    " + "\n    ".join(c.trace), file=sys.stderr)
            if name in ("twoshot", "oneshot"):
                t0 = time.time()
                f = check(name, W, K, drop=(0, 1))
                stuck_at_victim = (not f.ok) and f.violation.startswith("deadlock") and "rank 1 at ('wge'" in f.violation
                rows.append((name, W, K, "fault: rank 0 never signals rank 1", f.states, f.transitions,
                             "rank 1 waits forever, no wrong value (device: spin timeout -> error word)" if stuck_at_victim
                             else "UNEXPECTED: " + str(f.violation), time.time() - t0))
                bad += not stuck_at_victim
    if a.markdown:
        print("| protocol | W | calls | variant | states | transitions | result |")
        print("|---|---|---|---|---|---|---|")
        for r in rows:
            print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]} | {r[6]} |")
    else:
        for r in rows:
            print(f"{r[0]:12s} W={r[1]} calls={r[2]} {r[3]:42s} states={r[4]:<9d} transitions={r[5]:<10d} {r[6]}  [{r[7]:.1f}s]")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
