"""Model check of the multi-GPU exchange protocol (gpunb_b200.cu: run_job / merge_kernel / combine_kernel, DESIGN.md
section 5), on the CPU.  The model has exactly the ordering rules the library relies on:

  * per rank, call k runs in pipeline slot q = k mod NS: pair(k) on stream lo[q], merge(k) then combine(k) on hi[q];
    pair(k) waits for combine(k - NS) (the slot's buffers are free), merge(k) waits for pair(k);
  * merge(k) writes this rank's shard result into exchange slot k mod X of its OWN buffer -- but only after every peer
    has acknowledged call k - X (acks written by the peers' combine(k - X));
  * merge(k) then raises flag k in every peer; combine(k) waits for flag k from every rank, reads every rank's slot
    k mod X and finally acknowledges call k to every rank.

Kernels that are ready run in RANDOM order (GPUs give no ordering between streams or between ranks).  Checked: every
schedule terminates (no deadlock), no exchange slot is overwritten while a peer still has to read it, and every combine
reads exactly the data of its own call."""
import random

import pytest


def simulate(R, NS, X, ncalls, rng):
    done = set()                               # (kind, rank, k)
    slot_data = {}                             # (rank, xslot) -> call number whose data it holds
    readers_left = {}                          # (rank, xslot) -> ranks that still have to read the current content
    todo = [(kind, r, k) for k in range(1, ncalls + 1) for r in range(R) for kind in ("pair", "merge", "combine")]

    def ready(kind, r, k):
        if kind == "pair":
            return k <= NS or ("combine", r, k - NS) in done
        if kind == "merge":
            if ("pair", r, k) not in done:
                return False
            if k > NS and ("combine", r, k - NS) not in done:       # stream order on hi[q]
                return False
            return k <= X or all(("combine", p, k - X) in done for p in range(R))     # acks
        return ("merge", r, k) in done and all(("merge", p, k) in done for p in range(R))   # flags

    while todo:
        runnable = [t for t in todo if ready(*t)]
        assert runnable, f"deadlock with {len(todo)} kernels left (R={R}, NS={NS}, X={X})"
        kind, r, k = rng.choice(runnable)
        if kind == "merge":
            key = (r, k % X)
            assert not readers_left.get(key), f"rank {r} overwrites exchange slot {k % X} (call {slot_data.get(key)}) before {readers_left[key]} read it"
            slot_data[key] = k
            readers_left[key] = set(range(R))
        elif kind == "combine":
            for p in range(R):
                key = (p, k % X)
                assert slot_data.get(key) == k, f"combine({k}) on rank {r} reads call {slot_data.get(key)} from rank {p}"
                readers_left[key].discard(r)
        done.add((kind, r, k))
        todo.remove((kind, r, k))


@pytest.mark.parametrize("R,NS,X", [(2, 1, 8), (2, 3, 8), (8, 3, 8), (4, 4, 8), (8, 4, 2), (3, 2, 1)])
def test_exchange_protocol_never_deadlocks_or_overwrites(R, NS, X):
    rng = random.Random(1000 * R + 10 * NS + X)
    for trial in range(20):
        simulate(R, NS, X, ncalls=3 * X + NS + 2, rng=rng)


def test_model_detects_a_missing_ack_wait():
    """Sanity of the model itself: without the ack rule a slot IS overwritten under some schedule."""
    rng = random.Random(7)

    def broken(R, NS, X, ncalls):
        done, slot_data, readers_left = set(), {}, {}
        todo = [(kind, r, k) for k in range(1, ncalls + 1) for r in range(R) for kind in ("pair", "merge", "combine")]
        while todo:
            runnable = []
            for kind, r, k in todo:
                if kind == "pair":
                    ok = k <= NS or ("combine", r, k - NS) in done
                elif kind == "merge":
                    ok = ("pair", r, k) in done and (k <= NS or ("combine", r, k - NS) in done)      # no ack wait
                else:
                    ok = all(("merge", p, k) in done for p in range(R))
                if ok:
                    runnable.append((kind, r, k))
            kind, r, k = rng.choice(runnable)
            if kind == "merge":
                key = (r, k % X)
                if readers_left.get(key):
                    return True
                slot_data[key] = k; readers_left[key] = set(range(R))
            elif kind == "combine":
                for p in range(R):
                    readers_left[(p, k % X)].discard(r)
            done.add((kind, r, k)); todo.remove((kind, r, k))
        return False

    assert any(broken(4, 3, 2, 12) for _ in range(50))
