"""CPU models of two synchronisation protocols that were written without a GPU at hand, run under randomised
interleavings, so that a wrong mbarrier parity or a counter race shows up here as a deadlock / hazard instead of as a
hung kernel on the box.  The models restate the loops of the kernels statement by statement (file:function cited);
they do not prove the hardware semantics, they pin the protocol logic.

  1. gemm.cu gemm_tcgen05_persistent_kernel — TMA producer / MMA issuer / epilogue over the shared-memory ring
     (full/empty) and the double-buffered TMEM accumulator (tmem_full/tmem_empty), persistent over tiles.
  2. common.cuh FlagSync as engine.cu wires it — per-op completion counters + token epoch across g_step / g_body graphs.
"""
import random

import pytest


# ------------------------------------------------------------------------------------------------- mbarrier model
class MBar:
    """mbarrier with an arrival count and a transaction count.  try_wait.parity(P) succeeds iff the phase with parity P
    has completed, i.e. the barrier's current phase parity differs from P (a fresh barrier passes parity 1)."""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._maybe_complete()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        assert self.tx >= 0
        self._maybe_complete()

    def passed(self, parity):
        return (self.phase & 1) != parity


def run_persistent_gemm(tiles_for_cta, kblocks, stages, rng):
    """One CTA of the persistent kernel working through `tiles_for_cta` tiles.  Returns per-tile checks performed."""
    full = [MBar(1) for _ in range(stages)]
    empty = [MBar(1) for _ in range(stages)]
    tmem_full = [MBar(1), MBar(1)]
    tmem_empty = [MBar(128), MBar(128)]
    smem_fill = [None] * stages           # (tile, kb) whose operands sit in the stage, None = garbage
    smem_busy = [0] * stages              # MMAs issued on the stage that have not completed
    tmem_acc = [None, None]               # per buffer: (tile, k-blocks accumulated so far)
    tmem_readers = [0, 0]                 # epilogue threads still reading the buffer
    async_q = []                          # in-order completion queue of the tensor pipe: ("mma", s) | ("commit", bar)
    tma_q = []                            # TMA loads in flight: (stage, tile, kb)
    done_tiles = []

    def producer():
        s, ph = 0, 1
        for t in tiles_for_cta:
            for kb in range(kblocks):
                while not empty[s].passed(ph):
                    yield
                assert smem_busy[s] == 0, "producer overwrites a stage the tensor pipe is still reading"
                full[s].arrive_expect_tx(2)
                tma_q.append((s, t, kb))      # A and B boxes: one unit each
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
            yield

    def mma():
        s, ph, i = 0, 0, 0
        for t in tiles_for_cta:
            acc = i & 1
            while not tmem_empty[acc].passed(((i >> 1) & 1) ^ 1):
                yield
            assert tmem_readers[acc] == 0, "MMA overwrites an accumulator the epilogue is still reading"
            for kb in range(kblocks):
                while not full[s].passed(ph):
                    yield
                assert smem_fill[s] == (t, kb), f"MMA reads stage {s} holding {smem_fill[s]}, wanted {(t, kb)}"
                smem_busy[s] += 1
                async_q.append(("mma", s, acc, t, kb))
                async_q.append(("commit", empty[s]))
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
            async_q.append(("commit", tmem_full[acc]))
            i += 1
            yield

    def epilogue_thread(tid):
        i = 0
        for t in tiles_for_cta:
            acc = i & 1
            while not tmem_full[acc].passed((i >> 1) & 1):
                yield
            assert tmem_acc[acc] == (t, kblocks), f"epilogue reads {tmem_acc[acc]}, wanted tile {t} complete"
            tmem_readers[acc] += 1
            for _ in range(rng.randint(0, 3)):
                yield
            assert tmem_acc[acc] == (t, kblocks), "accumulator changed under the epilogue"
            tmem_readers[acc] -= 1
            tmem_empty[acc].arrive()
            if tid == 0:
                done_tiles.append(t)
            i += 1
            yield

    actors = [producer(), mma()] + [epilogue_thread(i) for i in range(128)]
    alive = set(range(len(actors)))
    idle_rounds = 0
    while alive:
        progressed = False
        # asynchronous agents: TMA completions (any order), tensor pipe (in order)
        if tma_q and rng.random() < 0.6:
            s, t, kb = tma_q.pop(rng.randrange(len(tma_q)))
            smem_fill[s] = (t, kb)
            full[s].complete_tx(2)
            progressed = True
        if async_q and rng.random() < 0.6:
            ev = async_q.pop(0)
            if ev[0] == "mma":
                _, s, acc, t, kb = ev
                assert smem_fill[s] == (t, kb), "operands were overwritten before the MMA read them"
                smem_busy[s] -= 1
                tmem_acc[acc] = (t, 1) if kb == 0 else (t, tmem_acc[acc][1] + 1)
            else:
                ev[1].arrive()
            progressed = True
        for a in rng.sample(sorted(alive), k=min(len(alive), 12)):
            state_before = (tuple(b.phase for b in full + empty + tmem_full + tmem_empty), len(async_q), len(tma_q),
                            len(done_tiles))
            try:
                next(actors[a])
            except StopIteration:
                alive.discard(a)
                progressed = True
                continue
            if state_before != (tuple(b.phase for b in full + empty + tmem_full + tmem_empty), len(async_q), len(tma_q),
                                len(done_tiles)):
                progressed = True
        idle_rounds = 0 if (progressed or tma_q or async_q) else idle_rounds + 1
        assert idle_rounds < 2000, "deadlock: no actor can make progress"
    assert done_tiles == list(tiles_for_cta)
    assert not tma_q and not async_q


@pytest.mark.parametrize("ntiles,kblocks,stages", [(1, 1, 4), (1, 7, 4), (2, 3, 4), (3, 4, 4), (5, 1, 4), (6, 9, 4),
                                                   (4, 2, 2), (7, 5, 3)])
def test_persistent_gemm_barrier_protocol(ntiles, kblocks, stages):
    for seed in range(6):
        rng = random.Random(1000 * ntiles + 10 * kblocks + seed)
        run_persistent_gemm([3 + 148 * i for i in range(ntiles)], kblocks, stages, rng)


# ----------------------------------------------------------------------------------------------- flag-sync model
def run_flagsync(tokens, nlayer_ops, ctas_per_op, rng):
    """`tokens` is a list of 'step' (with lm_head + argmax) / 'body' graphs launched back to back on one stream.
    Ops of a token: 0 = embed (1 CTA, full dependency), 1 … n = the GEMV/attention chain (PDL-launched: an op's CTAs may
    start as soon as every CTA of the previous op has STARTED), then for 'step': head (polls) and argmax (full
    dependency, advances the epoch); for 'body' the last chain op has a full dependency and advances the epoch.
    Checks: a CTA passes its wait only when every CTA of its producer op has finished in the SAME token."""
    ctr = {}
    epoch = [0]
    for tok_idx, kind in enumerate(tokens):
        n_ops = 1 + nlayer_ops + (1 if kind == "step" else 0)     # embed + chain (+ head); argmax handled apart
        finished = {op: 0 for op in range(n_ops)}
        started = {op: 0 for op in range(n_ops)}
        ctas = {0: 1}
        for op in range(1, n_ops):
            ctas[op] = ctas_per_op[(op - 1) % len(ctas_per_op)]
        last_chain = nlayer_ops if kind == "body" else None       # full-dependency node of a body-only token

        def cta(op, token_epoch_reads):
            started[op] += 1
            yield
            full_dep = (op == 0) or (op == last_chain)
            if full_dep:
                while op > 0 and finished[op - 1] < ctas[op - 1]:
                    yield                                          # the runtime holds the node back
            else:
                e = epoch[0]
                token_epoch_reads.append(e)
                target = (e + 1) * ctas[op - 1]
                while ctr.get(op - 1, 0) < target:
                    yield
                assert finished[op - 1] == ctas[op - 1], (
                    f"token {tok_idx} op {op}: wait passed with {finished[op - 1]}/{ctas[op - 1]} producers done")
            for _ in range(rng.randint(0, 2)):
                yield
            finished[op] += 1
            ctr[op] = ctr.get(op, 0) + 1
            if op == last_chain and False:
                pass

        reads = []
        pending_ops = list(range(n_ops))
        actors = []
        launched = set()
        while pending_ops or actors:
            # launch rule: op 0 immediately; op k when all CTAs of k-1 have started (PDL) — or finished for a full dep
            for op in list(pending_ops):
                if op == 0:
                    ok = True
                elif op == last_chain:
                    ok = finished[op - 1] == ctas[op - 1]
                else:
                    ok = (op - 1) in launched and started[op - 1] == ctas[op - 1]
                if ok:
                    actors += [cta(op, reads) for _ in range(ctas[op])]
                    launched.add(op)
                    pending_ops.remove(op)
                else:
                    break
            rng.shuffle(actors)
            nxt = []
            for a in actors:
                try:
                    next(a)
                    nxt.append(a)
                except StopIteration:
                    pass
            actors = nxt
        # end of token: the last node (argmax for 'step', the last chain op for 'body') advances the epoch — only after
        # every CTA of the token has read it (all actors are done here, which is what the full dependency guarantees)
        assert all(r == epoch[0] for r in reads)
        epoch[0] += 1
        # counters every later token relies on must have advanced by exactly ctas per token for the ops of BOTH graphs
        for op in range(1 + nlayer_ops):
            assert ctr[op] == epoch[0] * ctas[op], (op, ctr[op], epoch[0], ctas[op])


def test_flagsync_counter_protocol():
    for seed in range(8):
        rng = random.Random(seed)
        kinds = [rng.choice(["step", "body"]) for _ in range(7)]
        run_flagsync(kinds, nlayer_ops=10, ctas_per_op=[5, 3, 4, 6, 2], rng=rng)


# ------------------------------------------------------------------------------ persistent GEMM: tile walk + epilogue cover
@pytest.mark.parametrize("M,N,sms", [(128, 256, 148), (1, 128, 148), (130, 136, 148), (200, 896, 148), (16, 1152, 148),
                                     (2048, 4096, 148), (384, 40000, 148), (512, 12288, 7)])
def test_persistent_gemm_writes_every_element_once(M, N, sms):
    """gemm.cu gemm_tcgen05_persistent_kernel: tiles t = blockIdx.x + i·gridDim.x, m-fastest (m0 = (t % mt)·128,
    n0 = (t / mt)·256); epilogue warp quarter q / lane own row m0 + 32q + lane and walk 256 columns in 16-wide pieces,
    guarded by row < M and n0 + c < N (+ per-element guard in the ragged piece)."""
    import numpy as np
    BM, BN = 128, 256
    mt, nt = (M + BM - 1) // BM, (N + BN - 1) // BN
    tiles = mt * nt
    grid = min(tiles, sms)
    hits = np.zeros((M, N), dtype=np.int32)
    acc_use = {}
    for b in range(grid):
        i = 0
        for t in range(b, tiles, grid):
            m0, n0 = (t % mt) * BM, (t // mt) * BN
            acc_use.setdefault(b, []).append(i & 1)
            rows = np.arange(m0, m0 + BM)
            rows = rows[rows < M]
            for c in range(0, BN, 16):
                if n0 + c < N:
                    hits[np.ix_(rows, np.arange(n0 + c, min(N, n0 + c + 16)))] += 1
            i += 1
    assert (hits == 1).all()
    assert all(u == [k & 1 for k in range(len(u))] for u in acc_use.values())   # accumulator buffers strictly alternate
