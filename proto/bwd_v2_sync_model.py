"""Synchronisation design of "backward v2" (dataflow: proto/bwd_v2_blueprint.py), model-checked
before any CUDA exists: the warp roles are coroutines executing the PROPOSED mbarrier waits / arrivals / commits, the
tensor pipe is an in-order queue, the scheduler is random and adversarial (proto/fwd_v2_sync_model.py), and every access
asserts what it must find.  Chunks are processed last to first; `it` counts iterations, slot = it % 3.

  stage A (8 warps)   tiles of the slot + U rows by bulk copy                        -> a_done[slot] (8), full[slot] (8 of 10)
  Gram group F (2)    reads G1 (forward Gram blocks): Aqb^T, Aqk^T, Aak^T, T          -> full[slot] (2 of 10)
  MMA warp            G1(it+1) | B1: [R | dV] = dS [B~;K~]^T, R += dY^T Aqb, P2a      -> bar_r
                      B2: Z^T = R^T T                                                 -> bar_z
                      B3: dS += dY^T Q~ + Z^T A~                  (after bar_z)
                      G2: [dY;Z][U;V]^T                           (after z_done)       -> g2_ready
                      B5: dS^T update, P1, P2b, P3b               (after c_done, s0)   -> out_ready
  group C1 (4)        window rescale (resc) ; Z^T -> DYZn Z rows + Z^T tile            -> z_done (4)
  Gram group G (2)    reads G2: eight 16x16 gradient operand tiles                     -> c_done (2)
  group C2 (8)        output epilogue of chunk `it` (staging aliases the slot's UVn + DYZn), boundary term -> ok_free[it&1] (8),
                      empty[slot] (8), glp_done (8, at window boundaries)
out_ready is one barrier PER CHUNK PARITY.  With a single barrier (the shipped kernel's arrangement) nothing stops the MMA warp
from committing out_ready of chunk it+1 before group C2 has observed the phase of chunk it -- C2 is off the chain by design --
and a waiter that is two phases behind sees its own parity again and never wakes: the model finds that deadlock at once
under adversarial scheduling (last line of `--mutations`).  On the GPU it needs C2 to be descheduled for a whole chunk iteration
(~5000 cycles) while suspended on the barrier, which has not been observed, but two barriers close it for free: the next
commit on out_ready[u] is two iterations later, behind the wait on ok_free[u], which C2 signals after its out_ready wait.
Tensor memory: dS 0-63, dS^T 64-127, Z 128-143, R 144-159, G1 2 x 32 (160-223), G2 224-255, OK/OV double buffered 256-415.
"""
import random
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fwd_v2_sync_model import Bar  # noqa: E402

NS, WIN = 3, 4


class Kernel:
    def __init__(self, nC, single_out_ready=False):
        self.nC, self.single = nC, single_out_ready
        B = Bar
        self.full = [B(10) for _ in range(NS)]; self.empty = [B(8) for _ in range(NS)]; self.a_done = [B(8) for _ in range(NS)]
        self.blob_full = [B(1) for _ in range(NS)]
        self.g1_ready = [B(1), B(1)]; self.g2_ready = B(1)
        self.s0_full, self.bar_r, self.bar_z = B(1), B(1), B(1)
        self.out_ready = [B(1), B(1)]        # per chunk parity: see the note on the single-barrier version below
        self.z_done, self.c_done, self.resc, self.glp_done = B(4), B(2), B(4), B(8)
        self.ok_free = [B(8), B(8)]
        self.slot_a = [-1] * NS; self.slot_f = [-1] * NS; self.blob = [-1] * NS; self.slot_busy = [0] * NS
        self.G1 = [-1, -1]; self.G1_read = [-1, -1]; self.G2 = -1; self.G2_read = -1
        self.dS = 0; self.dST = 0            # iterations accumulated
        self.framed = 0                      # window rescales applied to dS / dS^T
        self.R = self.Z = self.ztiles = self.gtiles = self.s0 = -1
        self.OK = [-1, -1]; self.OK_read = [-1, -1]
        self.glp_taken = 0
        self.pipe = []
        self.copies = []                     # bulk copies in flight: an engine of its own, NOT ordered with the tensor pipe
        self.done_out = 0

    def out_bar(self, it):                   # (barrier, parity) a waiter of chunk `it` uses
        return (self.out_ready[0], it & 1) if self.single else (self.out_ready[it & 1], (it >> 1) & 1)

    def win_last(self, it):
        c = self.nC - 1 - it
        return (c % WIN == WIN - 1) or (c == self.nC - 1)

    def n_win(self, it):                     # window entries up to and including iteration `it`
        return sum(1 for j in range(it + 1) if self.win_last(j))

    def pipe_step(self):
        op = self.pipe.pop(0)
        kind, it = op[0], op[1]
        si, u = it % NS, it & 1
        if kind == "commit":
            op[2].arrive()
            return
        self.slot_busy[si] -= 1
        if kind == "g1":
            assert self.slot_a[si] == it, ("g1", it, self.slot_a[si])
            assert self.G1_read[u] == self.G1[u], ("g1 overwrites unread Gram blocks", self.G1[u])
            self.G1[u] = it
        elif kind == "b1":
            assert self.slot_a[si] == it and self.slot_f[si] == it and self.blob[si] == it, ("b1", it)
            assert self.dS == it and self.dST == it and self.framed == self.n_win(it), ("b1 frame", it, self.dS, self.dST, self.framed)
            assert self.OK_read[u] == self.OK[u], ("b1 overwrites undrained gradients", self.OK[u])
            self.R = it
            self.OK[u] = (it, "partial")
        elif kind == "b2":
            assert self.R == it and self.slot_f[si] == it
            self.Z = it
        elif kind == "b3":
            assert self.Z == it and self.dS == it and self.slot_a[si] == it
            self.dS = it + 1
        elif kind == "g2":
            assert self.ztiles == it and self.blob[si] == it and self.slot_a[si] == it, ("g2", it, self.ztiles)
            assert self.G2_read == self.G2, ("g2 overwrites unread Gram blocks", self.G2)
            self.G2 = it
        elif kind == "b5":
            assert self.gtiles == it and self.s0 == it and self.ztiles == it and self.dST == it, ("b5", it, self.gtiles, self.s0)
            assert self.OK[u] == (it, "partial")
            self.dST = it + 1
            self.OK[u] = it

    def copy_step(self):                     # S0^T checkpoint of chunk `it` lands in the single buffer
        it = self.copies.pop(0)
        assert self.s0 == it - 1 and self.dST == it, ("S0 buffer overwritten while B5 of the previous chunk may read it", self.s0, self.dST, it)
        self.s0 = it
        self.s0_full.arrive()

    # ---- roles --------------------------------------------------------------------------------------------
    def stage_a(self):
        for it in range(self.nC):
            si = it % NS
            if it >= NS:
                yield lambda: self.empty[si].done(((it // NS) - 1) & 1)
            assert self.slot_busy[si] == 0, ("stage A overwrites a slot still being read", it)
            self.slot_a[si] = it
            self.slot_busy[si] = 6           # g1, b1, b2, b3, g2, b5
            yield None
            self.blob[si] = it               # bulk copy of the U rows (completes some time later; in order here)
            self.blob_full[si].arrive()
            yield None
            self.a_done[si].arrive(8)
            self.full[si].arrive(8)

    def gram_f(self):
        for it in range(self.nC):
            si, u = it % NS, it & 1
            yield lambda: self.g1_ready[u].done((it >> 1) & 1)
            assert self.G1[u] == it, ("gram F reads", self.G1[u], "for", it)
            self.G1_read[u] = it
            yield None
            assert self.slot_a[si] == it
            self.slot_f[si] = it
            yield None
            self.full[si].arrive(2)

    def mma(self):
        yield lambda: self.a_done[0].done(0)
        self.pipe += [("g1", 0), ("commit", 0, self.g1_ready[0])]
        nw = 0
        for it in range(self.nC):
            si, u = it % NS, it & 1
            yield lambda: self.full[si].done((it // NS) & 1)
            yield lambda: self.blob_full[si].done((it // NS) & 1)
            if it > 0:
                yield lambda: self.out_bar(it - 1)[0].done(self.out_bar(it - 1)[1])
            if self.win_last(it):
                yield lambda nw=nw: self.resc.done(nw & 1)
                nw += 1
            if it >= 2:
                yield lambda: self.ok_free[u].done(((it >> 1) - 1) & 1)
            if it + 1 < self.nC:
                yield lambda: self.a_done[(it + 1) % NS].done(((it + 1) // NS) & 1)
            self.copies.append(it)
            self.pipe += [("b1", it), ("commit", it, self.bar_r)]
            if it + 1 < self.nC:
                self.pipe += [("g1", it + 1), ("commit", it + 1, self.g1_ready[(it + 1) & 1])]
            yield lambda: self.bar_r.done(it & 1)
            self.pipe += [("b2", it), ("commit", it, self.bar_z)]
            yield lambda: self.bar_z.done(it & 1)
            self.pipe += [("b3", it)]
            yield lambda: self.z_done.done(it & 1)
            self.pipe += [("g2", it), ("commit", it, self.g2_ready)]
            yield lambda: self.c_done.done(it & 1)
            yield lambda: self.s0_full.done(it & 1)
            self.pipe += [("b5", it), ("commit", it, self.out_bar(it)[0])]

    def group_c1(self):
        nw = 0
        for it in range(self.nC):
            si = it % NS
            if self.win_last(it):
                yield lambda: self.full[si].done((it // NS) & 1)
                if it > 0:
                    yield lambda nw=nw: self.glp_done.done((nw - 1) & 1)
                assert self.dS == it and self.dST == it, ("rescale before the previous chunk is final", it, self.dS, self.dST)
                assert it == 0 or self.glp_taken == nw, ("rescale before the boundary term was taken", it)
                self.framed += 1
                yield None
                self.resc.arrive(4)
                nw += 1
            yield lambda: self.bar_z.done(it & 1)
            assert self.Z == it, ("C1 reads Z of", self.Z, "for", it)
            assert self.slot_a[si] == it
            self.ztiles = it
            yield None
            self.z_done.arrive(4)

    def gram_g(self):
        for it in range(self.nC):
            yield lambda: self.g2_ready.done(it & 1)
            assert self.G2 == it
            self.G2_read = it
            yield None
            self.gtiles = it
            yield None
            self.c_done.arrive(2)

    def group_c2(self):
        for it in range(self.nC):
            si, u = it % NS, it & 1
            c = self.nC - 1 - it
            yield lambda: self.out_bar(it)[0].done(self.out_bar(it)[1])
            assert self.OK[u] == it, ("C2 drains", self.OK[u], "for", it)
            assert self.slot_busy[si] == 0, ("output staging overwrites tiles still being read", it)
            self.OK_read[u] = it
            yield None
            if c % WIN == 0 and c > 0:       # boundary term from dS^T before C1 rescales it
                assert self.dST == it + 1
                self.glp_taken += 1
                self.glp_done.arrive(8)
            self.ok_free[u].arrive(8)
            yield None
            self.empty[si].arrive(8)
            self.done_out += 1


def run(nC, seed, single_out_ready=False):
    rng = random.Random(seed)
    k = Kernel(nC, single_out_ready)
    roles = {"A": k.stage_a(), "GF": k.gram_f(), "M": k.mma(), "C1": k.group_c1(), "GG": k.gram_g(), "C2": k.group_c2()}
    waiting = {n: None for n in roles}
    alive = set(roles)
    victim = rng.choice(list(roles) + ["pipe", "copy", None, None])
    steps = 0
    while alive or k.pipe or k.copies:
        runnable = [n for n in alive if waiting[n] is None or waiting[n]()]
        if k.pipe:
            runnable.append("pipe")
        if k.copies:
            runnable.append("copy")
        assert runnable, f"deadlock: {sorted(alive)} after {steps} steps (nC={nC}, seed={seed})"
        others = [n for n in runnable if n != victim]
        n = rng.choice(others) if others and rng.random() > 0.03 else rng.choice(runnable)
        if n == "pipe":
            k.pipe_step()
        elif n == "copy":
            k.copy_step()
        else:
            try:
                waiting[n] = next(roles[n])
            except StopIteration:
                alive.discard(n)
        steps += 1
    assert k.done_out == nC and k.dS == nC and k.dST == nC
    return steps


MUTATIONS = [
    ("no z_done wait before G2", "            yield lambda: self.z_done.done(it & 1)\n", ""),
    ("no c_done wait before B5", "            yield lambda: self.c_done.done(it & 1)\n", ""),
    ("no s0_full wait before B5", "            yield lambda: self.s0_full.done(it & 1)\n", ""),
    ("no ok_free wait", "            if it >= 2:\n                yield lambda: self.ok_free[u].done(((it >> 1) - 1) & 1)\n", ""),
    ("no bar_r wait before B2", "            yield lambda: self.bar_r.done(it & 1)\n", ""),
    ("no glp_done wait before the rescale", "                if it > 0:\n                    yield lambda nw=nw: self.glp_done.done((nw - 1) & 1)\n", ""),
    ("no resc wait", "            if self.win_last(it):\n                yield lambda nw=nw: self.resc.done(nw & 1)\n                nw += 1\n", ""),
    ("stage A ignores empty[]", "            if it >= NS:\n                yield lambda: self.empty[si].done(((it // NS) - 1) & 1)\n", ""),
    ("no out_ready wait in the MMA warp", "            if it > 0:\n                yield lambda: self.out_bar(it - 1)[0].done(self.out_bar(it - 1)[1])\n", ""),
    ("G1(it+1) issued before the wait on full[it]", "            yield lambda: self.full[si].done((it // NS) & 1)\n            yield lambda: self.blob_full",
     "            if it + 1 < self.nC and it > 0:\n                self.pipe += [(\"g1\", it + 1), (\"commit\", it + 1, self.g1_ready[(it + 1) & 1])]\n"
     "            yield lambda: self.full[si].done((it // NS) & 1)\n            yield lambda: self.blob_full"),
]


def mutations():
    src = open(__file__).read()
    body = src[:src.index("MUTATIONS = [")]
    for name, old, new in MUTATIONS:
        assert old in body, name
        ns = {"__name__": "mutant", "__file__": __file__}
        exec(compile(body.replace(old, new), name, "exec"), ns)
        caught = None
        try:
            for nC in (5, 9, 13):
                for seed in range(150):
                    ns["run"](nC, seed)
        except (AssertionError, IndexError) as e:
            caught = str(e)[:90]
        print(f"  {name:46s} {'caught: ' + caught if caught else 'NOT caught'}")
    try:
        for seed in range(200):
            run(6, seed, single_out_ready=True)
        print("  single out_ready barrier                       NOT caught")
    except AssertionError as e:
        print("  single out_ready barrier                       caught:", str(e)[:80])


if __name__ == "__main__":
    if "--mutations" in sys.argv:
        mutations()
        sys.exit(0)
    total = 0
    for nC in (1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 16, 23):
        for seed in range(300):
            total += run(nC, seed)
    print("backward-v2 synchronisation design: no deadlock, no hazard over", total, "scheduled steps")
