"""Model check of the forward-v2 sketch's synchronisation (proto/wkv7_tc_fwd_v2.cu, inference variant): the warp roles are
coroutines that execute the kernel's mbarrier waits / arrivals / commits in program order, the tensor pipe is an in-order
queue whose commits arrive when everything issued before them has executed, and a random scheduler interleaves them.
Both variants (the training one adds the U^T tile hand-off, the lagged transposed-state update and the checkpoint
group).  Every access carries an assertion about WHAT it must find (which chunk's data a tile / tensor-memory buffer holds, and
that its previous contents have been consumed), so the run fails on a lost hand-off, an overwrite-before-read, a wrong
parity or a deadlock.  Barrier counts, parities and the issue order are transcribed from the .cu file.
`--mutations` seeds known protocol bugs into a copy of the model and reports which ones it catches (its sensitivity).
Not modelled: the tensor pipe is taken to complete instructions in issue order, so the hardware rule that an MMA reading
tensor memory written under a different accumulator needs a commit + wait in between is NOT checked here."""
import random
import sys

NSLOT, WIN = 5, 4


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self, n=1):
        for _ in range(n):
            self.pending -= 1
            assert self.pending >= 0
            if self.pending == 0:
                self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):                 # try_wait.parity: the phase with this parity has completed
        return (self.phase & 1) != parity


class Kernel:
    def __init__(self, nC, train=False):
        self.nC, self.train = nC, train
        self.ut_ready, self.st_ready, self.st_free = Bar(4), Bar(1), Bar(4)
        self.Ut = [-1, -1]                  # chunk whose U^T tile is in shared memory
        self.Ut_read = [-1, -1]
        self.ST = 0                         # chunks accumulated into the transposed state
        self.ck_out = 0                     # checkpoints read out of S^T (checkpoint 0 = initial state)
        self.empty = [Bar(1) for _ in range(NSLOT)]
        self.full = [Bar(2) for _ in range(NSLOT)]
        self.a_done = [Bar(8) for _ in range(NSLOT)]
        self.g_ready = [Bar(1), Bar(1)]
        self.p_done = Bar(1)
        self.y_ready, self.y_free = [Bar(1), Bar(1)], [Bar(4), Bar(4)]
        self.win_scaled = Bar(4)
        # what each resource holds
        self.slot_a = [-1] * NSLOT          # chunk whose stage-A tiles are in the slot
        self.slot_g = [-1] * NSLOT          # chunk whose Gram-group tiles (Aak/Aqk/Aqb/T) are in the slot
        self.slot_reads_left = [0] * NSLOT  # tensor-pipe reads of the slot still outstanding
        self.G = [-1, -1]                   # chunk whose Gram blocks are in tensor memory
        self.G_read = [-1, -1]              # last chunk whose Gram blocks have been read out
        self.S = 0                          # number of chunks accumulated into S^
        self.windows_scaled = 0             # window-end rescales applied (+1 for the initial state store)
        self.Z, self.U, self.Y = [-1, -1], [-1, -1], [-1, -1]
        self.Y_read = [-1, -1]
        self.pipe = []                      # in-order tensor pipe
        self.y_out = 0

    # ---- tensor pipe -----------------------------------------------------------------------------------
    def pipe_step(self):
        op = self.pipe.pop(0)
        kind, c = op[0], op[1]
        si, u = c % NSLOT, c & 1
        if kind == "commit":
            op[2].arrive()
        elif kind == "gram":
            assert self.slot_a[si] == c, ("gram reads stage-A tiles of", self.slot_a[si], "for", c)
            assert self.G_read[u] == self.G[u], ("gram overwrites unread Gram blocks", self.G[u])
            self.G[u] = c
            self.slot_reads_left[si] -= 1
        elif kind == "p1":
            assert self.slot_a[si] == c and self.slot_g[si] == c, ("p1", c, self.slot_a[si], self.slot_g[si])
            assert self.S == c, ("p1 reads S^ after", self.S, "chunks for", c)
            assert self.windows_scaled == c // WIN + (1 if c % WIN else 1), ("p1 frame", c, self.windows_scaled)
            assert self.Y_read[u] == self.Y[u], ("p1 overwrites unread Y", self.Y[u])
            self.Z[u] = c
            self.Y[u] = (c, "partial")
            self.slot_reads_left[si] -= 1
        elif kind == "p1b":
            assert self.Z[u] == c and self.slot_g[si] == c
            self.U[u] = c
            self.slot_reads_left[si] -= 1
        elif kind == "p2s":
            assert self.U[u] == c and self.S == c and self.slot_a[si] == c
            self.S = c + 1
            self.slot_reads_left[si] -= 1
        elif kind == "p2y":
            assert self.U[u] == c and self.Y[u] == (c, "partial") and self.slot_g[si] == c
            self.Y[u] = c
            self.slot_reads_left[si] -= 1
        elif kind == "st":                  # S^T += B~^T U + K~^T V of chunk c (issued in iteration c + 1)
            assert self.slot_a[si] == c and self.Ut[u] == c, ("st", c, self.slot_a[si], self.Ut[u])
            assert self.ST == c and self.ck_out == c + 1, ("st overwrites checkpoint", self.ST, self.ck_out, c)
            self.ST = c + 1
            self.Ut_read[u] = c
            self.slot_reads_left[si] -= 1

    # ---- roles (generators yield a predicate to wait on, or None to just yield the processor) --------------
    def stage_a(self):
        for c in range(self.nC):
            si = c % NSLOT
            if c >= NSLOT:
                yield lambda si=si, c=c: self.empty[si].done(((c // NSLOT) - 1) & 1)
            assert self.slot_reads_left[si] == 0, ("stage A overwrites a slot still being read", c)
            assert self.slot_g[si] in (-1, c - NSLOT), ("slot", si, "holds Gram tiles of", self.slot_g[si])
            self.slot_a[si] = c
            self.slot_reads_left[si] = 5 + (1 if self.train and c + 1 < self.nC else 0)   # gram, p1, p1b, p2s, p2y (+ st)
            yield None
            self.a_done[si].arrive(8)

    def gram_group(self, grp):
        for c in range(grp, self.nC, 2):
            si = c % NSLOT
            yield lambda c=c: self.g_ready[grp].done((c >> 1) & 1)
            assert self.G[grp] == c, ("gram group", grp, "reads blocks of", self.G[grp], "for", c)
            self.G_read[grp] = c
            yield None
            assert self.slot_a[si] == c, ("gram group writes into a slot that holds", self.slot_a[si])
            self.slot_g[si] = c
            yield None
            self.full[si].arrive(2)

    def mma_warp(self):
        ph = 0
        for c in range(min(2, self.nC)):
            yield lambda c=c: self.a_done[c % NSLOT].done(0)
            self.pipe += [("gram", c), ("commit", c, self.g_ready[c & 1])]
        for c in range(self.nC):
            si, u = c % NSLOT, c & 1
            yield lambda: self.full[si].done((c // NSLOT) & 1)
            if c % WIN == 0:
                yield lambda: self.win_scaled.done((c // WIN) & 1)
            if c >= 2:
                yield lambda: self.y_free[u].done(((c >> 1) - 1) & 1)
            if c + 2 < self.nC:
                yield lambda: self.a_done[(c + 2) % NSLOT].done(((c + 2) // NSLOT) & 1)
            self.pipe += [("p1", c), ("commit", c, self.p_done)]
            if c + 2 < self.nC:
                self.pipe += [("gram", c + 2), ("commit", c + 2, self.g_ready[(c + 2) & 1])]
            yield lambda ph=ph: self.p_done.done(ph)
            ph ^= 1
            self.pipe += [("p1b", c), ("commit", c, self.p_done)]
            yield lambda ph=ph: self.p_done.done(ph)
            ph ^= 1
            if self.train and c > 0:
                yield lambda: self.ut_ready.done((c - 1) & 1)
                yield lambda: self.st_free.done((c - 1) & 1)
            self.pipe += [("p2s", c), ("commit", c, self.p_done), ("p2y", c)]
            if not self.train:
                self.pipe += [("commit", c, self.empty[si])]
            self.pipe += [("commit", c, self.y_ready[u])]
            if self.train and c > 0:
                self.pipe += [("st", c - 1), ("commit", c, self.st_ready), ("commit", c, self.empty[(c - 1) % NSLOT])]
            yield lambda ph=ph: self.p_done.done(ph)
            ph ^= 1

    def epilogue(self):
        self.windows_scaled = 1                     # initial state stored
        yield None
        self.win_scaled.arrive(4)
        for c in range(self.nC):
            u = c & 1
            last = c == self.nC - 1
            win_end = (c % WIN == WIN - 1) or last
            yield lambda: self.y_ready[u].done((c >> 1) & 1)
            assert self.Y[u] == c, ("epilogue reads Y of", self.Y[u], "for", c)
            self.Y_read[u] = c
            assert self.U[u] == c
            if win_end:
                assert self.S == c + 1, ("window-end rescale sees", self.S, "chunks at", c)
                if not last:
                    self.windows_scaled += 1
            yield None
            self.y_free[u].arrive(4)
            if win_end:
                self.win_scaled.arrive(4)
            if self.train:
                assert self.Ut_read[u] == self.Ut[u], ("epilogue overwrites an unread U^T tile", self.Ut[u])
                self.Ut[u] = c
                yield None
                self.ut_ready.arrive(4)
            self.y_out += 1

    def ckpt_group(self):
        self.ck_out = 1                             # checkpoint 0 = the initial state
        yield None
        self.st_free.arrive(4)
        for c in range(self.nC - 1):
            yield lambda c=c: self.st_ready.done(c & 1)
            assert self.ST == c + 1, ("checkpoint", c + 1, "reads S^T after", self.ST, "chunks")
            self.ck_out = c + 2
            yield None
            self.st_free.arrive(4)


def run(nC, seed, train=False):
    rng = random.Random(seed)
    k = Kernel(nC, train)
    roles = {"A": k.stage_a(), "G0": k.gram_group(0), "G1": k.gram_group(1), "M": k.mma_warp(), "E": k.epilogue()}
    if train:
        roles["C"] = k.ckpt_group()
    waiting = {n: None for n in roles}
    alive = set(roles)
    steps = 0
    # adversarial scheduling: in most runs one role (or the tensor pipe) is starved -- it only runs when nothing else can,
    # or with a small probability -- so that every "the other side is surely done by then" assumption gets exercised
    victim = rng.choice(list(roles) + ["pipe", None, None])
    while alive or k.pipe:
        runnable = [n for n in alive if waiting[n] is None or waiting[n]()]
        if k.pipe:
            runnable.append("pipe")
        assert runnable, f"deadlock: waiting roles {sorted(alive)} after {steps} steps (nC={nC}, seed={seed})"
        others = [n for n in runnable if n != victim]
        n = rng.choice(others) if others and rng.random() > 0.03 else rng.choice(runnable)
        if n == "pipe":
            k.pipe_step()
        else:
            try:
                waiting[n] = next(roles[n])
            except StopIteration:
                alive.discard(n)
        steps += 1
    assert k.y_out == nC and k.S == nC and (not train or (k.ck_out == nC and k.ST == nC - 1))
    return steps


MUTATIONS = [
    ("Gram(c+2) issued before the wait on full[c]",
     "            yield lambda: self.full[si].done((c // NSLOT) & 1)\n",
     "            if c + 2 < self.nC:\n                self.pipe += [(\"gram\", c + 2), (\"commit\", c + 2, self.g_ready[(c + 2) & 1])]\n"
     "            yield lambda: self.full[si].done((c // NSLOT) & 1)\n"),
    ("no y_free wait", "            if c >= 2:\n                yield lambda: self.y_free[u].done(((c >> 1) - 1) & 1)\n", ""),
    ("full[] expects one arrival", "self.full = [Bar(2) for _ in range(NSLOT)]", "self.full = [Bar(1) for _ in range(NSLOT)]"),
    ("stage A ignores empty[]", "            if c >= NSLOT:\n                yield lambda si=si, c=c: self.empty[si].done(((c // NSLOT) - 1) & 1)\n", ""),
    ("no win_scaled wait", "            if c % WIN == 0:\n                yield lambda: self.win_scaled.done((c // WIN) & 1)\n", ""),
    ("g_ready parity off by one", "self.g_ready[grp].done((c >> 1) & 1)", "self.g_ready[grp].done(((c >> 1) + 1) & 1)"),
    ("no a_done wait for chunk c+2",
     "            if c + 2 < self.nC:\n                yield lambda: self.a_done[(c + 2) % NSLOT].done(((c + 2) // NSLOT) & 1)\n", ""),
    ("slot never released (inference)", '                self.pipe += [("commit", c, self.empty[si])]\n', "                pass\n"),
    ("training: no st_free wait", "                yield lambda: self.st_free.done((c - 1) & 1)\n", ""),
    ("training: no ut_ready wait", "                yield lambda: self.ut_ready.done((c - 1) & 1)\n", ""),
    ("training: slot released in phase 2 as in inference", "            if not self.train:\n                self.pipe += [(\"commit\", c, self.empty[si])]\n",
     "            self.pipe += [(\"commit\", c, self.empty[si])]\n"),
]


def mutations():
    src = open(__file__).read()
    body = src[:src.index("MUTATIONS = [")]
    for name, old, new in MUTATIONS:
        assert old in body, name
        ns = {}
        exec(compile(body.replace(old, new), name, "exec"), ns)
        caught = None
        try:
            for nC in (4, 7, 13):
                for seed in range(200):
                    ns["run"](nC, seed)
                    ns["run"](nC, seed, train=True)
        except AssertionError as e:
            caught = str(e)[:100]
        print(f"  {name:46s} {'caught: ' + caught if caught else 'NOT caught'}")


if __name__ == "__main__":
    if "--mutations" in sys.argv:
        mutations()
        sys.exit(0)
    total = 0
    for nC in (1, 2, 3, 4, 5, 6, 7, 9, 13, 16, 23, 64):
        for seed in range(300 if nC < 30 else 40):
            total += run(nC, seed) + run(nC, seed, train=True)
    print("forward-v2 synchronisation model: no deadlock, no hazard over", total, "scheduled steps")
