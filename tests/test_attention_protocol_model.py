"""Model checks of the attention kernel (gigaam_b200/csrc/attention_sm100.cu) on the CPU: its barrier protocol (first part)
and the arithmetic of its lazy-reference online softmax (second part, at the end of the file).

The kernel is six concurrent roles -- TMA producer, two MMA issuer warps, two softmax warpgroups, one output warpgroup --
that hand shared-memory tiles (Q, K / V sets) and tensor-memory regions (S, P, O per stream) to each other through
mbarriers, with work items whose shape varies (packed rows: 0 .. 6 query tiles / key blocks per utterance).  A GPU test
exercises the few interleavings the hardware happens to produce; this file restates every role's control flow as a Python
coroutine (same loops, same waits, same phase arithmetic, line for line) and runs them under a RANDOM scheduler over
thousands of item sequences, with asynchronous engines for the TMA loads and the tensor-core queue.  It checks what a
hang or a silent corruption on the GPU would mean:

  * no deadlock: every role finishes;
  * no phase aliasing: a wait for completion j of a barrier is only ever evaluated when the barrier has completed j or
    j + 1 times (parity waits cannot tell j + 1 from j + 3), and the arrivals of one phase come from distinct parties;
  * no data hazard: an S block is not overwritten before the softmax has read it, a P block not before its P.V has
    executed, O not before the output warpgroup has read it, a Q tile / K-V set not while MMAs that read it are pending,
    and every MMA reads the tiles of ITS item (versions are tagged).

The model is test infrastructure for the kernel's design: when the kernel's protocol changes, this file changes with it.
"""
import random

import pytest

K_MAX_KB = 6


class ProtocolError(AssertionError):
    pass


class Barrier:
    """mbarrier with `count` expected arrivals per phase; completions are counted instead of keeping a parity bit, so
    that a wait can state which completion it means and aliasing becomes visible."""

    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completions, self.parties = name, count, count, 0, []

    def arrive(self, who):
        if who in self.parties:
            raise ProtocolError(f"{self.name}: second arrival of {who} within one phase")
        self.parties.append(who)
        self.pending -= 1
        if self.pending == 0:
            self.completions += 1
            self.pending = self.count
            self.parties = []

    def ready(self, j):
        """wait for completion j (0-based); j = -1 is the 'preceding phase' of a fresh barrier"""
        if self.completions > j + 1:
            raise ProtocolError(f"{self.name}: waiting for completion {j} but {self.completions} have happened (parity aliasing)")
        return self.completions == j + 1


class Sim:
    def __init__(self, items, nkb, ring, rng, look_ring=None, unbounded_look=False):
        self.items, self.nkb, self.ring, self.rng = items, nkb, ring, rng          # items: list of (nq, nk)
        self.look_ring = ring if look_ring is None else look_ring                  # what the `look` test believes (sanity tests)
        self.unbounded_look = unbounded_look
        B = Barrier
        self.kv_full = [[B(f"kv_full[{s}][{k}]", 1) for k in range(K_MAX_KB)] for s in range(2)]
        self.kv_empty = [B(f"kv_empty[{s}]", 2) for s in range(2)]
        self.q_full = [B(f"q_full[{w}]", 1) for w in range(2)]
        self.q_empty = [B(f"q_empty[{w}]", 1) for w in range(2)]
        self.s_full = [B(f"s_full[{w}]", 1) for w in range(2)]
        self.s_empty = [B(f"s_empty[{w}]", 1) for w in range(2)]     # 4 warps in the kernel: one party here
        self.p_full = [B(f"p_full[{w}]", 1) for w in range(2)]
        self.p_empty = [B(f"p_empty[{w}]", 1) for w in range(2)]
        self.o_full = [B(f"o_full[{w}]", 1) for w in range(2)]
        self.o_empty = [B(f"o_empty[{w}]", 1) for w in range(2)]
        # tagged resources
        self.q_tile = [None, None]                      # (item, qt) or "loading"
        self.kv = [[None] * K_MAX_KB for _ in range(2)]  # item index or "loading"
        self.v_ready = [[set() for _ in range(K_MAX_KB)] for _ in range(2)]   # streams that prepared V (ones / zero tail)
        self.s_reg = [None, None]                       # (item, qt, kb) written and not yet read
        self.p_reg = [None, None]
        self.o_reg = [None, None]                       # [item, qt, blocks accumulated]
        self.tma = []                                   # pending loads: callables
        self.mma_q = [[], []]                           # per issuer warp: in-order queue of ops
        self.pending_reads = {"q": [0, 0], "kv": [0, 0]}   # MMAs queued that read a Q tile / a K-V set

    # ------------------------------------------------------------------ asynchronous engines
    def tma_load(self, fn):
        self.tma.append(fn)

    def enqueue(self, wg, op):
        self.mma_q[wg].append(op)

    # ------------------------------------------------------------------ roles (generators yield predicates to wait on)
    def producer(self):
        it, n_q = 0, [0, 0]
        for item, (nq, nk) in enumerate(self.items):
            if nq == 0:
                continue
            st, use = it % self.ring, it // self.ring
            it += 1
            yield lambda st=st, use=use: self.kv_empty[st].ready(use - 1)
            if self.pending_reads["kv"][st]:
                raise ProtocolError("K / V set reloaded while MMAs that read it are queued")
            for kb in range(self.nkb):
                if kb < nk:
                    self.kv[st][kb] = "loading"
                    self.v_ready[st][kb] = set()

                    def done(st=st, kb=kb, item=item):
                        self.kv[st][kb] = item
                        self.kv_full[st][kb].arrive("tma")
                    self.tma_load(done)
                else:
                    self.kv_full[st][kb].arrive("producer")
            for qt in range(nq):
                wg = qt & 1
                yield lambda wg=wg, j=n_q[wg] - 1: self.q_empty[wg].ready(j)
                if self.pending_reads["q"][wg]:
                    raise ProtocolError("Q tile reloaded while MMAs that read it are queued")
                self.q_tile[wg] = "loading"

                def qdone(wg=wg, item=item, qt=qt):
                    self.q_tile[wg] = (item, qt)
                    self.q_full[wg].arrive("tma")
                self.tma_load(qdone)
                n_q[wg] += 1

    def mma(self, wg):
        n_q = n_s = n_p = 0
        state = {"it": 0, "item": 0, "qt_next": 0, "cur_nq": 0, "cur_nk": 0, "cur_set": 0, "cur_use": 0, "cur_item": -1}
        n_items = len(self.items)

        def g_ahead_nq():
            return self.items[state["item"]][0] if state["item"] < n_items else 0

        def next_tile():
            while True:
                if state["qt_next"] < state["cur_nq"]:
                    qt = state["qt_next"]
                    t = dict(valid=1, nk=state["cur_nk"], set=state["cur_set"], first_q=qt == wg, last_q=qt + 2 >= state["cur_nq"],
                             use=state["cur_use"], item=state["cur_item"], qt=qt)
                    state["qt_next"] += 2
                    return t
                if state["item"] >= n_items:
                    return dict(valid=0)
                item = state["item"]
                nq, nk = self.items[item]
                state["item"] += 1
                state["cur_nq"] = 0
                if nq == 0:
                    continue
                state["cur_set"], state["cur_use"] = state["it"] % self.ring, state["it"] // self.ring
                state["it"] += 1
                if wg >= nq:
                    st, use = state["cur_set"], state["cur_use"]
                    yield lambda st=st, use=use: self.kv_full[st][0].ready(use)
                    self.kv_empty[st].arrive(f"stream{wg}")
                    continue
                state["cur_nq"], state["cur_nk"], state["qt_next"], state["cur_item"] = nq, nk, wg, item

        def issue_s(t, kb):
            nonlocal n_s
            yield lambda j=n_s - 1: self.s_empty[wg].ready(j)
            if t["first_q"]:
                yield lambda: self.kv_full[t["set"]][kb].ready(t["use"])
            self.pending_reads["q"][wg] += 1
            self.pending_reads["kv"][t["set"]] += 1

            def run(t=t, kb=kb):
                if self.q_tile[wg] != (t["item"], t["qt"]):
                    raise ProtocolError(f"S MMA of stream {wg} read Q tile {self.q_tile[wg]}, wanted {(t['item'], t['qt'])}")
                if self.kv[t["set"]][kb] != t["item"]:
                    raise ProtocolError(f"S MMA read K block of item {self.kv[t['set']][kb]}, wanted {t['item']}")
                if self.s_reg[wg] is not None:
                    raise ProtocolError(f"S region of stream {wg} overwritten before the softmax read {self.s_reg[wg]}")
                self.s_reg[wg] = (t["item"], t["qt"], kb)
                self.pending_reads["q"][wg] -= 1
                self.pending_reads["kv"][t["set"]] -= 1
            self.enqueue(wg, run)
            self.enqueue(wg, lambda: self.s_full[wg].arrive("commit"))
            if kb == t["nk"] - 1:
                self.enqueue(wg, lambda: self.q_empty[wg].arrive("commit"))
            n_s += 1

        def issue_pv(t, pb):
            nonlocal n_p
            yield lambda j=n_p: self.p_full[wg].ready(j)
            if pb == 0:
                yield lambda j=n_q - 1: self.o_empty[wg].ready(j)
            self.pending_reads["kv"][t["set"]] += 1

            def run(t=t, pb=pb):
                if self.p_reg[wg] != (t["item"], t["qt"], pb):
                    raise ProtocolError(f"P.V read P block {self.p_reg[wg]}, wanted {(t['item'], t['qt'], pb)}")
                if self.kv[t["set"]][pb] != t["item"] or wg not in self.v_ready[t["set"]][pb]:
                    raise ProtocolError("P.V read a V block of another item / before its tail and ones column were written")
                if pb == 0:
                    if self.o_reg[wg] is not None:
                        raise ProtocolError(f"O of stream {wg} overwritten before it was read out: {self.o_reg[wg]}")
                    self.o_reg[wg] = [t["item"], t["qt"], 1]
                else:
                    if self.o_reg[wg] != [t["item"], t["qt"], pb]:
                        raise ProtocolError(f"P.V accumulates into O {self.o_reg[wg]}, wanted {[t['item'], t['qt'], pb]}")
                    self.o_reg[wg][2] += 1
                self.p_reg[wg] = None
                self.pending_reads["kv"][t["set"]] -= 1
            self.enqueue(wg, run)
            self.enqueue(wg, lambda: self.p_empty[wg].arrive("commit"))
            if pb == t["nk"] - 1:
                self.enqueue(wg, lambda: self.o_full[wg].arrive("commit"))
                if t["last_q"]:
                    self.enqueue(wg, lambda st=t["set"]: self.kv_empty[st].arrive(f"stream{wg}"))
            n_p += 1

        t = yield from next_tile()
        s0_issued = False
        while t["valid"]:
            if not s0_issued:
                yield lambda j=n_q: self.q_full[wg].ready(j)
                yield from issue_s(t, 0)
            for kb in range(1, t["nk"]):
                yield from issue_s(t, kb)
                yield from issue_pv(t, kb - 1)
            n = dict(valid=0)
            look = state["qt_next"] < state["cur_nq"] or (self.look_ring == 2 and state["item"] < n_items and
                                                          (g_ahead_nq() > wg or self.unbounded_look))
            s0_issued = False
            if look:
                n = yield from next_tile()
                if n["valid"]:
                    yield lambda j=n_q + 1: self.q_full[wg].ready(j)
                    yield from issue_s(n, 0)
                    s0_issued = True
            yield from issue_pv(t, t["nk"] - 1)
            n_q += 1
            if not look:
                n = yield from next_tile()
            t = n

    def softmax(self, wg):
        it = n_s = 0
        for item, (nq, nk) in enumerate(self.items):
            if nq == 0:
                continue
            st = it % self.ring
            it += 1
            for qt in range(wg, nq, 2):
                for kb in range(nk):
                    yield lambda j=n_s: self.s_full[wg].ready(j)
                    if self.s_reg[wg] != (item, qt, kb):
                        raise ProtocolError(f"softmax of stream {wg} read S {self.s_reg[wg]}, wanted {(item, qt, kb)}")
                    if qt == wg:
                        if self.kv[st][kb] != item:
                            raise ProtocolError("V tail / ones column written into a tile of another item")
                        self.v_ready[st][kb].add(wg)
                    self.s_reg[wg] = None                      # block is in registers
                    self.s_empty[wg].arrive("softmax")
                    yield lambda j=n_s - 1: self.p_empty[wg].ready(j)
                    if kb > 0 and self.o_reg[wg] != [item, qt, kb]:       # a rescale would touch O here
                        raise ProtocolError(f"O of stream {wg} is {self.o_reg[wg]} when block {kb} of {(item, qt)} may rescale it")
                    if self.p_reg[wg] is not None:
                        raise ProtocolError("P block overwritten before its P.V executed")
                    self.p_reg[wg] = (item, qt, kb)
                    self.p_full[wg].arrive("softmax")
                    n_s += 1

    def output(self):
        n_o = [0, 0]
        for item, (nq, nk) in enumerate(self.items):
            for qt in range(nq):
                wg = qt & 1
                yield lambda wg=wg, j=n_o[wg]: self.o_full[wg].ready(j)
                n_o[wg] += 1
                if self.o_reg[wg] != [item, qt, nk]:
                    raise ProtocolError(f"output warpgroup read O {self.o_reg[wg]}, wanted {[item, qt, nk]}")
                self.o_reg[wg] = None
                self.o_empty[wg].arrive("output")

    # ------------------------------------------------------------------ scheduler
    def run(self, max_steps=2_000_000):
        roles = {"producer": self.producer(), "mma0": self.mma(0), "mma1": self.mma(1), "softmax0": self.softmax(0),
                 "softmax1": self.softmax(1), "output": self.output()}
        waiting = {}
        for name, gen in list(roles.items()):
            try:
                waiting[name] = next(gen)
            except StopIteration:
                del roles[name]
        for _ in range(max_steps):
            choices = [("role", n) for n in roles if waiting[n]()]
            if self.tma:
                choices.append(("tma", None))
            choices += [("mma", w) for w in range(2) if self.mma_q[w]]
            if not choices:
                if roles:
                    raise ProtocolError(f"deadlock: {sorted(roles)} blocked")
                return
            kind, who = self.rng.choice(choices)
            if kind == "tma":
                self.tma.pop(self.rng.randrange(len(self.tma)))()      # loads complete in any order
            elif kind == "mma":
                self.mma_q[who].pop(0)()                                # in order per issuing warp
            else:
                try:
                    waiting[who] = roles[who].send(None)
                except StopIteration:
                    del roles[who], waiting[who]
        raise ProtocolError("did not finish")


def _random_items(rng, nkb, packed):
    n = rng.randint(1, 9)
    items = []
    for _ in range(n):
        if packed:
            kb = rng.choice([0, 1, 1, 2, nkb, rng.randint(0, nkb)])
            kb = min(kb, nkb)
            items.append((kb, max(1, kb)))             # query tiles = key blocks = ceil(len / 128); empty utterance: skipped
        else:
            items.append((nkb, rng.randint(1, nkb)))   # padded layout of the unit tests: every query tile, keys by length
    return items


@pytest.mark.parametrize("nkb", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("packed", [True, False])
def test_attention_barrier_protocol_has_no_deadlock_aliasing_or_hazard(nkb, packed):
    ring = 2 if nkb <= 3 else 1                         # launch_attention's choice
    rng = random.Random(1000 * nkb + packed)
    for trial in range(400):
        items = _random_items(rng, nkb, packed)
        Sim(items, nkb, ring, rng).run()


def test_the_model_catches_a_lookahead_that_crosses_a_shared_buffer_set():
    """Sanity of the checker itself: with a single K / V set, letting the MMA warp look ahead across an item boundary (what the
    kernel's `look` condition forbids) must show up as a deadlock -- the next item's loads wait for this tile's last P.V, which
    waits behind the look-ahead."""
    rng = random.Random(7)
    with pytest.raises(ProtocolError, match="deadlock"):
        for _ in range(50):
            Sim([(2, 2), (2, 2), (2, 2)], nkb=2, ring=1, rng=rng, look_ring=2).run()


def test_the_model_catches_an_unbounded_lookahead_with_two_sets():
    """... and with two sets: a stream that skips over an item it has no tile in and opens the item AFTER it is back on the
    current item's K / V set (the kernel bounds the look-ahead to the next item that has a tile for the stream)."""
    rng = random.Random(11)
    with pytest.raises(ProtocolError, match="deadlock"):
        for _ in range(200):
            Sim([(2, 2), (1, 1), (2, 2), (1, 1), (2, 2)], nkb=2, ring=2, rng=rng, unbounded_look=True).run()


# ------------------------------------------------------------------------------------------ numerics of the same kernel
def _lazy_online_attention(q, k, v, klen, lazy=8.0, block=128):
    """The arithmetic of attention_kernel restated with torch on the CPU: key blocks of 128, a reference point that the first
    block sets to its exact maximum and later blocks move only when they exceed it by more than 2^lazy (O and its row-sum
    column are rescaled then), P rounded to fp16 before it meets V, fp32 accumulation, and the softmax denominator taken as
    the row sum of the ROUNDED P (the ones column of V on the tensor core)."""
    import torch
    T, dk = q.shape
    scale = 1.4426950408889634 / dk ** 0.5
    out = torch.zeros(T, v.shape[1])
    acc = torch.zeros(T, v.shape[1])
    den = torch.zeros(T)
    mc = torch.zeros(T)
    pmax = 0.0
    moves = 0
    for kb in range((klen + block - 1) // block):
        lo, hi = kb * block, min(klen, (kb + 1) * block)
        s = (q @ k[lo:hi].T) * scale                                   # log2 domain
        bm = s.max(dim=1).values
        if kb == 0:
            mc = bm.clone()
        else:
            move = bm > mc + lazy
            corr = torch.where(move, torch.exp2(mc - bm), torch.ones(T))
            acc, den = acc * corr[:, None], den * corr
            mc = torch.where(move, bm, mc)
            moves += int(move.sum())
        p = torch.exp2(s - mc[:, None]).half().float()
        pmax = max(pmax, float(p.max()))
        acc = acc + p @ v[lo:hi]
        den = den + p.sum(1)
    out = (acc / den[:, None]).half().float()
    return out, pmax, moves


@pytest.mark.parametrize("case", ["randn", "peaked_ramp", "descending", "wide"])
def test_lazy_reference_softmax_is_accurate_and_stays_inside_fp16(case):
    """Why 2^8: P never exceeds 256 (fp16 tops out at 65 504), rows whose maximum sits in a LATER block than the reference
    lose no precision (fp16 keeps 11 bits at every magnitude), and the rounded-P row sum keeps numerator and denominator
    consistent.  Checked against a float64 softmax on score rows built to move the reference, not to move it, and to
    span +-60 in the exponent."""
    import torch
    g = torch.Generator().manual_seed(3)
    T, dk = 300, 48
    q, k, v = (torch.randn(T, dk, generator=g) for _ in range(3))
    if case == "peaked_ramp":
        q = q * 6.0
        k = k * (0.1 + 2.4 * torch.arange(T) / T)[:, None]
    elif case == "descending":
        q = q * 6.0
        k = k * (2.5 - 2.4 * torch.arange(T) / T)[:, None]
    elif case == "wide":
        q = q * 20.0
    q, k, v = q.half().float(), k.half().float(), v.half().float()
    got, pmax, moves = _lazy_online_attention(q, k, v, klen=T)
    want = (torch.softmax((q.double() @ k.double().T) / dk ** 0.5, -1) @ v.double()).float()
    rel = float((got - want).norm() / want.norm())
    assert torch.isfinite(got).all()
    assert pmax <= 256.0 * 1.001                # the lazy reference never trails by more than 2^8
    assert rel < 1e-3, (case, rel)
    if case == "peaked_ramp":
        assert moves > T                        # the case does exercise the rescaling path
    if case == "randn":
        assert moves == 0                       # ... and ordinary rows never pay for it
