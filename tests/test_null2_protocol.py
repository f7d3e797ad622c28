"""Model check of the mbarrier protocol of the K2 null-scan kernel (derfinder_b200/csrc/fstat_tc2.cu).

compute-sanitizer is closed on the GPU pool this was built on (profiles/r2_sanitizer_closed.log), so the
hand-rolled producer / MMA / epilogue hand-off is checked here instead: a host simulation of the kernel's
roles with the hardware's parity semantics (`try_wait.parity(p)` succeeds iff the barrier's current phase
parity differs from p), run under many random interleavings.  Invariants: no deadlock; an accumulator stage is
never overwritten before the epilogue group released its previous use; an A stage or the B group is never
refilled before every MMA that reads it has retired.  The test also shows the bug the sequence hand-over of the
two MMA issuers fixes: without it, an issuer two uses ahead on a shared stage takes the barrier's initial
state for "released" (it hung the first version of the kernel on a GPU)."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def test(self, parity):
        return self.phase != parity


class Violation(Exception):
    pass


def simulate(NT, EW, MW, NA, groups, cg, chunks, tiles, use_seq, seed, max_rounds=200000):
    rnd = random.Random(seed)
    tfull = [Bar(1) for _ in range(NT)]
    tempty = [Bar(1) for _ in range(NT)]  # the four warps of a group arrive together: modelled as one
    afull = [Bar(1) for _ in range(NA)]
    aempty = [Bar(MW) for _ in range(NA)]
    bfull, bempty = Bar(1), Bar(MW)
    st = dict(issue_seq=0, issued=[0] * NT, released=[0] * NT, a_loaded=[0] * NA, a_retired=[0] * NA,
              b_loaded=0, b_retired=0)

    def producer():
        stage, aph = 0, 0
        for g in range(groups):
            if g > 0:
                while not bempty.test((g - 1) & 1):
                    yield
            if st["b_retired"] != g * MW:
                raise Violation("B group replaced before every issuer retired the previous pass")
            st["b_loaded"] += 1
            bfull.arrive()
            for _ in range(tiles):
                while not aempty[stage].test(aph ^ 1):
                    yield
                if st["a_retired"][stage] != st["a_loaded"][stage] * MW:
                    raise Violation("A stage refilled while MMAs may still read it")
                st["a_loaded"][stage] += 1
                afull[stage].arrive()
                stage += 1
                if stage == NA:
                    stage, aph = 0, aph ^ 1
                yield

    def mma(me):
        turn = s = ph = stage = aph = unit = 0
        for g in range(groups):
            cgl = min(cg, chunks - g * cg)
            while not bfull.test(g & 1):
                yield
            for _ in range(tiles):
                while not afull[stage].test(aph):
                    yield
                for _c in range(cgl):
                    mine = turn == me
                    turn = (turn + 1) % MW
                    u = unit
                    unit += 1
                    if mine:
                        if use_seq:
                            while st["issue_seq"] != u:
                                yield
                        while not tempty[s].test(ph ^ 1):
                            yield
                        if st["released"][s] != st["issued"][s]:
                            raise Violation("accumulator stage %d overwritten before its previous use was released" % s)
                        if use_seq:
                            st["issue_seq"] = u + 1
                        yield
                        st["issued"][s] += 1
                        tfull[s].arrive()  # commit: arrives when the MMAs have completed (in order)
                    s += 1
                    if s == NT:
                        s, ph = 0, ph ^ 1
                st["a_retired"][stage] += 1
                aempty[stage].arrive()
                stage += 1
                if stage == NA:
                    stage, aph = 0, aph ^ 1
            st["b_retired"] += 1
            bempty.arrive()

    def epilogue(h):
        turn = s = ph = 0
        for g in range(groups):
            cgl = min(cg, chunks - g * cg)
            for _ in range(tiles):
                for _c in range(cgl):
                    mine = turn == h
                    turn = (turn + 1) % EW
                    sm, pm = s, ph
                    s += 1
                    if s == NT:
                        s, ph = 0, ph ^ 1
                    if not mine:
                        continue
                    while not tfull[sm].test(pm):
                        yield
                    if st["issued"][sm] != st["released"][sm] + 1:
                        raise Violation("epilogue group read a stage that holds no fresh accumulator")
                    yield
                    st["released"][sm] += 1
                    tempty[sm].arrive()

    threads = [producer()] + [mma(m) for m in range(MW)] + [epilogue(h) for h in range(EW)]
    alive = [True] * len(threads)
    idle_rounds = 0
    for _ in range(max_rounds):
        if not any(alive):
            return "done"
        before = (st["issue_seq"], tuple(st["issued"]), tuple(st["released"]), tuple(st["a_loaded"]), st["b_loaded"],
                  tuple(alive))
        order = [i for i in range(len(threads)) if alive[i]]
        rnd.shuffle(order)
        # random subset: some threads stall this round (models arbitrary relative speeds)
        for i in order:
            if rnd.random() < 0.35:
                continue
            try:
                next(threads[i])
            except StopIteration:
                alive[i] = False
        after = (st["issue_seq"], tuple(st["issued"]), tuple(st["released"]), tuple(st["a_loaded"]), st["b_loaded"],
                 tuple(alive))
        idle_rounds = idle_rounds + 1 if after == before else 0
        if idle_rounds > 400:
            return "deadlock"
    return "timeout"


# (NT stages, EW epilogue groups, MW issuers): the geometries null2_geometry() produces
GEOMETRIES = [(3, 3, 2), (4, 4, 2), (2, 2, 2), (4, 2, 2), (3, 3, 1), (4, 4, 1), (2, 1, 2), (1, 1, 1)]


@pytest.mark.parametrize("NT,EW,MW", GEOMETRIES)
def test_protocol_random_interleavings(NT, EW, MW):
    use_seq = MW > 1 and NT % MW != 0
    for seed in range(40):
        for groups, cg, chunks, tiles in ((1, 7, 7, 3), (5, 7, 32, 2), (2, 4, 8, 1), (8, 4, 32, 1), (3, 1, 3, 4)):
            assert simulate(NT, EW, MW, 3, groups, cg, chunks, tiles, use_seq, seed) == "done"
            assert simulate(NT, EW, MW, 2, groups, cg, chunks, tiles, use_seq, seed + 1000) == "done"


def test_protocol_needs_the_issue_sequence_when_stages_are_shared():
    """NT = 3 stages shared by two issuers WITHOUT the sequence hand-over: some interleaving lets an issuer
    overwrite an accumulator that has not been scanned (or deadlocks afterwards)."""
    bad = 0
    for seed in range(300):
        try:
            if simulate(3, 3, 2, 3, 5, 7, 32, 1, False, seed) != "done":
                bad += 1
        except Violation:
            bad += 1
    assert bad > 0
