"""Barrier-protocol model of the multi-query score kernel (6dgs_b200/csrc/score_tc_mq.cu, all three key formats), runnable
on the CPU: the five roles of one CTA pair (key producer x2, token producer x2, MMA issuer, 16 epilogue warps) are
stepped in random interleavings against a model of mbarrier phases, with TMA loads landing late and out of order and
MMA completions (tcgen05.commit arrivals) retiring in issue order.  Checked: every role terminates (no deadlock) and,
at the time an MMA completes, both CTAs' shared-memory slots hold exactly the (tile, query, k-block) operands that MMA
was issued for (no slot is refilled early).  The loop structure mirrors the kernel's; keep the two in step."""
import random

import pytest

class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        assert s.pending>=0
        if s.pending==0: s.phase^=1; s.pending=s.count
    def passed(s, parity):  # try_wait.parity: true if the phase with this parity has completed
        return s.phase != parity
def run(n_tiles, nq, S, seed):
    rnd=random.Random(seed)
    KB=6
    # leader-side barriers (full) and per-CTA empties; model both CTAs
    k_full=[Bar(2) for _ in range(KB)]; q_full=[Bar(2) for _ in range(S)]
    k_empty=[[Bar(1) for _ in range(KB)] for _ in range(2)]; q_empty=[[Bar(1) for _ in range(S)] for _ in range(2)]
    tmem_full=[[Bar(1) for _ in range(2)] for _ in range(2)]; tmem_empty=[Bar(16) for _ in range(2)]
    kt=[[None]*KB for _ in range(2)]; qs=[[None]*S for _ in range(2)]
    pending_mma=[]  # queue of callbacks executed in order when "MMA completes"
    def kprod(c):
        phase=0
        for t in range(n_tiles):
            for kb in range(KB):
                while not k_empty[c][kb].passed(phase^1): yield
                # load lands later
                def land(c=c,kb=kb,t=t): kt[c][kb]=("K",t,kb); k_full[kb].arrive()
                lands.append(land)
                yield
            phase^=1
    def qprod(c):
        stage=0; phase=0
        for t in range(n_tiles):
            for b in range(nq):
                for kb in range(KB):
                    while not q_empty[c][stage].passed(phase^1): yield
                    def land(c=c,stage=stage,t=t,b=b,kb=kb): qs[c][stage]=("Q",t,b,kb); q_full[stage].arrive()
                    lands.append(land)
                    stage+=1
                    if stage==S: stage=0; phase^=1
                    yield
    def mma():
        stage=0; qphase=0; kphase=0; it=0
        for t in range(n_tiles):
            for b in range(nq):
                acc=it&1; ap=(it>>1)&1
                while not tmem_empty[acc].passed(ap^1): yield
                for kb in range(KB):
                    if b==0:
                        while not k_full[kb].passed(kphase): yield
                    while not q_full[stage].passed(qphase): yield
                    # issue MMA: record operand check at completion time
                    def comp(stage=stage,kb=kb,t=t,b=b):
                        for c in range(2):
                            assert kt[c][kb]==("K",t,kb),(kt[c][kb],t,b,kb)
                            assert qs[c][stage]==("Q",t,b,kb),(qs[c][stage],t,b,kb)
                    pending_mma.append(comp)
                    def ce(stage=stage):
                        for c in range(2): q_empty[c][stage].arrive()
                    pending_mma.append(ce)
                    if b==nq-1:
                        def ke(kb=kb):
                            for c in range(2): k_empty[c][kb].arrive()
                        pending_mma.append(ke)
                    stage+=1
                    if stage==S: stage=0; qphase^=1
                    yield
                def tf(acc=acc):
                    for c in range(2): tmem_full[c][acc].arrive()
                pending_mma.append(tf)
                it+=1
            kphase^=1
    done=[0]
    def epi(c,w):
        it=0
        for t in range(n_tiles):
            for b in range(nq):
                acc=it&1; ap=(it>>1)&1
                while not tmem_full[c][acc].passed(ap): yield
                for _ in range(rnd.randint(0,3)): yield
                tmem_empty[acc].arrive()
                it+=1
        done[0]+=1
    lands=[]
    actors=[kprod(0),kprod(1),qprod(0),qprod(1),mma()]+[epi(c,w) for c in range(2) for w in range(8)]
    alive=set(range(len(actors)))
    idle=0
    while alive:
        progressed=False
        i=rnd.choice(sorted(alive))
        try: next(actors[i])
        except StopIteration: alive.discard(i)
        # random async completions, in order
        if lands and rnd.random()<0.5: lands.pop(rnd.randrange(min(3,len(lands))))()
        if pending_mma and rnd.random()<0.5: pending_mma.pop(0)()
        idle+=1
        if idle>5_000_000: print("DEADLOCK?", n_tiles,nq,S,seed); return False
    while pending_mma: pending_mma.pop(0)()
    assert done[0]==16
    return True


@pytest.mark.parametrize("n_tiles,nq", [(1, 1), (1, 8), (3, 1), (3, 2), (5, 8), (4, 3)])
def test_multi_query_kernel_barrier_protocol(n_tiles, nq):
    for seed in range(4):
        assert run(n_tiles, nq, 6, seed)


# ---------------------------------------------------------------------------------------------------------------------
# The two fp16-pair formats stream more than one block per k-block through the ring and hold two ring stages at once:
#   f16x2: ring = Qh[kb], Kl[kb], Ql[kb];  MMAs (Qh,Kh[kb]) | (Qh,Kl) -> frees Qh, Kl | (Ql,Kh[kb]) -> frees Ql   (7 stages)
#   f16f8: ring = Qh[2j], Qh[2j+1], Qh8[j], Kl8[j], Ql8[j]; resident = 6 x Kh + 3 x Kh8                          (4 stages)
# Same model as above, driven by a per-(tile, query) schedule that mirrors the kernel's producer and issuer loops.
def _schedule(fmt):
    """-> (n_resident, ring labels in push order, MMA ops); an op = (ring labels it reads, resident index or None,
    ring labels it frees, resident index it frees after the tile's last query or None)"""
    if fmt == "bf16":
        ring = [("Qh", kb) for kb in range(6)]
        ops = [([("Qh", kb)], kb, [("Qh", kb)], kb) for kb in range(6)]
        return 6, ring, ops
    if fmt == "f16x2":
        ring, ops = [], []
        for kb in range(6):
            ring += [("Qh", kb), ("Kl", kb), ("Ql", kb)]
            ops += [([("Qh", kb)], kb, [], None), ([("Qh", kb), ("Kl", kb)], None, [("Qh", kb), ("Kl", kb)], None),
                    ([("Ql", kb)], kb, [("Ql", kb)], kb)]
        return 6, ring, ops
    ring, ops = [], []
    for j in range(3):
        ring += [("Qh", 2 * j), ("Qh", 2 * j + 1), ("Qh8", j), ("Kl8", j), ("Ql8", j)]
        for h in range(2):
            ops.append(([("Qh", 2 * j + h)], 2 * j + h, [("Qh", 2 * j + h)], 2 * j + h))
        ops.append(([("Qh8", j), ("Kl8", j)], None, [("Qh8", j), ("Kl8", j)], None))
        ops.append(([("Ql8", j)], 6 + j, [("Ql8", j)], 6 + j))
    return 9, ring, ops


def run_fmt(n_tiles, nq, fmt, S, seed):
    rnd = random.Random(seed)
    RES, ring, ops = _schedule(fmt)
    k_full = [Bar(2) for _ in range(RES)]
    q_full = [Bar(2) for _ in range(S)]
    k_empty = [[Bar(1) for _ in range(RES)] for _ in range(2)]
    q_empty = [[Bar(1) for _ in range(S)] for _ in range(2)]
    tmem_full = [[Bar(1) for _ in range(2)] for _ in range(2)]
    tmem_empty = [Bar(16) for _ in range(2)]
    kt = [[None] * RES for _ in range(2)]
    qs = [[None] * S for _ in range(2)]
    pending_mma, lands, done = [], [], [0]

    def kprod(c):
        phase = 0
        for t in range(n_tiles):
            for kb in range(RES):
                while not k_empty[c][kb].passed(phase ^ 1): yield
                def land(c=c, kb=kb, t=t): kt[c][kb] = ("K", t, kb); k_full[kb].arrive()
                lands.append(land)
                yield
            phase ^= 1

    def qprod(c):
        stage, phase = 0, 0
        for t in range(n_tiles):
            for b in range(nq):
                for lab in ring:
                    while not q_empty[c][stage].passed(phase ^ 1): yield
                    def land(c=c, stage=stage, t=t, b=b, lab=lab): qs[c][stage] = (lab, t, b); q_full[stage].arrive()
                    lands.append(land)
                    stage += 1
                    if stage == S: stage, phase = 0, phase ^ 1
                    yield

    def mma():
        stage, qphase, kphase, it = 0, 0, 0, 0
        for t in range(n_tiles):
            for b in range(nq):
                acc, ap = it & 1, (it >> 1) & 1
                while not tmem_empty[acc].passed(ap ^ 1): yield
                held = {}  # ring label -> stage, for blocks that stay in the ring across ops
                for reads, res, frees, res_free in ops:
                    if res is not None and b == 0:
                        while not k_full[res].passed(kphase): yield
                    for lab in reads:
                        if lab not in held:  # next block of the ring, in push order
                            while not q_full[stage].passed(qphase): yield
                            held[lab] = stage
                            stage += 1
                            if stage == S: stage, qphase = 0, qphase ^ 1
                    def comp(reads=tuple(reads), res=res, t=t, b=b, held=dict(held)):
                        for c in range(2):
                            if res is not None: assert kt[c][res] == ("K", t, res), (kt[c][res], t, b, res)
                            for lab in reads: assert qs[c][held[lab]] == (lab, t, b), (qs[c][held[lab]], lab, t, b)
                    pending_mma.append(comp)
                    for lab in frees:
                        def ce(st=held.pop(lab)):
                            for c in range(2): q_empty[c][st].arrive()
                        pending_mma.append(ce)
                    if res_free is not None and b == nq - 1:
                        def ke(kb=res_free):
                            for c in range(2): k_empty[c][kb].arrive()
                        pending_mma.append(ke)
                    yield
                assert not held
                def tf(acc=acc):
                    for c in range(2): tmem_full[c][acc].arrive()
                pending_mma.append(tf)
                it += 1
            kphase ^= 1

    def epi(c, w):
        it = 0
        for t in range(n_tiles):
            for b in range(nq):
                acc, ap = it & 1, (it >> 1) & 1
                while not tmem_full[c][acc].passed(ap): yield
                for _ in range(rnd.randint(0, 3)): yield
                tmem_empty[acc].arrive()
                it += 1
        done[0] += 1

    actors = [kprod(0), kprod(1), qprod(0), qprod(1), mma()] + [epi(c, w) for c in range(2) for w in range(8)]
    alive = set(range(len(actors)))
    steps = 0
    while alive:
        i = rnd.choice(sorted(alive))
        try: next(actors[i])
        except StopIteration: alive.discard(i)
        if lands and rnd.random() < 0.5: lands.pop(rnd.randrange(min(3, len(lands))))()
        if pending_mma and rnd.random() < 0.5: pending_mma.pop(0)()
        steps += 1
        if steps > 5_000_000: return False  # deadlock
    while pending_mma: pending_mma.pop(0)()
    return done[0] == 16


@pytest.mark.parametrize("fmt,stages", [("bf16", 6), ("f16x2", 7), ("f16f8", 4)])
@pytest.mark.parametrize("n_tiles,nq", [(1, 1), (2, 8), (3, 2), (4, 3)])
def test_fp16_pair_formats_barrier_protocol(fmt, stages, n_tiles, nq):
    """no deadlock and no early refill for the ring schedules of the three key formats at the ring depths the kernel uses
    (f16f8 has only 4 stages for 5 blocks per step and holds two of them across one MMA group)"""
    for seed in range(3):
        assert run_fmt(n_tiles, nq, fmt, stages, seed), (fmt, n_tiles, nq, seed)


def test_fp16_pair_ring_needs_two_stages_at_least():
    """the model does catch an impossible configuration: a group that holds two ring stages cannot run on a 1-stage ring"""
    assert not run_fmt(1, 1, "f16x2", 1, 0)
