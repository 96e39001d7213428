"""Barrier-protocol model of the experimental multi-query score kernel (6dgs_b200/csrc/score_tc_mq.cu), runnable
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
