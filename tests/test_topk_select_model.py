"""CPU model of the digit-selection rule of the radix top-k (6dgs_b200/csrc/topk.cu: topk_select_kernel).  The kernel
picks, from a 256-bin histogram and the remaining rank k_rem, the digit d of the k-th key with 256 threads and suffix sums:
the unique thread t with (t == 0 or suf[t] >= k_rem) and suf[t + 1] < k_rem, then k_rem -= suf[t + 1].  Round 1 did the same
with a serial walk from the top bin; the two must agree for every histogram, including empty bins, k_rem = 1, k_rem = total
and the out-of-contract k_rem > total (both fall back to bin 0)."""
import random

import pytest


def serial(hist, k_rem):
    above, d = 0, 255
    while d > 0:
        c = hist[d]
        if above + c >= k_rem:
            break
        above += c
        d -= 1
    return d, k_rem - above


def parallel(hist, k_rem):
    suf = [0] * 257
    for t in range(255, -1, -1):
        suf[t] = suf[t + 1] + hist[t]
    hits = [t for t in range(256) if (t == 0 or suf[t] >= k_rem) and suf[t + 1] < k_rem]
    assert len(hits) == 1, hits  # exactly one thread writes the state
    return hits[0], k_rem - suf[hits[0] + 1]


@pytest.mark.parametrize("seed", range(6))
def test_suffix_scan_selection_equals_serial_walk(seed):
    rnd = random.Random(seed)
    for _ in range(300):
        kind = rnd.randrange(4)
        if kind == 0:
            hist = [rnd.randrange(0, 5) for _ in range(256)]
        elif kind == 1:
            hist = [0] * 256
            for _ in range(rnd.randrange(1, 6)):
                hist[rnd.randrange(256)] = rnd.randrange(1, 10_000_000)
        elif kind == 2:
            hist = [rnd.randrange(0, 2) * rnd.randrange(0, 1000) for _ in range(256)]
        else:
            hist = [0] * 256
            hist[rnd.randrange(256)] = rnd.randrange(1, 1 << 30)  # everything in one bin (heavy ties)
        total = sum(hist)
        for k_rem in {1, max(1, total // 2), max(1, total), total + 7, rnd.randrange(1, max(2, total + 1))}:
            assert parallel(hist, k_rem) == serial(hist, k_rem), (hist, k_rem)
