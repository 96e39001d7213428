"""The digit selection of the radix top-k exists twice in 6dgs_b200/csrc/topk.cu: a serial walk over the 256 bins
(topk_select_kernel, validated on the GPU) and a parallel suffix-sum rule evaluated by 256 threads
(topk_hist_select_kernel, experimental).  This model checks on the CPU that the two rules always pick the same
digit and the same remaining count."""
import random


def serial_rule(hist, k_rem):
    above, d = 0, 255
    while d > 0:
        c = hist[d]
        if above + c >= k_rem:
            break
        above += c
        d -= 1
    return d, k_rem - above


def parallel_rule(hist, k_rem):
    hits = []
    for t in range(256):
        s = sum(hist[t:])          # elements with digit >= t
        above = s - hist[t]        # elements with digit >  t
        if (s >= k_rem and above < k_rem) or (t == 0 and s < k_rem):
            hits.append((t, k_rem - above))
    assert len(hits) == 1, hits    # exactly one thread writes the state
    return hits[0]


def test_parallel_digit_selection_equals_serial_walk():
    rnd = random.Random(0)
    for trial in range(3000):
        style = trial % 4
        if style == 0:
            hist = [rnd.randint(0, 50) for _ in range(256)]
        elif style == 1:   # sparse
            hist = [rnd.randint(1, 2000) if rnd.random() < 0.03 else 0 for _ in range(256)]
        elif style == 2:   # everything in one bin
            hist = [0] * 256
            hist[rnd.randrange(256)] = rnd.randint(1, 10_000)
        else:              # fewer elements than requested (cannot happen with n >= k; the rules must still agree)
            hist = [rnd.randint(0, 2) for _ in range(256)]
        total = sum(hist)
        for k_rem in {1, 2, max(1, total // 2), max(1, total), total + 5, rnd.randint(1, max(1, total))}:
            assert serial_rule(hist, k_rem) == parallel_rule(hist, k_rem), (hist, k_rem)
