"""CPU model of the ordered compaction's decoupled look-back (csrc/sift_tc.cu: scan_lookback, the
one-kernel form of the match path's tail).  getGoodMatches emits matches in ascending queryIdx
(featureMatchingCommon.cpp:43-49); on the device a unit (256 query rows of a pair) knows its own
kept count and needs the sum of the counts of the units before it, while all of them run
concurrently.  The model runs the kernel's protocol word for word -- {epoch:30, state:2, count:32}
words, 32-unit windows read nearest-first, wait only for the units up to the nearest published
prefix -- under random interleavings, with the array holding words of EARLIER launches (other
epochs, other counts, both states), and checks that every unit ends with the exclusive prefix sum
and that the protocol never waits on a unit behind it."""
import numpy as np
import pytest

AGG, PREFIX = 1, 2


def word(epoch, state, v):
    return (((epoch << 2) | state) << 32) | v


def unit_program(scan, epoch, unit, n, log):
    """Generator: one `yield` per memory round trip of the unit's warp (a window read)."""
    if unit == 0:
        scan[0] = word(epoch, PREFIX, n)
        return 0
    scan[unit] = word(epoch, AGG, n)
    base = 0
    j = unit - 1
    while True:
        yield                                    # the window is read at one instant (one load per lane)
        ws = []
        for lane in range(32):
            k = j - lane
            ws.append(word(epoch, PREFIX, 0) if k < 0 else int(scan[k]))
            if k >= 0:
                log.append((unit, k))
        hi = [w >> 32 for w in ws]
        ready = [(h >> 2) == epoch and (h & 3) != 0 for h in hi]
        is_prefix = [r and (h & 3) == PREFIX for r, h in zip(ready, hi)]
        first = is_prefix.index(True) if True in is_prefix else 32
        need = range(32) if first >= 32 else range(first + 1)
        if not all(ready[l] for l in need):
            continue                             # spin: read the window again later
        base += sum(ws[l] & 0xFFFFFFFF for l in range(32) if l <= first)
        if first < 32:
            break
        j -= 32
    scan[unit] = word(epoch, PREFIX, base + n)
    return base


@pytest.mark.parametrize("n_units,seed", [(1, 0), (2, 1), (33, 2), (40, 3), (97, 4), (320, 5)])
def test_every_interleaving_gives_the_exclusive_prefix(n_units, seed):
    rng = np.random.default_rng(seed)
    for trial in range(6):
        epoch = int(rng.integers(2, 1 << 30))
        counts = rng.integers(0, 257, n_units)
        if trial == 1:
            counts[:] = 0
        # what earlier launches left behind: any epoch but this one, any state, any count
        scan = np.zeros(n_units, dtype=object)
        for u in range(n_units):
            e = int(rng.integers(0, 1 << 30))
            e = e if e != epoch else e - 1
            scan[u] = word(e, int(rng.integers(0, 3)), int(rng.integers(0, 1 << 20)))
        log, result = [], {}
        # units START in index order (the hardware dispatches blocks in order) but run at random speeds
        started, running = 0, {}
        while len(result) < n_units:
            if started < n_units and (not running or rng.random() < 0.3):
                running[started] = unit_program(scan, epoch, started, int(counts[started]), log)
                started += 1
            u = list(running)[int(rng.integers(0, len(running)))]
            try:
                next(running[u])
            except StopIteration as stop:
                result[u] = stop.value
                del running[u]
        want = np.concatenate([[0], np.cumsum(counts)[:-1]])
        assert [result[u] for u in range(n_units)] == [int(x) for x in want]
        assert all(k < u for u, k in log)        # a unit only ever looks (and so waits) backwards
        assert int(scan[n_units - 1]) & 0xFFFFFFFF == int(counts.sum())


def test_a_slow_predecessor_only_delays_never_corrupts():
    """Unit 5 publishes nothing for a long time: the units behind it spin, the ones before finish."""
    epoch, counts = 7, [3, 1, 4, 1, 5, 9, 2, 6]
    scan = np.zeros(8, dtype=object)
    log, progs, result = [], {}, {}
    for u in range(8):
        if u != 5:
            progs[u] = unit_program(scan, epoch, u, counts[u], log)
    for _ in range(50):
        for u in list(progs):
            try:
                next(progs[u])
            except StopIteration as stop:
                result[u] = stop.value
                del progs[u]
    assert sorted(result) == [0, 1, 2, 3, 4] and sorted(progs) == [6, 7]
    progs[5] = unit_program(scan, epoch, 5, counts[5], log)
    while progs:
        for u in list(progs):
            try:
                next(progs[u])
            except StopIteration as stop:
                result[u] = stop.value
                del progs[u]
    assert [result[u] for u in range(8)] == [0, 3, 4, 8, 9, 14, 23, 25]
