"""A CPU model of the candidate logic behind the tensor-core matchers (csrc/sift_tc.cu), checked
against brute force: per 32-column chunk the four 8-column group minima, the running top-2 over
chunks with strict '<' (top2_chunk_insert), the merge of slot records (sift_merge_kernel /
tc_tail_fused_kernel), and the two ways the candidates are consumed --

  * match output: only the best group's 8 columns are evaluated; the second distance is
    min(second of that group, L), L = min(second group minimum of the best chunk, second chunk);
  * raw k-NN output: the best chunk's 32 columns + the 8 columns of the second chunk's group.

The model works on exact integer distances (what the tcgen05 accumulators hold for cv::SIFT rows and
for ORB bits), with masses of ties, ragged column counts and arbitrary slot boundaries: the claims
"ties go to the lowest train index" and "40 candidates are enough" are properties of this logic,
not of the hardware.  The kernels themselves are tested on the GPU against the oracle."""
import numpy as np

INF = float("inf")
ABSENT = 0xFFFF


def slot_record(row, c0, c1):
    """Epilogue of one thread over columns [c0, c1) (multiples of 32): returns (m1, i1, s1, m2, i2)."""
    m1 = m2 = s1 = INF
    i1 = i2 = -1
    for c in range(c0, c1, 32):
        g = [row[c + 8 * j: c + 8 * j + 8].min() for j in range(4)]
        gid0 = c // 8
        m01, m23 = min(g[0], g[1]), min(g[2], g[3])
        j01 = gid0 + 1 if g[1] < g[0] else gid0
        j23 = gid0 + 3 if g[3] < g[2] else gid0 + 2
        cm = min(m01, m23)
        gid = j23 if m23 < m01 else j01
        c2 = min(max(m01, m23), max(g[0], g[1]), max(g[2], g[3]))     # second smallest of the four
        lt1, lt2 = cm < m1, cm < m2                                   # top2_chunk_insert
        m2, i2 = (m1, i1) if lt1 else ((cm, gid) if lt2 else (m2, i2))
        if lt1:
            m1, i1, s1 = cm, gid, c2
    return m1, (i1 & 0xFFFF), s1, m2, (i2 & 0xFFFF)


def lt_fi(va, ia, vb, ib):
    return va < vb or (va == vb and ia < ib)


def merge(records):
    v0 = v1 = s0 = INF
    g0 = g1 = ABSENT
    for a, ia, sa, b, ib in records:
        if ia != ABSENT:
            if lt_fi(a, ia, v0, g0):
                v1, g1, v0, g0, s0 = v0, g0, a, ia, sa
            elif lt_fi(a, ia, v1, g1):
                v1, g1 = a, ia
        if ib != ABSENT and lt_fi(b, ib, v1, g1):
            v1, g1 = b, ib
    return v0, g0, s0, v1, g1


def brute_top2(row, t_n):
    keys = sorted((row[c], c) for c in range(t_n))
    return keys[:2]


def run_case(rng, q_rows, t_n, vmax):
    t_pad = (t_n + 255) // 256 * 256
    D = rng.integers(0, vmax, (q_rows, t_pad)).astype(np.float64)
    D[:, t_n:] = INF                                              # padding columns
    # slots: contiguous column ranges, boundaries on multiples of 128 (half tiles), random cuts
    cuts = sorted(set([0, t_pad] + [int(c) * 128 for c in rng.integers(0, t_pad // 128 + 1, 3)]))
    for r in range(q_rows):
        row = D[r]
        recs = [slot_record(row, a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        v0, g0, s0, v1, g1 = merge(recs)
        top = brute_top2(row, t_n)
        if t_n == 0:
            assert g0 == ABSENT
            continue
        assert g0 != ABSENT and v0 == top[0][0]
        # match output: the best group only
        L = min(s0, v1)
        cols = [c for c in range(g0 * 8, g0 * 8 + 8) if c < t_n]
        grp = sorted((row[c], c) for c in cols)
        assert grp[0] == top[0]                                    # best index, lowest on ties
        d1 = min(grp[1][0] if len(grp) > 1 else INF, L)
        assert d1 == (top[1][0] if len(top) > 1 else INF)          # the second distance, exactly
        # raw k-NN output: 32 + 8 candidates
        cand = [c for c in range((g0 >> 2) * 32, (g0 >> 2) * 32 + 32) if c < t_n]
        if g1 != ABSENT:
            cand += [c for c in range(g1 * 8, g1 * 8 + 8) if c < t_n]
        keys = sorted((row[c], c) for c in cand)
        assert keys[:2] == top


def test_candidate_logic_equals_brute_force():
    rng = np.random.default_rng(404)
    for t_n in (0, 1, 2, 7, 8, 9, 31, 33, 255, 256, 257, 700, 1023):
        for vmax in (2, 5, 300):                                   # masses of ties ... few ties
            run_case(rng, 24, t_n, vmax)


# ---- general floats: approximate candidates + certificate (sift_tc_kernel<GEN>, sift_gen_rerank_kernel) ----
def gen_slot_record(approx, c0, c1):
    """GEN epilogue: best four chunks (value, group of the minimum), fifth chunk minimum s, smallest
    second group minimum s2 of any chunk."""
    m = [INF] * 4
    i = [-1] * 4
    s = s2 = INF
    for c in range(c0, c1, 32):
        g = [approx[c + 8 * j: c + 8 * j + 8].min() for j in range(4)]
        gid0 = c // 8
        m01, m23 = min(g[0], g[1]), min(g[2], g[3])
        j01 = gid0 + 1 if g[1] < g[0] else gid0
        j23 = gid0 + 3 if g[3] < g[2] else gid0 + 2
        cm = min(m01, m23)
        gid = j23 if m23 < m01 else j01
        c2 = min(max(m01, m23), max(g[0], g[1]), max(g[2], g[3]))
        l = [cm < m[k] for k in range(4)]
        s = m[3] if l[3] else min(s, cm)
        s2 = min(s2, c2)
        m[3], i[3] = (m[2], i[2]) if l[2] else ((cm, gid) if l[3] else (m[3], i[3]))
        m[2], i[2] = (m[1], i[1]) if l[1] else ((cm, gid) if l[2] else (m[2], i[2]))
        m[1], i[1] = (m[0], i[0]) if l[0] else ((cm, gid) if l[1] else (m[1], i[1]))
        if l[0]:
            m[0], i[0] = cm, gid
    return m, i, s, s2


def gen_rerank(exact, approx, t_n, cuts, E):
    """Returns (certified, top2) following sift_gen_rerank_kernel's two stages."""
    entries, rest, rest2 = [], INF, INF
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        m, i, s, s2 = gen_slot_record(approx, a, b)
        rest, rest2 = min(rest, s), min(rest2, s2)
        entries += [(m[k], i[k]) for k in range(4) if i[k] >= 0]
    entries.sort()
    chosen, others = entries[:4], entries[4:]
    rest = min([rest] + [v for v, _ in others])

    def certified_against(bound, keys):
        if not bound < INF:
            return True
        if len(keys) < 2:
            return False
        return 2.0 * bound - E > keys[1][0]      # every other column is provably farther than the 2nd

    cols = [c for _, g in chosen for c in range(g * 8, g * 8 + 8) if c < t_n]
    keys = sorted((exact[c], c) for c in cols)
    if certified_against(min(rest, rest2), keys):
        return True, keys[:2]
    cols = [c for _, g in chosen for c in range((g >> 2) * 32, (g >> 2) * 32 + 32) if c < t_n]
    keys = sorted((exact[c], c) for c in cols)
    return certified_against(rest, keys), keys[:2]


def test_certificate_is_sound_under_bounded_error():
    """Whatever the approximation errors (within the bound E on d^2), a certified answer is the
    brute-force one, ties included; uncertified rows are the only ones left to the fallback."""
    rng = np.random.default_rng(405)
    n_cert = n_total = 0
    for t_n in (1, 2, 9, 40, 257, 900):
        t_pad = (t_n + 255) // 256 * 256
        for vmax, E in ((6, 3.0), (50, 4.0), (1000, 6.0), (100000, 8.0)):
            for _ in range(40):
                exact = rng.integers(0, vmax, t_pad).astype(np.float64)     # d^2, many ties when vmax is small
                approx = exact / 2 + rng.uniform(-E / 2, E / 2, t_pad)      # accumulator ~ d^2 / 2, |2 a - d^2| <= E
                exact[t_n:] = INF
                approx[t_n:] = INF
                cuts = sorted(set([0, t_pad] + [int(c) * 128 for c in rng.integers(0, t_pad // 128 + 1, 2)]))
                ok, top = gen_rerank(exact, approx, t_n, cuts, E)
                n_total += 1
                if ok:
                    n_cert += 1
                    assert top == brute_top2(exact, t_n)
    assert n_cert > n_total // 4       # the certificate does fire (spread-out distances certify)
