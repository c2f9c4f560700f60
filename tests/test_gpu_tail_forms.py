"""GPU: the tail of the tensor-core match path (slot merge, exact pruning, best-group rerank, ratio
test, ORDERED compaction -- getGoodMatches, featureMatchingCommon.cpp:37-50) in each of its forms:
inside the tcgen05 kernel (two tail warps per CTA, look-back compaction), as one kernel behind it,
as tail + compaction kernels, as separate kernels.  All must give the oracle's match list bit for
bit; the in-kernel form needs train sets of >= 12 column tiles and no empty shares, so the shapes
here straddle those limits (ragged sizes, partial last tiles, query sets that end inside a block)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, synth
from slam_indoor_code_b200.feature_matching import MatcherType

FORMS = (-1, 3, 2, 1, 0)   # -1: the default (by batch size)


@pytest.fixture()
def tail_ctx(ctx):
    yield ctx
    ctx.debug_tail_form(-1)


@pytest.mark.parametrize("nq,nt,seed", [(3000, 3072, 11), (2049, 3100, 12), (130, 5000, 13), (4097, 7001, 14),
                                        (1, 3073, 15), (257, 40000, 16)])
def test_sift_pair_every_form_equals_oracle(tail_ctx, nq, nt, seed):
    ctx = tail_ctx
    q, t = synth.sift_pair(nq, nt, seed)
    ref = c_oracle.match_features(0, q, t, 0.7)
    Q, T = ctx.upload(q), ctx.upload(t)
    for form in FORMS:
        ctx.debug_tail_form(form)
        for _ in range(2):   # twice: the look-back words and segment counters of the first call are reused
            assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, 0.7), ref), form
    Q.free(); T.free()


@pytest.mark.parametrize("nq,nt,seed", [(2500, 3072, 21), (1000, 9000, 22), (4100, 3333, 23)])
def test_orb_pair_every_form_equals_oracle(tail_ctx, nq, nt, seed):
    ctx = tail_ctx
    q, t = synth.orb_pair(nq, nt, seed)
    ref = c_oracle.match_features(int(MatcherType.ORB_BF), q, t, 0.8)
    Q, T = ctx.upload(q), ctx.upload(t)
    for form in FORMS:
        ctx.debug_tail_form(form)
        assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.ORB_BF, 0.8), ref), form
    Q.free(); T.free()


@pytest.mark.parametrize("ratio", [0.0, 0.5, 1.0, 1.5])
def test_ratios_and_ties_every_form(tail_ctx, ratio):
    """Masses of equal distances (lowest train index wins at every level) and ratios on both sides
    of 1 -- with ratio >= 1 nothing is pruned and every row's best group is evaluated."""
    ctx = tail_ctx
    rng = np.random.default_rng(31)
    base = rng.integers(0, 4, (40, 128)).astype(np.float32) * 20
    q = base[rng.integers(0, 40, 1500)]
    t = base[rng.integers(0, 40, 3500)]
    ref = c_oracle.match_features(0, q, t, ratio)
    Q, T = ctx.upload(q), ctx.upload(t)
    for form in FORMS:
        ctx.debug_tail_form(form)
        assert np.array_equal(ctx.matchFeatures(Q, T, MatcherType.SIFT_BF, ratio), ref), form
    Q.free(); T.free()


def test_ragged_batch_every_form(tail_ctx):
    """One query frame against train frames of different sizes, all large enough for the in-kernel
    tail, then the same batch with one small frame (which sends the whole batch to the kernel behind)."""
    ctx = tail_ctx
    q = synth.sift_like(2700, 41)
    sizes = [3072, 9000, 4001, 12345, 3329, 6000, 3072, 5555]
    trains = [synth.sift_train_from_query(q, n, 42 + i) for i, n in enumerate(sizes)]
    refs = [c_oracle.match_features(0, q, t, 0.7) for t in trains]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    small = synth.sift_train_from_query(q, 700, 77)
    Tsmall = ctx.upload(small)
    ref_small = c_oracle.match_features(0, q, small, 0.7)
    for form in FORMS:
        ctx.debug_tail_form(form)
        got = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
        assert all(np.array_equal(g, r) for g, r in zip(got, refs)), form
        got = ctx.matchBatch(Q, Ts[:3] + [Tsmall] + Ts[3:], MatcherType.SIFT_BF, 0.7)
        assert all(np.array_equal(g, r) for g, r in zip(got, refs[:3] + [ref_small] + refs[3:])), form


def test_window_of_equal_frames_forms_agree(tail_ctx):
    """48 pairs of 10 000 x 10 000 rows (the bench's shape, shorter): the forms agree pair by pair,
    and the first and last pairs equal the oracle."""
    ctx = tail_ctx
    q = synth.sift_like(10000, 51)
    trains = [synth.sift_train_from_query(q, 10000, 52 + i) for i in range(48)]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    ctx.debug_tail_form(3)
    got3 = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    again = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    assert all(np.array_equal(a, b) for a, b in zip(got3, again))
    for form in (2, 1):
        ctx.debug_tail_form(form)
        got = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
        assert all(np.array_equal(a, b) for a, b in zip(got3, got)), form
    for k in (0, 47):
        assert np.array_equal(got3[k], c_oracle.match_features(0, q, trains[k], 0.7))


def test_lookback_compaction_is_stable_under_load(tail_ctx):
    """The one-kernel tail orders its output by a look-back over blocks that run concurrently; 40
    launches of a 40-pair batch (1600 blocks in flight, more than fit on the GPU at once) must give
    the same lists every time, equal to the two-kernel form's."""
    ctx = tail_ctx
    q = synth.sift_like(5000, 61)
    trains = [synth.sift_train_from_query(q, 3000 + 257 * (i % 7), 62 + i) for i in range(40)]
    Q = ctx.upload(q)
    Ts = [ctx.upload(t) for t in trains]
    ctx.debug_tail_form(1)
    want = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    ctx.debug_tail_form(2)
    for it in range(40):
        got = ctx.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), it
