"""GPU: the in-process device set (include/slamb200.h) -- one process, all GPUs of the
box behind the reference's batch call shape (batch.cpp:162-226) -- returns what the single-device
slamb200_match_batch returns, pair for pair, whatever the number of members and wherever the train
frames live.  With one visible GPU the one-member set is still exercised; the multi-member cases
are SKIPPED there (visibly), not passed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import synth_inputs as synth
from oracle import c_oracle
from slam_indoor_code_b200.device_set import DeviceSet
from slam_indoor_code_b200.feature_matching import MatcherType


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _window(n_trains, rows=2500, seed=9100):
    q = synth.sift_like(rows, seed)
    sizes = [rows, rows - 300, 1, 0, rows + 700] + [rows] * max(0, n_trains - 5)
    trains = [synth.sift_train_from_query(q, max(s, 1), seed + 1 + i)[:s] for i, s in enumerate(sizes[:n_trains])]
    return q, trains


def _check(dset, q, trains, want, placement):
    Q = dset.upload(q)                                   # replicated over the members
    Ts = [dset.upload(t, placement(i)) for i, t in enumerate(trains)]
    got = dset.matchBatch(Q, Ts, MatcherType.SIFT_BF, 0.7)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    dset.matchBatchEnqueue(Q, Ts, MatcherType.SIFT_BF, 0.7)
    got2, n_out, ms = dset.batchFetch()
    for g, w in zip(got2, want):
        assert np.array_equal(g, w)
    for h in Ts + [Q]:
        h.free()
    return ms


def test_one_member_set_equals_context(ctx):
    q, trains = _window(9)
    want = ctx.matchBatch(ctx.upload(q), [ctx.upload(t) for t in trains], MatcherType.SIFT_BF, 0.7)
    assert np.array_equal(want[0], c_oracle.match_features(0, q, trains[0], 0.7))
    with DeviceSet(1) as ds:
        assert ds.n == 1
        _check(ds, q, trains, want, lambda i: 0)
        _check(ds, q, trains, want, lambda i: -1)


@pytest.mark.parametrize("members", [2, 4, 8])
def test_multi_member_set_equals_single_device(ctx, members):
    if _n_gpus() < members:
        pytest.skip(f"{members}-member device set needs {members} GPUs, {_n_gpus()} visible "
                    "(run under gpurun --gpus N; bench.py --gpus N checks it at every N as well)")
    q, trains = _window(23)
    want = ctx.matchBatch(ctx.upload(q), [ctx.upload(t) for t in trains], MatcherType.SIFT_BF, 0.7)
    with DeviceSet(members) as ds:
        assert ds.n == members
        n = len(trains)
        ms = _check(ds, q, trains, want, lambda i: ds.owner(i, n))          # contiguous split
        assert np.all(ms > 0)                                               # every member worked
        _check(ds, q, trains, want, lambda i: i % members)                  # round robin: scattered results
        _check(ds, q, trains, want, lambda i: -1)                           # replicated trains: balanced
        _check(ds, q, trains, want, lambda i: members - 1)                  # everything on the last member
    # ORB through the same path
    qo, _ = synth.orb_pair(1800, 10, 9200)
    to = [synth.orb_pair(10, 1500 + 100 * i, 9201 + i)[1] for i in range(7)]
    want = ctx.matchBatch(ctx.upload(qo), [ctx.upload(t) for t in to], MatcherType.ORB_BF, 0.7)
    with DeviceSet(members) as ds:
        Q = ds.upload(qo)
        Ts = [ds.upload(t, ds.owner(i, len(to))) for i, t in enumerate(to)]
        for g, w in zip(ds.matchBatch(Q, Ts, MatcherType.ORB_BF, 0.7), want):
            assert np.array_equal(g, w)
