"""Multi-GPU sharding of the keyframe window (BASELINE cfg4: 8 frames x 50 000 rows,
all i<j pairs), one process per GPU over torch.distributed.

Each rank holds the host descriptors of the frames it owns (frame f belongs to rank f mod G -- the
rank that extracted it).  A pair needs both of its frames, so this is the one place on the path
with a real exchange step: the ranks all-gather their frames' descriptor rows (NCCL over NVLink on
GPUs; 8 x 25.6 MB of fp32 rows for cfg4), every rank turns the gathered rows into resident
descriptor sets on its own device (slamb200_upload_desc_device: device-to-device prep, no host
round trip) and matches its share of the pairs.  The pair list is dealt round-robin (3-4 of the 28
pairs per rank at G = 8).  No reduction over partial distances is ever needed; the only other
collective is the gather of the per-pair match counts.

The collective plumbing is independent of the matcher: `upload_fn` and `match_fn` are injected
(product: Context.upload_device / Context.matchBatch; the CPU tests run the same plumbing over gloo
with host arrays).
"""
import numpy as np


def frame_owner(f, world):
    return f % world


def window_pairs(n_frames):
    """All i<j pairs in the order slamb200_match_window enumerates them."""
    return [(i, j) for i in range(n_frames) for j in range(i + 1, n_frames)]


def my_window_pairs(rank, world, n_frames):
    """Round-robin deal of the pair list: |share sizes| differ by at most one."""
    return [p for k, p in enumerate(window_pairs(n_frames)) if k % world == rank]


def exchange_frames(local_frames, n_frames, width, dtype, dist, device, group=None):
    """local_frames: {frame index: host array [rows, width]} for the frames this rank owns.
    Returns (rows_per_frame [n_frames], gathered tensor [world, slots, max_rows, width] on `device`)
    where frame f sits at [f % world, f // world, :rows[f]]."""
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    for f in local_frames:
        if frame_owner(f, world) != rank:
            raise ValueError(f"frame {f} is not owned by rank {rank}")
    rows = torch.zeros(n_frames, dtype=torch.int64, device=device)
    for f, a in local_frames.items():
        rows[f] = int(np.asarray(a).shape[0])
    dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group)      # every frame has one owner
    rows_host = [int(x) for x in rows.cpu().tolist()]
    max_rows = max(max(rows_host), 1)
    slots = (n_frames + world - 1) // world
    tdtype = {np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
    send = torch.zeros((slots, max_rows, width), dtype=tdtype, device=device)
    for f, a in local_frames.items():
        a = np.ascontiguousarray(a, dtype).reshape(-1, width)
        if a.shape[0]:
            send[f // world, : a.shape[0]] = torch.from_numpy(a).to(device, non_blocking=True)
    recv = torch.empty((world, slots, max_rows, width), dtype=tdtype, device=device)
    dist.all_gather_into_tensor(recv.view(world * slots, max_rows, width), send, group=group)
    return rows_host, recv


def match_window_sharded(local_frames, n_frames, width, dtype, dist, device, upload_fn, match_fn,
                         group=None):
    """Returns ({(i, j): matches} for this rank's pairs, counts [n_pairs] of ALL pairs, gathered).

    upload_fn(tensor_slice [rows, width] on `device`) -> a descriptor-set handle;
    match_fn(query_handle, [train handles]) -> list of match arrays (one per train)."""
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows, recv = exchange_frames(local_frames, n_frames, width, dtype, dist, device, group)
    mine = my_window_pairs(rank, world, n_frames)
    needed = sorted({f for p in mine for f in p})
    sets = {f: upload_fn(recv[frame_owner(f, world), f // world, : rows[f]]) for f in needed}
    out = {}
    for i in sorted({p[0] for p in mine}):
        js = [j for (a, j) in mine if a == i]
        for j, m in zip(js, match_fn(sets[i], [sets[j] for j in js])):
            out[(i, j)] = m
    pairs = window_pairs(n_frames)
    counts = torch.zeros(len(pairs), dtype=torch.int64, device=device)
    for k, p in enumerate(pairs):
        if p in out:
            counts[k] = len(out[p])
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return out, [int(x) for x in counts.cpu().tolist()], sets


def match_window_on_gpus(ctx, local_frames, n_frames, matcherType, knnMatcherDistance, dist, device,
                         group=None):
    """The product wiring of match_window_sharded: NCCL all-gather on torch's current stream, the
    gathered rows become resident descriptor sets through slamb200_upload_desc_device (ordered
    behind the all-gather by the stream handle), every rank matches its pairs with
    slamb200_match_batch.  Returns ({(i, j): matches}, counts of all pairs)."""
    import torch
    from . import _capi
    from .feature_matching import MatcherType
    orb = int(matcherType) == int(MatcherType.ORB_BF)
    width, dtype, kind = (32, np.uint8, _capi.DESC_U8X32) if orb else (128, np.float32, _capi.DESC_F32X128)
    stream = torch.cuda.current_stream(device).cuda_stream

    def upload(t):
        return ctx.upload_device(t.data_ptr(), int(t.shape[0]), kind, row_stride=width * t.element_size(),
                                 stream=stream)

    def match(q, ts):
        return ctx.matchBatch(q, ts, matcherType, knnMatcherDistance) if ts else []

    out, counts, sets = match_window_sharded(local_frames, n_frames, width, dtype, dist, device, upload,
                                             match, group)
    ctx.synchronize()      # the gathered rows may be released once every prep kernel has read them
    for s in sets.values():
        s.free()
    return out, counts


class PeerWindow:
    """The fused form of the exchange: no rows are gathered.  Every rank publishes the frames it
    owns once (slamb200_upload_desc_shared + a 128-byte IPC record, all-gathered), and matches its
    round-robin share of the pairs reading the other ranks' PREPARED operands over NVLink: a remote
    query set is consumed in place (the tcgen05 kernel TMA-loads its tiles from peer memory while it
    computes -- one pass over the operand per pair, +5 % kernel time measured), a remote train set
    is pulled once as a prepared slab (slamb200_desc_localize, 55 MB in 0.11 ms) because every
    query block re-reads it.  Mappings and local copies persist across match() calls, as they would
    while a window slides over the same frames."""

    def __init__(self, ctx, dist, device, group=None):
        self.ctx, self.dist, self.device, self.group = ctx, dist, device, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.own, self.records, self.peer, self.local_copy = {}, {}, {}, {}

    def publish(self, local_frames, n_frames):
        """Collective: upload the owned frames, exchange the IPC records of all n_frames."""
        import torch
        rec = torch.zeros((n_frames, 128), dtype=torch.uint8, device=self.device)
        for f, a in local_frames.items():
            if frame_owner(f, self.world) != self.rank:
                raise ValueError(f"frame {f} is not owned by rank {self.rank}")
            if f not in self.own:
                self.own[f] = self.ctx.upload_shared(np.ascontiguousarray(a))
            rec[f] = torch.frombuffer(bytearray(self.own[f].export_ipc()), dtype=torch.uint8).to(self.device)
        self.dist.all_reduce(rec, op=self.dist.ReduceOp.SUM, group=self.group)   # one owner per frame
        host = rec.cpu().numpy()
        for f in range(n_frames):
            self.records[f] = host[f].tobytes()

    def _query(self, f):
        if f in self.own:
            return self.own[f]
        if f not in self.peer:
            self.peer[f] = self.ctx.import_ipc(self.records[f])
        return self.peer[f]

    def _train(self, f):
        if f in self.own:
            return self.own[f]
        if f not in self.local_copy:
            self.local_copy[f] = self.ctx.localize(self._query(f))
        return self.local_copy[f]

    def match(self, n_frames, matcherType, knnMatcherDistance, gather_counts=True):
        """This rank's pairs {(i, j): matches} and (optionally, collective) all pairs' counts."""
        import torch
        mine = my_window_pairs(self.rank, self.world, n_frames)
        out = {}
        for i in sorted({p[0] for p in mine}):
            js = [j for (a, j) in mine if a == i]
            for j, m in zip(js, self.ctx.matchBatch(self._query(i), [self._train(j) for j in js],
                                                    matcherType, knnMatcherDistance)):
                out[(i, j)] = m
        if not gather_counts:
            return out, None
        pairs = window_pairs(n_frames)
        counts = torch.zeros(len(pairs), dtype=torch.int64, device=self.device)
        for k, p in enumerate(pairs):
            if p in out:
                counts[k] = len(out[p])
        self.dist.all_reduce(counts, op=self.dist.ReduceOp.SUM, group=self.group)
        return out, [int(x) for x in counts.cpu().tolist()]

    def close(self):
        """Collective: mappings are closed before any owner frees its set."""
        for ds in list(self.local_copy.values()) + list(self.peer.values()):
            ds.free()
        self.local_copy, self.peer = {}, {}
        self.dist.barrier(group=self.group)
        for ds in self.own.values():
            ds.free()
        self.own = {}


def match_window_peer(ctx, local_frames, n_frames, matcherType, knnMatcherDistance, dist, device,
                      group=None):
    """publish + match + close in one call: ({(i, j): matches}, counts of all pairs)."""
    w = PeerWindow(ctx, dist, device, group)
    w.publish(local_frames, n_frames)
    out, counts = w.match(n_frames, matcherType, knnMatcherDistance)
    w.close()
    return out, counts
