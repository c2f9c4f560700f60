"""One process, every GPU of the box (include/slamb200.h "device set").

The reference is a single process whose batch search strides the framesBatchSize window with host
threads (src/mainModule/cycleProcessing/batch.cpp:162-226).  A DeviceSet gives that process all its
GPUs behind the same call shape: the query frame is replicated (peer copies of the prepared set
over NVLink), each train frame lives on the device that will match it, one call matches the whole
window.  Results equal Context.matchBatch on one device, pair for pair.
"""
import ctypes

import numpy as np

from . import _capi
from ._capi import DMATCH, check, ptr


class SetDescriptor:
    def __init__(self, dset, h, n, member):
        self._set, self._h, self.n, self.member = dset, h, n, member

    def free(self):
        if self._h is not None and self._set._h is not None:
            self._set._lib.slamb200_set_free_desc(self._set._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceSet:
    """slamb200_set: contexts on devices 0..n-1 of this process (n <= 0: every visible device)."""

    def __init__(self, n_devices=0):
        self._lib = _capi.load()
        h = ctypes.c_void_p()
        check(self._lib.slamb200_set_init(int(n_devices), ctypes.byref(h)))
        self._h = h
        self.n = int(self._lib.slamb200_set_devices(h))
        self._last = (0, 1)

    def close(self):
        if self._h is not None:
            self._lib.slamb200_set_shutdown(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def owner(self, i, n):
        """Member that owns item i of n under the contiguous split."""
        return int(self._lib.slamb200_set_owner(self._h, int(i), int(n)))

    def set_pack_threads(self, n):
        for i in range(self.n):
            check(self._lib.slamb200_set_pack_threads(self._lib.slamb200_set_ctx(self._h, i), int(n)))

    def upload(self, desc, member=-1):
        """Host descriptor Mat -> resident on `member`, or replicated on every member (member < 0)."""
        desc = np.asarray(desc)
        kind = _capi.DESC_U8X32 if desc.dtype == np.uint8 else _capi.DESC_F32X128
        width = 128 if kind == _capi.DESC_F32X128 else 32
        if desc.size == 0:
            desc = np.zeros((0, width), desc.dtype)
        if desc.ndim != 2 or desc.shape[1] != width:
            raise ValueError(f"descriptor shape {desc.shape}, expected (N, {width})")
        if desc.strides[1] != desc.itemsize:
            desc = np.ascontiguousarray(desc)
        stride = desc.strides[0] if desc.shape[0] > 1 else width * desc.itemsize
        h = ctypes.c_void_p()
        check(self._lib.slamb200_set_upload(self._h, int(member), kind, ptr(desc), desc.shape[0], stride,
                                            ctypes.byref(h)))
        return SetDescriptor(self, h, desc.shape[0], member)

    def matchBatch(self, prevDesc, trainDescs, matcherType, knnMatcherDistance=0.7, out=None, n_out=None):
        """The batch window of batch.cpp:120-148 over every device: list of good-match arrays."""
        P = len(trainDescs)
        cap = max(prevDesc.n, 1)
        own = out is None
        if own:
            out = np.empty((max(P, 1), cap), DMATCH)
            n_out = np.zeros(max(P, 1), np.int32)
        arr = (ctypes.c_void_p * max(P, 1))(*[t._h for t in trainDescs])
        check(self._lib.slamb200_set_match_batch(self._h, int(matcherType), prevDesc._h, arr, P,
                                                 float(knnMatcherDistance), ptr(out), out.shape[1], ptr(n_out)))
        return [out[p, : n_out[p]].copy() if own else out[p, : n_out[p]] for p in range(P)]

    def matchBatchEnqueue(self, prevDesc, trainDescs, matcherType, knnMatcherDistance=0.7):
        P = len(trainDescs)
        arr = (ctypes.c_void_p * max(P, 1))(*[t._h for t in trainDescs])
        check(self._lib.slamb200_set_match_batch_enqueue(self._h, int(matcherType), prevDesc._h, arr, P,
                                                         float(knnMatcherDistance)))
        self._last = (P, max(prevDesc.n, 1))

    def batchFetch(self, out=None, n_out=None):
        """Match lists of the last enqueue and each member's device time (ms) for its share."""
        P, cap = self._last
        if out is None:
            out = np.empty((max(P, 1), cap), DMATCH)
            n_out = np.zeros(max(P, 1), np.int32)
        ms = np.zeros(self.n, np.float32)
        check(self._lib.slamb200_set_batch_fetch(self._h, ptr(out), out.shape[1], ptr(n_out), ptr(ms)))
        return [out[p, : n_out[p]] for p in range(P)], n_out[:P], ms
