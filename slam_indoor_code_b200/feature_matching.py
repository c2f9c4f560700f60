"""Host-side mirror of the reference's matching interface over the libslamb200 C ABI.

Names, argument meaning and error behaviour follow
  src/mainModule/featureMatching/featureMatching.h:12-53        (matchFramesPairFeatures)
  src/mainModule/featureMatching/featureMatchingCPU.cpp:17-43   (matchFeatures)
  src/mainModule/featureMatching/featureMatchingCommon.{h,cpp}  (MatcherType, getGoodMatches,
                                                                 getKeyPointCoordsFromFramePair)
so that the parity tests read like the reference's call sites (batch.cpp:113-130).  Descriptors are
NumPy arrays in cv::Mat layout (N x 128 float32 for SIFT, N x 32 uint8 for ORB); matches come
back as a structured array bit-compatible with cv::DMatch.  The compiled-host twin of this file
is host/featureMatchingB200.cpp.
"""
import ctypes
import enum

import numpy as np

from . import _capi
from ._capi import DMATCH, check, ptr


class MatcherType(enum.IntEnum):
    """featureMatchingCommon.h:8-12"""
    SIFT_BF = 0
    SIFT_FLANN = 1
    ORB_BF = 2
    # not in the reference enum: what useFM-SIFT-BF means in its OpenCV-CUDA build (NORM_L1,
    # featureMatchingCUDA.cpp:28); SURVEY.md 8f-4
    SIFT_BF_L1 = 3


def getMatcherTypeIndex(config):
    """featureMatchingCommon.cpp:13-21: flag priority useFM-SIFT-BF > useFM-SIFT-FLANN > useFM-ORB;
    none set -> the reference throws."""
    if config.get("useFM-SIFT-BF"):
        return MatcherType.SIFT_BF
    if config.get("useFM-SIFT-FLANN"):
        return MatcherType.SIFT_FLANN
    if config.get("useFM-ORB"):
        return MatcherType.ORB_BF
    raise RuntimeError("no matcher flag set (useFM-SIFT-BF / useFM-SIFT-FLANN / useFM-ORB)")


def _kind_of(matcher_type):
    if matcher_type in (MatcherType.SIFT_BF, MatcherType.SIFT_FLANN, MatcherType.SIFT_BF_L1):
        return _capi.DESC_F32X128
    if matcher_type == MatcherType.ORB_BF:
        return _capi.DESC_U8X32
    # featureMatchingCPU.cpp:36-37 / :62-63: default -> throw std::exception()
    raise RuntimeError(f"bad matcher type {matcher_type!r}")


class Context:
    """One libslamb200 context bound to one CUDA device."""

    def __init__(self, device=0):
        self._lib = _capi.load()
        h = ctypes.c_void_p()
        check(self._lib.slamb200_init(int(device), ctypes.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.slamb200_shutdown(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._lib.slamb200_synchronize(self._h))

    def launch_count(self):
        return int(self._lib.slamb200_launch_count(self._h))

    KERNELS = ("sift_tc", "sift_exact", "orb", "ransac", "sift_rerank", "finalize", "sift_tc_gen",
               "sift_gen_rerank", "pnp", "sift_l1", "orb_desc", "fast", "sift_desc")

    def debug_orb_kernel(self, tensor_cores=True):
        """Developer switch: ORB pairs through the tcgen05 kernel on e4m3 0/1 bytes (default) or
        through the XOR/POPC kernel.  Both give cv::BFMatcher(NORM_HAMMING)'s results."""
        f = self._lib.slamb200_dbg_set_tc_orb
        f.argtypes = [ctypes.c_void_p, ctypes.c_int]
        f.restype = ctypes.c_int
        check(f(self._h, 1 if tensor_cores else 0))

    def debug_tail_form(self, form=-1):
        """Developer switch for the tail of the tensor-core match path (slot merge, best-group rerank,
        ratio test, ordered compaction): 2 = one kernel behind the tcgen05 kernel, 1 = tail kernel +
        compaction kernel, 0 = separate merge / rerank / finalize kernels, 3 = experimental, inside the
        tcgen05 kernel where the batch allows it; -1 (default) = 2 for batches of up to 32 pairs, 1
        beyond.  Every form gives the same match lists."""
        f = self._lib.slamb200_dbg_set_fused_tail
        f.argtypes = [ctypes.c_void_p, ctypes.c_int]
        f.restype = ctypes.c_int
        check(f(self._h, int(form)))

    def profile_enable(self, on=True, kinds=None):
        """Per-kernel-class device times (profile_read).  kinds: names from KERNELS to restrict the
        events to (a timed region that only wants its dominant kernel's duration)."""
        if on and kinds is not None:
            mask = 0
            for k in kinds:
                mask |= 1 << self.KERNELS.index(k)
            check(self._lib.slamb200_profile_enable_kinds(self._h, mask))
        else:
            check(self._lib.slamb200_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kernel: (summed device ms, launches)} since the last read (synchronises)."""
        ms = np.zeros(len(self.KERNELS), np.float64)
        n = np.zeros(len(self.KERNELS), np.int64)
        check(self._lib.slamb200_profile_read(self._h, ptr(ms), ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNELS)}

    # -- descriptor sets ("upload once per frame") --------------------------------------------
    def upload(self, desc, kind=None, _pinned=False):
        """Host descriptor Mat (N x 128 float32 or N x 32 uint8, row pitch honoured) -> HBM."""
        desc = np.asarray(desc)
        if kind is None:
            kind = _capi.DESC_U8X32 if desc.dtype == np.uint8 else _capi.DESC_F32X128
        width = 128 if kind == _capi.DESC_F32X128 else 32
        want = np.float32 if kind == _capi.DESC_F32X128 else np.uint8
        if desc.dtype != want:
            raise TypeError(f"descriptor dtype {desc.dtype} does not fit kind {kind}")
        if desc.size == 0:
            desc = np.zeros((0, width), want)
        if desc.ndim != 2 or desc.shape[1] != width:
            raise ValueError(f"descriptor shape {desc.shape}, expected (N, {width})")
        if desc.strides[1] != desc.itemsize:
            desc = np.ascontiguousarray(desc)
        stride = desc.strides[0] if desc.shape[0] > 1 else width * desc.itemsize
        h = ctypes.c_void_p()
        fn = {False: self._lib.slamb200_upload_desc, True: self._lib.slamb200_upload_desc_pinned,
              "packed": self._lib.slamb200_upload_desc_packed,
              "shared": self._lib.slamb200_upload_desc_shared}[_pinned]
        check(fn(self._h, kind, ptr(desc), desc.shape[0], stride, ctypes.byref(h)))
        return DescriptorSet(self, h, desc.shape[0], kind)

    def upload_pinned(self, desc, kind=None):
        """Pipelined upload from page-locked host memory (no synchronisation; the caller keeps
        `desc` alive and unmodified until synchronize() or until results were fetched)."""
        return self.upload(desc, kind, _pinned=True)

    def set_pack_threads(self, n):
        """Host threads sharing the narrowing inside upload_packed (before its first use)."""
        check(self._lib.slamb200_set_pack_threads(self._h, int(n)))

    def upload_packed(self, desc, kind=None):
        """Upload that narrows integer-valued SIFT rows to bytes on the calling thread (every
        element verified) so that a quarter of the bytes cross PCIe; `desc` is consumed on return.
        Anything else takes upload()'s path.  Same results as upload()."""
        return self.upload(desc, kind, _pinned="packed")

    # -- sets shared between the ranks of a node (one process per GPU; CUDA IPC over NVLink) --------
    def upload_shared(self, desc, kind=None):
        """Like upload(), into an allocation other processes can map (export_ipc / import_ipc)."""
        return self.upload(desc, kind, _pinned="shared")

    def import_ipc(self, record):
        """A peer rank's exported set (the 128 bytes of DescriptorSet.export_ipc()) as a handle
        on this rank: its prepared operands are read over NVLink by whatever kernel uses it."""
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(bytes(record))
        h = ctypes.c_void_p()
        check(self._lib.slamb200_desc_import(self._h, ctypes.byref(buf), ctypes.byref(h)))
        n = int(self._lib.slamb200_desc_rows(h))
        kind = int(self._lib.slamb200_desc_kind(h))
        return DescriptorSet(self, h, n, kind)

    def localize(self, ds):
        """A local copy of a (peer-mapped) set: one device-to-device transfer of the prepared slab."""
        h = ctypes.c_void_p()
        check(self._lib.slamb200_desc_localize(self._h, ds._h, ctypes.byref(h)))
        return DescriptorSet(self, h, ds.n, ds.kind)

    def upload_device(self, dev_ptr, n, kind, row_stride=0, stream=None):
        """Rows already resident on this device (raw pointer, e.g. torch.Tensor.data_ptr())."""
        h = ctypes.c_void_p()
        check(self._lib.slamb200_upload_desc_device(self._h, kind, ctypes.c_void_p(dev_ptr), n,
                                                    row_stride, ctypes.c_void_p(stream or 0),
                                                    ctypes.byref(h)))
        return DescriptorSet(self, h, n, kind)

    def upload_keypoints(self, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        h = ctypes.c_void_p()
        check(self._lib.slamb200_upload_pts(self._h, ptr(xy), xy.shape[0], 8, ctypes.byref(h)))
        return KeypointSet(self, h, xy.shape[0])

    # -- hot path A ------------------------------------------------------------------------------
    def knnMatch(self, matcher_type, query, train):
        """DescriptorMatcher::knnMatch(query, train, k=2) raw: idx[nq,2] (-1 absent), dist[nq,2]."""
        _kind_of(matcher_type)
        idx = np.full((query.n, 2), -1, np.int32)
        dist = np.zeros((query.n, 2), np.float32)
        check(self._lib.slamb200_knn2(self._h, int(matcher_type), query._h, train._h, ptr(idx),
                                      ptr(dist)))
        return idx, dist

    def matchFeatures(self, prevDesc, curDesc, matcherType, knnMatcherDistance=0.7):
        """featureMatchingCPU.cpp:17-43 on resident descriptor sets: knnMatch(prev, cur, 2) then
        getGoodMatches.  Returns the good matches (DMATCH structured array, ascending queryIdx)."""
        _kind_of(matcherType)
        cap = max(prevDesc.n, 1)
        out = np.zeros(cap, DMATCH)
        n = ctypes.c_int(0)
        check(self._lib.slamb200_match_pair(self._h, int(matcherType), prevDesc._h, curDesc._h,
                                            float(knnMatcherDistance), ptr(out), cap,
                                            ctypes.byref(n)))
        return out[: n.value].copy()

    def matchBatch(self, prevDesc, trainDescs, matcherType, knnMatcherDistance=0.7, out=None,
                   n_out=None):
        """The batch window of batch.cpp:120-148 in one call: list of good-match arrays.  `out`
        ([P, rows(query)] DMATCH) and `n_out` ([P] int32) may be caller-owned buffers that are
        reused across calls; the returned arrays are then views into `out`."""
        _kind_of(matcherType)
        P = len(trainDescs)
        cap = max(prevDesc.n, 1)
        own = out is None
        if own:
            out = np.empty((max(P, 1), cap), DMATCH)
            n_out = np.zeros(max(P, 1), np.int32)
        assert out.shape[0] >= P and out.shape[1] >= cap and out.dtype == DMATCH
        arr = (ctypes.c_void_p * max(P, 1))(*[t._h for t in trainDescs])
        check(self._lib.slamb200_match_batch(self._h, int(matcherType), prevDesc._h, arr, P,
                                             float(knnMatcherDistance), ptr(out), out.shape[1],
                                             ptr(n_out)))
        return [out[p, : n_out[p]].copy() if own else out[p, : n_out[p]] for p in range(P)]

    def matchBatchHost(self, prevDesc, trainDescs, matcherType, knnMatcherDistance=0.7, out=None, n_out=None):
        """The batch window from HOST descriptor Mats (numpy arrays, pageable is fine) in one call:
        uploads and matching overlap inside the library, nothing stays resident.  List of
        good-match arrays (views into `out` when the caller owns it)."""
        kind = _kind_of(matcherType)
        width, want = (128, np.float32) if kind == _capi.DESC_F32X128 else (32, np.uint8)

        def mat(a):
            a = np.asarray(a)
            if a.dtype != want or a.ndim != 2 or a.shape[1] != width:
                raise ValueError(f"descriptor Mat {a.dtype} {a.shape} does not fit the matcher")
            return a if a.strides[1] == a.itemsize else np.ascontiguousarray(a)
        q = mat(prevDesc)
        ts = [mat(t) for t in trainDescs]
        P = len(ts)
        cap = max(q.shape[0], 1)
        own = out is None
        if own:
            out = np.empty((max(P, 1), cap), DMATCH)
            n_out = np.zeros(max(P, 1), np.int32)
        ptrs = (ctypes.c_void_p * max(P, 1))(*[t.ctypes.data for t in ts])
        rows = np.array([t.shape[0] for t in ts] or [0], np.int32)
        strides = (ctypes.c_size_t * max(P, 1))(*[t.strides[0] if t.shape[0] > 1 else width * t.itemsize for t in ts])
        check(self._lib.slamb200_match_batch_host(self._h, int(matcherType), ptr(q), q.shape[0],
                                                  q.strides[0] if q.shape[0] > 1 else width * q.itemsize, ptrs, ptr(rows),
                                                  strides, P, float(knnMatcherDistance), ptr(out), out.shape[1],
                                                  ptr(n_out)))
        return [out[p, : n_out[p]].copy() if own else out[p, : n_out[p]] for p in range(P)]

    def matchWindow(self, frames, matcherType, knnMatcherDistance=0.7):
        """All i<j pairs of a frame window: dict {(i, j): matches}."""
        _kind_of(matcherType)
        F = len(frames)
        P = F * (F - 1) // 2
        cap = max([f.n for f in frames] + [1])
        out = np.zeros((max(P, 1), cap), DMATCH)
        n_out = np.zeros(max(P, 1), np.int32)
        arr = (ctypes.c_void_p * max(F, 1))(*[f._h for f in frames])
        check(self._lib.slamb200_match_window(self._h, int(matcherType), arr, F,
                                              float(knnMatcherDistance), ptr(out), cap, ptr(n_out)))
        res, p = {}, 0
        for i in range(F):
            for j in range(i + 1, F):
                res[(i, j)] = out[p, : n_out[p]].copy()
                p += 1
        return res

    # device-resident batch (bench / pipelines): enqueue on a stream, fetch later
    def matchBatchEnqueue(self, prevDesc, trainDescs, matcherType, knnMatcherDistance=0.7,
                          stream=None):
        _kind_of(matcherType)
        P = len(trainDescs)
        arr = (ctypes.c_void_p * max(P, 1))(*[t._h for t in trainDescs])
        check(self._lib.slamb200_match_batch_enqueue(self._h, int(matcherType), prevDesc._h, arr,
                                                     P, float(knnMatcherDistance),
                                                     ctypes.c_void_p(stream or 0)))
        self._last = (P, max(prevDesc.n, 1))

    def batchFetch(self, stream=None):
        P, cap = self._last
        out = np.empty((max(P, 1), cap), DMATCH)
        n_out = np.zeros(max(P, 1), np.int32)
        check(self._lib.slamb200_batch_fetch(self._h, ptr(out), cap, ptr(n_out),
                                             ctypes.c_void_p(stream or 0)))
        return [out[p, : n_out[p]] for p in range(P)], n_out[:P]


class DescriptorSet:
    """One frame's descriptors resident in HBM (slamb200_desc)."""

    def __init__(self, ctx, h, n, kind):
        self._ctx, self._h, self.n, self.kind = ctx, h, n, kind

    @property
    def exact_mode(self):
        return int(self._ctx._lib.slamb200_desc_exact_mode(self._h))

    def export_ipc(self):
        """128-byte record another rank of the node can pass to Context.import_ipc (the set must
        come from upload_shared and stay alive until the importers have freed their mappings)."""
        buf = (ctypes.c_ubyte * 128)()
        check(self._ctx._lib.slamb200_desc_export(self._ctx._h, self._h, ctypes.byref(buf)))
        return bytes(buf)

    def free(self):
        if self._h and self._ctx._h:
            self._ctx._lib.slamb200_free_desc(self._ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class KeypointSet:
    def __init__(self, ctx, h, n):
        self._ctx, self._h, self.n = ctx, h, n

    def free(self):
        if self._h and self._ctx._h:
            self._ctx._lib.slamb200_free_pts(self._ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def matchFramesPairFeatures(ctx, firstFrameDescriptor, secondDescriptor, matcherType,
                            knnMatcherDistance=0.7):
    """featureMatching.h:47-53 with the descriptor of the second frame already extracted
    (extractDescriptor stays OpenCV-CPU; it only produces this function's input): uploads both
    Mats, matches, returns the good matches."""
    kind = _kind_of(matcherType)
    a = ctx.upload(firstFrameDescriptor, kind)
    b = ctx.upload(secondDescriptor, kind)
    try:
        return ctx.matchFeatures(a, b, matcherType, knnMatcherDistance)
    finally:
        a.free()
        b.free()


def getKeyPointCoordsFromFramePair(prevFrameFeatures, nextFrameFeatures, matches):
    """featureMatchingCommon.cpp:23-33 (host gather; the device twin runs inside
    slamb200_score_batch_enqueue)."""
    prev = np.asarray(prevFrameFeatures, np.float32).reshape(-1, 2)
    nxt = np.asarray(nextFrameFeatures, np.float32).reshape(-1, 2)
    return prev[matches["queryIdx"]].copy(), nxt[matches["trainIdx"]].copy()
