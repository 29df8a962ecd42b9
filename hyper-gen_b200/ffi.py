"""ctypes binding of include/hypergen_b200.h — the stub a host language adds on its side.

Every function here is a 1:1 call into the C ABI; there is no Python or CPU implementation
of the hot path behind it.  If the shared library (or a CUDA device) is missing the calls
fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

HG_OK, HG_E_INVALID, HG_E_CUDA, HG_E_CAPACITY, HG_E_RANGE, HG_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

EXPORTS = [
    "hg_init", "hg_destroy", "hg_sync", "hg_host_alloc", "hg_host_free", "hg_host_register", "hg_host_unregister", "hg_last_error", "hg_version", "hg_stream_handle", "hg_launch_count",
    "hg_set_profiling", "hg_stage_ms", "hg_int_peak", "hg_tensor_peak", "hg_encode_sets", "hg_encode_sets_dev",
    "hg_fasta_merge", "hg_sketch_fasta_batch",
    "hg_kmer_hash", "hg_sketch_batch", "hg_sketch_batch_dev", "hg_sketch_status", "hg_unpack", "hg_unpack_dev",
    "hg_dist", "hg_dist_dev", "hg_dist_status", "hg_dist_last_path", "hg_dist_last_reason", "hg_sort_hits_dev", "hg_dist_sorted", "hg_dist_packed",
    "hg_group_create", "hg_group_destroy", "hg_group_size", "hg_group_ctx", "hg_group_sketch_fasta_batch", "hg_group_dist_packed",
    "hg_peer_window_need", "hg_peer_create", "hg_peer_connect", "hg_peer_create_local", "hg_peer_destroy", "hg_peer_rank",
    "hg_peer_world", "hg_peer_barrier", "hg_peer_stage_ms", "hg_peer_timeline", "hg_peer_plan_tiles", "hg_peer_plan_push", "hg_dist_sharded_dev", "hg_dist_sharded_hits", "hg_peer_hit_buffers",
]
HG_MAX_PEERS, HG_IPC_HANDLE_BYTES = 8, 64


class SketchParams(C.Structure):
    """struct hg_sketch_params (reference knobs: src/types.rs:97-113, FileSketch fields)."""
    _fields_ = [("scaled", C.c_uint64), ("seed", C.c_uint64), ("hv_d", C.c_uint32), ("ksize", C.c_uint8),
                ("canonical", C.c_uint8), ("reserved", C.c_uint8 * 2)]


HIT_DTYPE = np.dtype([("i", np.uint32), ("j", np.uint32), ("dot", np.int32), ("ani", np.float32)])


class HyperGenError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("hypergen_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """dlopen the in-tree library (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("HG_LIB")  # kernel-variant experiments only
    if override:
        L = C.CDLL(override)
    else:
        if _build.needs_build():
            _build.build()
        L = C.CDLL(lib_path())
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    pp = C.POINTER(SketchParams)
    L.hg_init.restype = i32; L.hg_init.argtypes = [i32, C.POINTER(vp)]
    L.hg_destroy.restype = None; L.hg_destroy.argtypes = [vp]
    L.hg_sync.restype = i32; L.hg_sync.argtypes = [vp]
    L.hg_host_alloc.restype = i32; L.hg_host_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
    L.hg_host_free.restype = i32; L.hg_host_free.argtypes = [vp]
    L.hg_host_register.restype = i32; L.hg_host_register.argtypes = [vp, C.c_uint64, C.POINTER(C.c_void_p)]
    L.hg_host_unregister.restype = i32; L.hg_host_unregister.argtypes = [vp]
    L.hg_last_error.restype = C.c_char_p; L.hg_last_error.argtypes = []
    L.hg_version.restype = C.c_char_p; L.hg_version.argtypes = []
    L.hg_stream_handle.restype = u64; L.hg_stream_handle.argtypes = [vp]
    L.hg_launch_count.restype = u64; L.hg_launch_count.argtypes = [vp]
    L.hg_set_profiling.restype = i32; L.hg_set_profiling.argtypes = [vp, i32]
    L.hg_stage_ms.restype = i32; L.hg_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.hg_int_peak.restype = i32; L.hg_int_peak.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.hg_tensor_peak.restype = i32; L.hg_tensor_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.hg_encode_sets.restype = i32; L.hg_encode_sets.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp, vp]
    L.hg_encode_sets_dev.restype = i32; L.hg_encode_sets_dev.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp, vp]
    L.hg_fasta_merge.restype = i32; L.hg_fasta_merge.argtypes = [vp, vp, vp, u32, vp, u64, vp]
    L.hg_sketch_fasta_batch.restype = i32; L.hg_sketch_fasta_batch.argtypes = [vp, vp, vp, u32, pp, vp, vp, vp, vp, vp]
    L.hg_kmer_hash.restype = i32; L.hg_kmer_hash.argtypes = [vp, vp, vp, u32, pp, vp, u64, vp]
    L.hg_sketch_batch.restype = i32; L.hg_sketch_batch.argtypes = [vp, vp, vp, u32, pp, vp, vp, vp, vp, vp]
    L.hg_sketch_batch_dev.restype = i32; L.hg_sketch_batch_dev.argtypes = [vp, vp, vp, u32, pp, vp, vp, vp, vp, vp]
    L.hg_sketch_status.restype = i32; L.hg_sketch_status.argtypes = [vp]
    L.hg_unpack.restype = i32; L.hg_unpack.argtypes = [vp, vp, u64, vp, u32, u32, vp]
    L.hg_unpack_dev.restype = i32; L.hg_unpack_dev.argtypes = [vp, vp, u64, vp, u32, u32, vp]
    L.hg_dist.restype = i32
    L.hg_dist.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, u32, C.c_float, i32, i32, vp, u64, C.POINTER(u64)]
    L.hg_dist_dev.restype = i32
    L.hg_dist_dev.argtypes = [vp, vp, vp, u32, u32, vp, vp, u32, u32, u32, u32, C.c_float, i32, i32, vp, u64, vp]
    L.hg_dist_status.restype = i32; L.hg_dist_status.argtypes = [vp]
    L.hg_sort_hits_dev.restype = i32; L.hg_sort_hits_dev.argtypes = [vp, vp, u64, vp]
    L.hg_dist_sorted.restype = i32
    L.hg_dist_sorted.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, u32, C.c_float, i32, i32, vp, vp, u64, C.POINTER(u64)]
    L.hg_dist_packed.restype = i32
    L.hg_dist_packed.argtypes = [vp, vp, u64, vp, vp, u32, vp, u64, vp, vp, u32, u32, u32, C.c_float, i32, i32, i32, vp, vp,
                                 u64, C.POINTER(u64)]
    L.hg_group_create.restype = i32; L.hg_group_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.hg_group_destroy.restype = None; L.hg_group_destroy.argtypes = [vp]
    L.hg_group_size.restype = i32; L.hg_group_size.argtypes = [vp]
    L.hg_group_ctx.restype = vp; L.hg_group_ctx.argtypes = [vp, i32]
    L.hg_group_sketch_fasta_batch.restype = i32; L.hg_group_sketch_fasta_batch.argtypes = [vp, vp, vp, u32, pp, vp, vp, vp, vp, vp]
    L.hg_group_dist_packed.restype = i32
    L.hg_group_dist_packed.argtypes = [vp, vp, u64, vp, vp, u32, vp, u64, vp, vp, u32, u32, u32, C.c_float, i32, i32, vp, vp, u64,
                                       C.POINTER(u64)]
    L.hg_peer_window_need.restype = u64; L.hg_peer_window_need.argtypes = [u32, u32, u64]
    L.hg_peer_create.restype = i32; L.hg_peer_create.argtypes = [vp, i32, i32, u64, vp, C.POINTER(vp)]
    L.hg_peer_connect.restype = i32; L.hg_peer_connect.argtypes = [vp, vp]
    L.hg_peer_create_local.restype = i32; L.hg_peer_create_local.argtypes = [vp, i32, u64, vp]
    L.hg_peer_destroy.restype = None; L.hg_peer_destroy.argtypes = [vp]
    L.hg_peer_rank.restype = i32; L.hg_peer_rank.argtypes = [vp]
    L.hg_peer_world.restype = i32; L.hg_peer_world.argtypes = [vp]
    L.hg_peer_barrier.restype = i32; L.hg_peer_barrier.argtypes = [vp]
    L.hg_peer_plan_tiles.restype = i32; L.hg_peer_plan_tiles.argtypes = [i32, i32, i32, i32, u32, u32, vp, vp, u64, C.POINTER(u64)]
    L.hg_peer_plan_push.restype = i32; L.hg_peer_plan_push.argtypes = [i32, i32, i32, i32, u32, vp, vp, vp, u64, C.POINTER(u64), C.POINTER(C.c_int)]
    L.hg_peer_stage_ms.restype = i32; L.hg_peer_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.hg_peer_timeline.restype = i32; L.hg_peer_timeline.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.hg_dist_sharded_dev.restype = i32
    L.hg_dist_sharded_dev.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp, u32, u32, C.c_float, i32, i32, i32, u64, vp]
    L.hg_dist_sharded_hits.restype = i32; L.hg_dist_sharded_hits.argtypes = [vp, i32, vp, vp, u64, C.POINTER(u64)]
    L.hg_peer_hit_buffers.restype = i32; L.hg_peer_hit_buffers.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    L.hg_dist_last_path.restype = i32; L.hg_dist_last_path.argtypes = [vp]
    L.hg_dist_last_reason.restype = C.c_char_p; L.hg_dist_last_reason.argtypes = [vp]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != HG_OK:
        raise HyperGenError(rc, load().hg_last_error().decode("utf-8", "replace"))


def _ptr(a) -> int:
    """host numpy array / device pointer int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def make_params(k=21, scaled=1500, seed=123, canonical=True, hv_d=4096) -> SketchParams:
    return SketchParams(scaled=scaled, seed=seed, hv_d=hv_d, ksize=k, canonical=int(bool(canonical)))


class Context:
    """hg_ctx: one CUDA device + stream (replaces CudaDevice::new(0), sketch_cuda.rs:52)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(load().hg_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            load().hg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing --
    def sync(self):
        _check(load().hg_sync(self._h))

    @property
    def stream(self) -> int:
        return int(load().hg_stream_handle(self._h))

    @property
    def launches(self) -> int:
        return int(load().hg_launch_count(self._h))

    def set_profiling(self, enabled: bool):
        _check(load().hg_set_profiling(self._h, int(enabled)))

    def stage_ms(self):
        """device ms of the last call's stages: [staging, kmer_hash, encode, dist] (-1 = not run)"""
        out = (C.c_float * 4)()
        _check(load().hg_stage_ms(self._h, out))
        return [float(x) for x in out]

    def int_peak(self, which: int = 2) -> float:
        v = C.c_double(0)
        _check(load().hg_int_peak(self._h, which, C.byref(v)))
        return v.value

    def tensor_peak(self) -> float:
        v = C.c_double(0)
        _check(load().hg_tensor_peak(self._h, C.byref(v)))
        return v.value

    # -- stage 1 --
    def kmer_hash(self, seq: np.ndarray, seg_off: np.ndarray, params: SketchParams):
        """hg_kmer_hash: per-genome sorted unique sampled hashes -> (hashes, hash_off)."""
        seq = np.ascontiguousarray(seq, np.uint8)
        off = np.ascontiguousarray(seg_off, np.uint64)
        n = off.size - 1
        hash_off = np.zeros(n + 1, np.uint64)
        total_len = int(off[-1] - off[0]) if n else 0
        cap = total_len // max(int(params.scaled), 1) * 2 + 4096 * max(n, 1)
        hashes = np.empty(cap, np.uint64)
        rc = load().hg_kmer_hash(self._h, _ptr(seq) if seq.size else None, _ptr(off), n, C.byref(params),
                                 _ptr(hashes), cap, _ptr(hash_off))
        if rc == HG_E_CAPACITY and int(hash_off[-1]) > cap:
            cap = int(hash_off[-1])
            hashes = np.empty(cap, np.uint64)
            rc = load().hg_kmer_hash(self._h, _ptr(seq), _ptr(off), n, C.byref(params), _ptr(hashes), cap,
                                     _ptr(hash_off))
        _check(rc)
        return hashes[: int(hash_off[-1])].copy(), hash_off

    def sketch_batch(self, seq: np.ndarray, seg_off: np.ndarray, params: SketchParams, want_hv: bool = True):
        """hg_sketch_batch with host buffers -> dict(hv, packed, quant_bits, norm2, n_hashes)."""
        seq = np.ascontiguousarray(seq, np.uint8)
        off = np.ascontiguousarray(seg_off, np.uint64)
        n = off.size - 1
        D = int(params.hv_d)
        hv = np.empty((n, D), np.int16) if want_hv else None
        packed = np.empty((n, 2 * D), np.uint8)
        qb = np.empty(n, np.uint8)
        norm2 = np.empty(n, np.int32)
        nh = np.empty(n, np.uint32)
        _check(load().hg_sketch_batch(self._h, _ptr(seq) if seq.size else None, _ptr(off), n, C.byref(params),
                                      _ptr(hv), _ptr(packed), _ptr(qb), _ptr(norm2), _ptr(nh)))
        return dict(hv=hv, packed=packed, quant_bits=qb, norm2=norm2, n_hashes=nh)

    def sketch_batch_dev(self, d_seq: int, seg_off: np.ndarray, params: SketchParams, d_hv, d_packed,
                         d_quant_bits: int, d_norm2: int, d_n_hashes: int):
        """hg_sketch_batch_dev: device pointers (ints), asynchronous on self.stream."""
        off = np.ascontiguousarray(seg_off, np.uint64)
        _check(load().hg_sketch_batch_dev(self._h, d_seq, _ptr(off), off.size - 1, C.byref(params), d_hv, d_packed,
                                          d_quant_bits, d_norm2, d_n_hashes))

    def encode_sets(self, sets, hv_d: int = 4096, want_hv: bool = True):
        """hg_encode_sets: list of unique-u64 arrays -> dict(hv, packed, quant_bits, norm2)."""
        n = len(sets)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in sets])
        hashes = np.ascontiguousarray(np.concatenate(sets) if n and int(off[-1]) else np.zeros(0), np.uint64)
        hv = np.empty((n, hv_d), np.int16) if want_hv else None
        packed = np.empty((n, 2 * hv_d), np.uint8)
        qb = np.empty(n, np.uint8)
        norm2 = np.empty(n, np.int32)
        _check(load().hg_encode_sets(self._h, _ptr(hashes) if hashes.size else None, _ptr(off), n, hv_d, _ptr(hv),
                                     _ptr(packed), _ptr(qb), _ptr(norm2)))
        return dict(hv=hv, packed=packed, quant_bits=qb, norm2=norm2)

    def encode_sets_dev(self, d_hashes: int, hash_off: np.ndarray, hv_d: int, d_hv, d_packed, d_quant_bits, d_norm2):
        off = np.ascontiguousarray(hash_off, np.uint64)
        _check(load().hg_encode_sets_dev(self._h, d_hashes, _ptr(off), off.size - 1, hv_d, d_hv, d_packed,
                                         d_quant_bits, d_norm2))

    # -- raw FASTA in --
    @staticmethod
    def _concat_files(files):
        off = np.zeros(len(files) + 1, np.uint64)
        off[1:] = np.cumsum([len(f) for f in files])
        raw = np.frombuffer(b"".join(bytes(f) for f in files), dtype=np.uint8) if int(off[-1]) else np.zeros(0, np.uint8)
        return np.ascontiguousarray(raw), off

    def fasta_merge(self, files):
        """hg_fasta_merge: list of raw FASTA file contents (bytes) -> list of merged sequences (uint8 arrays)."""
        raw, off = self._concat_files(files)
        n = len(files)
        merged = np.empty(max(int(off[-1]), 1), np.uint8)
        moff = np.zeros(n + 1, np.uint64)
        _check(load().hg_fasta_merge(self._h, _ptr(raw) if raw.size else None, _ptr(off), n, _ptr(merged), merged.size,
                                     _ptr(moff)))
        return [merged[int(moff[f]):int(moff[f + 1])].copy() for f in range(n)]

    def sketch_fasta_batch(self, files, params: SketchParams, want_hv: bool = True):
        """hg_sketch_fasta_batch: raw FASTA file contents in, sketches out (as sketch_batch)."""
        raw, off = self._concat_files(files)
        n = len(files)
        D = int(params.hv_d)
        hv = np.empty((n, D), np.int16) if want_hv else None
        packed = np.empty((n, 2 * D), np.uint8)
        qb = np.empty(n, np.uint8)
        norm2 = np.empty(n, np.int32)
        nh = np.empty(n, np.uint32)
        _check(load().hg_sketch_fasta_batch(self._h, _ptr(raw) if raw.size else None, _ptr(off), n, C.byref(params),
                                            _ptr(hv), _ptr(packed), _ptr(qb), _ptr(norm2), _ptr(nh)))
        return dict(hv=hv, packed=packed, quant_bits=qb, norm2=norm2, n_hashes=nh)

    def sketch_status(self):
        _check(load().hg_sketch_status(self._h))

    # -- format --
    def unpack(self, packed: np.ndarray, quant_bits: np.ndarray, hv_d: int) -> np.ndarray:
        packed = np.ascontiguousarray(packed, np.uint8)
        qb = np.ascontiguousarray(quant_bits, np.uint8)
        n = qb.size
        hv = np.empty((n, hv_d), np.int16)
        _check(load().hg_unpack(self._h, _ptr(packed), packed.shape[1], _ptr(qb), n, hv_d, _ptr(hv)))
        return hv

    def unpack_dev(self, d_packed: int, row_stride: int, d_quant_bits: int, n: int, hv_d: int, d_hv: int):
        _check(load().hg_unpack_dev(self._h, d_packed, row_stride, d_quant_bits, n, hv_d, d_hv))

    # -- stage 2 --
    def dist(self, ref_hv, ref_norm2, qry_hv, qry_norm2, ksize=21, ani_th=85.0, symmetric=False, path=0,
             cap=None, sorted_output=False, want_milli=False):
        """hg_dist with host buffers -> structured array of hits (i, j, dot, ani), unspecified order.
        sorted_output: hg_dist_sorted instead - the hits in the reference's output order (sorted on the
        GPU); with want_milli also the ANI in thousandths as `{:.3}` rounds it -> (hits, milli)."""
        r = np.ascontiguousarray(ref_hv, np.int16)
        rn = np.ascontiguousarray(ref_norm2, np.int32)
        same = qry_hv is ref_hv
        q = r if same else np.ascontiguousarray(qry_hv, np.int16)
        qn = rn if same and qry_norm2 is ref_norm2 else np.ascontiguousarray(qry_norm2, np.int32)
        R, D = r.shape
        Q = q.shape[0]
        if cap is None:
            cap = max(1024, R * Q // 64)
        while True:
            hits = np.empty(cap, HIT_DTYPE)
            n_hits = C.c_uint64(0)
            milli = np.empty(cap, np.uint32) if (sorted_output and want_milli) else None
            if sorted_output:
                rc = load().hg_dist_sorted(self._h, _ptr(r), _ptr(rn), R, _ptr(q), _ptr(qn), Q, D, ksize, ani_th,
                                           int(symmetric), path, _ptr(hits), _ptr(milli), cap, C.byref(n_hits))
            else:
                rc = load().hg_dist(self._h, _ptr(r), _ptr(rn), R, _ptr(q), _ptr(qn), Q, D, ksize, ani_th,
                                    int(symmetric), path, _ptr(hits), cap, C.byref(n_hits))
            if rc == HG_E_CAPACITY and n_hits.value > cap:
                cap = int(n_hits.value)
                continue
            _check(rc)
            if milli is not None:
                return hits[: n_hits.value].copy(), milli[: n_hits.value].copy()
            return hits[: n_hits.value].copy()

    def dist_dev(self, d_ref, d_ref_norm2, n_ref, i0, d_qry, d_qry_norm2, n_qry, j0, hv_d, ksize, ani_th,
                 symmetric, path, d_hits, cap, d_n_hits):
        _check(load().hg_dist_dev(self._h, d_ref, d_ref_norm2, n_ref, i0, d_qry, d_qry_norm2, n_qry, j0, hv_d, ksize,
                                  ani_th, int(symmetric), path, d_hits, cap, d_n_hits))

    def dist_status(self):
        """hg_dist_status: the verdict a dist_dev(path=3) did not wait for"""
        _check(load().hg_dist_status(self._h))

    def dist_packed(self, ref_packed, ref_bits, ref_norm2, qry_packed, qry_bits, qry_norm2, hv_d, ksize=21, ani_th=85.0,
                    symmetric=False, path=0, sorted_output=True, cap=None):
        """hg_dist_packed: packed sketch rows (n x stride uint8) + quant bits + norms -> (hits, ani_milli)."""
        rp = np.ascontiguousarray(ref_packed, np.uint8)
        rb = np.ascontiguousarray(ref_bits, np.uint8)
        rn = np.ascontiguousarray(ref_norm2, np.int32)
        same = qry_packed is ref_packed
        qp = rp if same else np.ascontiguousarray(qry_packed, np.uint8)
        qb = rb if same else np.ascontiguousarray(qry_bits, np.uint8)
        qn = rn if same else np.ascontiguousarray(qry_norm2, np.int32)
        R, Q = rp.shape[0], qp.shape[0]
        if cap is None:
            cap = max(1024, R * Q // 64)
        while True:
            hits = np.empty(cap, HIT_DTYPE)
            milli = np.empty(cap, np.uint32)
            n_hits = C.c_uint64(0)
            rc = load().hg_dist_packed(self._h, _ptr(rp), rp.strides[0] if R else 0, _ptr(rb), _ptr(rn), R, _ptr(qp),
                                       qp.strides[0] if Q else 0, _ptr(qb), _ptr(qn), Q, hv_d, ksize, ani_th,
                                       int(symmetric), path, int(sorted_output), _ptr(hits), _ptr(milli), cap,
                                       C.byref(n_hits))
            if rc == HG_E_CAPACITY and n_hits.value > cap:
                cap = int(n_hits.value)
                continue
            _check(rc)
            # ani_milli is written by the output-stage sort only (hypergen_b200.h): unsorted calls get None
            return hits[: n_hits.value].copy(), (milli[: n_hits.value].copy() if sorted_output else None)

    def sort_hits_dev(self, d_hits, n, d_ani_milli=None):
        """hg_sort_hits_dev: device records to the reference's output order, in place."""
        _check(load().hg_sort_hits_dev(self._h, d_hits, n, d_ani_milli))

    @property
    def dist_last_path(self) -> int:
        return int(load().hg_dist_last_path(self._h))

    @property
    def dist_last_reason(self) -> str:
        return load().hg_dist_last_reason(self._h).decode()


class Peer:
    """hg_peer: this process's GPU as one member of a group of GPUs on the box (one process per GPU).

    `exchange(handle: bytes) -> list[bytes]` is the host's own all-gather of the 64-byte window handles
    (multigpu.PeerGroup passes one built on torch.distributed); world == 1 needs none."""

    def __init__(self, ctx: Context, rank: int, world: int, window_bytes: int, exchange=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        handle = (C.c_uint8 * HG_IPC_HANDLE_BYTES)()
        _check(load().hg_peer_create(ctx._h, rank, world, window_bytes, handle, C.byref(self._h)))
        if world > 1:
            allh = exchange(bytes(handle))
            buf = (C.c_uint8 * (HG_IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(allh))
            _check(load().hg_peer_connect(self._h, buf))

    def close(self):
        if self._h:
            load().hg_peer_destroy(self._h)
            self._h = C.c_void_p()

    def barrier(self):
        _check(load().hg_peer_barrier(self._h))

    def dist_sharded_dev(self, d_ref, d_ref_norm2, n_ref_local, ref_row0, d_qry, d_qry_norm2, qry_bounds, hv_d, ksize, ani_th,
                         symmetric, path, root, cap, mapped_hits=None):
        """hg_dist_sharded_dev (collective; device pointers as ints; qry_bounds: world + 1 row offsets)."""
        qb = np.ascontiguousarray(qry_bounds, np.uint32)
        assert qb.size == self.world + 1
        _check(load().hg_dist_sharded_dev(self._h, d_ref, d_ref_norm2, n_ref_local, ref_row0, d_qry, d_qry_norm2, _ptr(qb), hv_d,
                                          ksize, ani_th, int(symmetric), path, root, cap, mapped_hits))

    def dist_sharded_hits(self, cap: int, sorted_output: bool = False, hits=None, milli=None):
        """hg_dist_sharded_hits -> (hits, milli or None) on the root, (empty, None) elsewhere.  `hits` / `milli`
        may be preallocated (pinned) numpy arrays of at least `cap` records."""
        own = hits is None
        if own:
            hits = np.empty(cap, HIT_DTYPE)
        if sorted_output and milli is None:
            milli = np.empty(cap, np.uint32)
        n = C.c_uint64(0)
        _check(load().hg_dist_sharded_hits(self._h, int(sorted_output), _ptr(hits), _ptr(milli) if sorted_output else None, cap,
                                           C.byref(n)))
        h = hits[: n.value]
        return (h.copy() if own else h), (milli[: n.value] if sorted_output else None)

    def stage_ms(self):
        """device ms of the last sharded dist: [operands, chunked push, kernel incl. waits for the peers' chunks, final barrier]"""
        out = (C.c_float * 4)()
        _check(load().hg_peer_stage_ms(self._h, out))
        return [float(x) for x in out]

    def timeline(self):
        """globaltimer stamps (ns) of the last sharded dist (HG_PEER_TIMELINE=1); see hg_peer_timeline"""
        out = (C.c_uint64 * 32)()
        _check(load().hg_peer_timeline(self._h, out))
        return [int(x) for x in out]

    def hit_buffers(self, root: int = 0):
        dh, dc = C.c_void_p(), C.c_void_p()
        _check(load().hg_peer_hit_buffers(self._h, root, C.byref(dh), C.byref(dc)))
        return dh.value, dc.value


def host_register(arr: np.ndarray) -> int:
    """hg_host_register on a numpy array's memory -> the device pointer kernels of this process use for it"""
    dev = C.c_void_p()
    _check(load().hg_host_register(arr.ctypes.data, arr.nbytes, C.byref(dev)))
    return dev.value


def host_unregister(arr: np.ndarray) -> None:
    _check(load().hg_host_unregister(arr.ctypes.data))


def peer_plan_tiles(world: int, rank: int, symmetric: bool, path: int, n_ref_local: int, qry_bounds, hv_d: int = 4096):
    """hg_peer_plan_tiles -> (tile_row, tile_col, need_mask) arrays in walk order (host logic, no GPU)"""
    qb = np.ascontiguousarray(qry_bounds, np.uint32)
    n = C.c_uint64(0)
    load().hg_peer_plan_tiles(world, rank, int(symmetric), path, hv_d, n_ref_local, _ptr(qb), None, 0, C.byref(n))
    out = np.zeros((max(int(n.value), 1), 2), np.uint32)
    _check(load().hg_peer_plan_tiles(world, rank, int(symmetric), path, hv_d, n_ref_local, _ptr(qb), _ptr(out), out.shape[0], C.byref(n)))
    out = out[: n.value]
    return out[:, 0] & 0xFFFF, out[:, 0] >> 16, out[:, 1]


def peer_plan_push(world: int, rank: int, symmetric: bool, path: int, qry_bounds, hv_d: int = 4096):
    """hg_peer_plan_push -> (chunk row boundaries [5], [(set, destination mask)] in sending order, ring?)"""
    qb = np.ascontiguousarray(qry_bounds, np.uint32)
    rows = np.zeros(5, np.uint32)
    units = np.zeros((40, 2), np.uint32)
    n, ring = C.c_uint64(0), C.c_int(0)
    _check(load().hg_peer_plan_push(world, rank, int(symmetric), path, hv_d, _ptr(qb), _ptr(rows), _ptr(units), units.shape[0],
                                    C.byref(n), C.byref(ring)))
    return rows.tolist(), [(int(a), int(b)) for a, b in units[: n.value]], bool(ring.value)


def peer_window_need(gathered_rows: int, hv_d: int, hit_cap: int) -> int:
    return int(load().hg_peer_window_need(gathered_rows, hv_d, hit_cap))


class Group:
    """hg_group: one process driving several GPUs of the box (what the `hyper-gen` CLI uses)."""

    def __init__(self, n_devices: int = 0, ordinals=None):
        self._h = C.c_void_p()
        arr = None
        if ordinals is not None:
            arr = (C.c_int * len(ordinals))(*ordinals)
            n_devices = len(ordinals)
        _check(load().hg_group_create(n_devices, arr, C.byref(self._h)))

    def close(self):
        if self._h:
            load().hg_group_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def size(self) -> int:
        return int(load().hg_group_size(self._h))

    def dist_last_path(self, i: int = 0) -> int:
        return int(load().hg_dist_last_path(load().hg_group_ctx(self._h, i)))

    def sketch_fasta_batch(self, files, params: SketchParams, want_hv: bool = True):
        raw, off = Context._concat_files(files)
        n = len(files)
        D = int(params.hv_d)
        hv = np.empty((n, D), np.int16) if want_hv else None
        packed = np.empty((n, 2 * D), np.uint8)
        qb = np.empty(n, np.uint8)
        norm2 = np.empty(n, np.int32)
        nh = np.empty(n, np.uint32)
        _check(load().hg_group_sketch_fasta_batch(self._h, _ptr(raw) if raw.size else None, _ptr(off), n, C.byref(params),
                                                  _ptr(hv), _ptr(packed), _ptr(qb), _ptr(norm2), _ptr(nh)))
        return dict(hv=hv, packed=packed, quant_bits=qb, norm2=norm2, n_hashes=nh)

    def dist_packed(self, ref_packed, ref_bits, ref_norm2, qry_packed, qry_bits, qry_norm2, hv_d, ksize=21, ani_th=85.0,
                    symmetric=False, sorted_output=True, cap=None):
        rp = np.ascontiguousarray(ref_packed, np.uint8)
        rb = np.ascontiguousarray(ref_bits, np.uint8)
        rn = np.ascontiguousarray(ref_norm2, np.int32)
        same = qry_packed is ref_packed
        qp = rp if same else np.ascontiguousarray(qry_packed, np.uint8)
        qb = rb if same else np.ascontiguousarray(qry_bits, np.uint8)
        qn = rn if same else np.ascontiguousarray(qry_norm2, np.int32)
        R, Q = rp.shape[0], qp.shape[0]
        if cap is None:
            cap = max(1024, R * Q // 64)
        while True:
            hits = np.empty(cap, HIT_DTYPE)
            milli = np.empty(cap, np.uint32)
            n_hits = C.c_uint64(0)
            rc = load().hg_group_dist_packed(self._h, _ptr(rp), rp.strides[0] if R else 0, _ptr(rb), _ptr(rn), R, _ptr(qp),
                                             qp.strides[0] if Q else 0, _ptr(qb), _ptr(qn), Q, hv_d, ksize, ani_th,
                                             int(symmetric), int(sorted_output), _ptr(hits), _ptr(milli), cap, C.byref(n_hits))
            if rc == HG_E_CAPACITY and n_hits.value > cap:
                cap = int(n_hits.value)
                continue
            _check(rc)
            return hits[: n_hits.value].copy(), (milli[: n_hits.value].copy() if sorted_output else None)
