"""CPU oracle for the HyperGen sketch -> dist hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``hyper-gen_b200/``) never
does.  It wraps ``oracle/hg_oracle.c`` (a C restatement of the reference's CPU algorithm,
each function citing reference file:line) and, when present, ``oracle/_ref/libhgref.so``
(the reference's own ``src/cuda_kernel.cu`` compiled as host code by ``oracle/Makefile``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libhg_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libhgref.so")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is mounted)."""
    src = os.path.join(_HERE, "hg_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libhg_oracle.so"])
    have_ref = os.path.exists("/root/reference/src/cuda_kernel.cu")
    if have_ref and (force or not os.path.exists(_REF_PATH)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    # the reference's GPU kernel as PTX (its own build.rs recipe) and as an sm_100a cubin, for bench.py's
    # `reference_gpu_kernel` leg on the GPU box (oracle/ref_gpu.py)
    if have_ref and (force or not os.path.exists(os.path.join(_HERE, "_ref", "cuda_kernel_sm100a.cubin"))):
        subprocess.call(["make", "-s", "-C", _HERE, "ref_gpu"], stderr=subprocess.DEVNULL)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.hgo_t1ha2_atonce.restype = C.c_uint64
        L.hgo_t1ha2_atonce.argtypes = [_u8p, C.c_uint64, C.c_uint64]
        L.hgo_read_merge_seq.restype = C.c_uint64
        L.hgo_read_merge_seq.argtypes = [_u8p, C.c_uint64, C.c_void_p]
        L.hgo_kmer_hash_set.restype = C.c_uint64
        L.hgo_kmer_hash_set.argtypes = [_u8p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int,
                                        C.c_void_p, C.c_uint64]
        L.hgo_wyrng_next.restype = C.c_uint64
        L.hgo_wyrng_next.argtypes = [C.POINTER(C.c_uint64)]
        L.hgo_encode_hd.restype = None
        L.hgo_encode_hd.argtypes = [_u64p, C.c_uint64, C.c_uint32, C.c_int, _i16p]
        L.hgo_encode_hd_avx2.restype = None
        L.hgo_encode_hd_avx2.argtypes = [_u64p, C.c_uint64, C.c_uint32, _i16p]
        L.hgo_hv_l2_norm_sq.restype = C.c_int32
        L.hgo_hv_l2_norm_sq.argtypes = [_i16p, C.c_uint32]
        L.hgo_quant_bits.restype = C.c_uint32
        L.hgo_quant_bits.argtypes = [_i16p, C.c_uint32]
        L.hgo_compress_hd_sketch.restype = C.c_uint32
        L.hgo_compress_hd_sketch.argtypes = [_i16p, C.c_uint32, _u8p]
        L.hgo_decompress_hd_sketch.restype = None
        L.hgo_decompress_hd_sketch.argtypes = [_u8p, C.c_uint32, C.c_uint32, _i16p]
        L.hgo_dot_i16.restype = C.c_int32
        L.hgo_dot_i16.argtypes = [_i16p, _i16p, C.c_uint32]
        L.hgo_ani_from_dot.restype = C.c_float
        L.hgo_ani_from_dot.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint32]
        L.hgo_logf_glibc.restype = C.c_float
        L.hgo_logf_glibc.argtypes = [C.c_float]
        L.hgo_logf_selfcheck.restype = C.c_uint64
        L.hgo_logf_selfcheck.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.hgo_set_threads.restype = C.c_int
        L.hgo_set_threads.argtypes = [C.c_int]
        L.hgo_sketch_batch.restype = None
        L.hgo_sketch_batch.argtypes = [_u8p, _u64p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int,
                                       C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.hgo_kmer_count_batch.restype = C.c_uint64
        L.hgo_kmer_count_batch.argtypes = [_u8p, _u64p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                           C.c_int]
        L.hgo_dist_all.restype = C.c_uint64
        L.hgo_dist_all.argtypes = [_i16p, _i32p, C.c_uint32, _i16p, _i32p, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_int, _f32p, C.c_void_p]
        L.hgo_ani_output_order.restype = C.c_uint64
        L.hgo_ani_output_order.argtypes = [_f32p, C.c_uint64, C.c_float, _u64p]
        _lib = L
    return _lib


def ref():
    """The reference's own kernel source compiled as host code, or None if not built."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF_PATH):
            return None
        R = C.CDLL(_REF_PATH)
        R.hgref_t1ha2_atonce.restype = C.c_uint64
        R.hgref_t1ha2_atonce.argtypes = [_u8p, C.c_uint64, C.c_uint64]
        R.hgref_extract_kmer_t1ha2.restype = C.c_uint64
        R.hgref_extract_kmer_t1ha2.argtypes = [_u8p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64,
                                               C.c_int, C.c_void_p, C.c_uint64]
        _ref = R
    return _ref


def _bytes(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray)):
        x = np.frombuffer(bytes(x), dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.uint8)
    if x.size == 0:
        x = np.zeros(1, np.uint8)[:0].copy()
    return x


def _pad(x: np.ndarray) -> np.ndarray:
    # ndpointer rejects nothing about size, but a 0-length array may have a NULL base
    return x if x.size else np.zeros(1, x.dtype)


def t1ha2_atonce(data, seed: int) -> int:
    d = _bytes(data)
    return int(lib().hgo_t1ha2_atonce(_pad(d), d.size, seed))


def ref_t1ha2_atonce(data, seed: int) -> int:
    d = _bytes(data)
    return int(ref().hgref_t1ha2_atonce(_pad(d), d.size, seed))


def read_merge_seq(file_bytes) -> np.ndarray:
    f = _bytes(file_bytes)
    n = lib().hgo_read_merge_seq(_pad(f), f.size, None)
    out = np.empty(max(int(n), 1), np.uint8)
    lib().hgo_read_merge_seq(_pad(f), f.size, out.ctypes.data)
    return out[: int(n)]


def kmer_hash_set(seq, k=21, scaled=1500, seed=123, canonical=True) -> np.ndarray:
    s = _bytes(seq)
    cap = s.size // max(scaled, 1) * 2 + 4096
    out = np.empty(cap, np.uint64)
    n = int(lib().hgo_kmer_hash_set(_pad(s), s.size, k, scaled, seed, int(canonical), out.ctypes.data, cap))
    if n > cap:
        out = np.empty(n, np.uint64)
        n = int(lib().hgo_kmer_hash_set(_pad(s), s.size, k, scaled, seed, int(canonical), out.ctypes.data, n))
    return out[:n].copy()


def ref_kmer_hash_set(seq, k=21, scaled=1500, seed=123, canonical=True) -> np.ndarray:
    s = _bytes(seq)
    cap = s.size // max(scaled, 1) * 2 + 4096
    out = np.empty(cap, np.uint64)
    n = int(ref().hgref_extract_kmer_t1ha2(_pad(s), s.size, k, scaled, seed, int(canonical), out.ctypes.data, cap))
    assert n <= cap
    return out[:n].copy()


def wyrng_words(state: int, n: int) -> list[int]:
    st = C.c_uint64(state)
    return [int(lib().hgo_wyrng_next(C.byref(st))) for _ in range(n)]


def encode_hd(hashes, hv_d=4096, layout="avx2") -> np.ndarray:
    h = np.ascontiguousarray(hashes, dtype=np.uint64)
    hv = np.empty(hv_d, np.int16)
    lib().hgo_encode_hd(_pad(h), h.size, hv_d, 0 if layout == "avx2" else 1, hv)
    return hv


def encode_hd_avx2_intrinsics(hashes, hv_d=4096) -> np.ndarray:
    h = np.ascontiguousarray(hashes, dtype=np.uint64)
    hv = np.empty(hv_d, np.int16)
    lib().hgo_encode_hd_avx2(_pad(h), h.size, hv_d, hv)
    return hv


def hv_l2_norm_sq(hv) -> int:
    hv = np.ascontiguousarray(hv, dtype=np.int16)
    return int(lib().hgo_hv_l2_norm_sq(hv, hv.size))


def compress_hd_sketch(hv):
    hv = np.ascontiguousarray(hv, dtype=np.int16)
    packed = np.zeros(2 * hv.size, np.uint8)
    b = int(lib().hgo_compress_hd_sketch(hv, hv.size, packed))
    return b, packed[: b * hv.size // 8].copy()


def decompress_hd_sketch(packed, hv_d: int, b: int) -> np.ndarray:
    p = np.ascontiguousarray(packed, dtype=np.uint8)
    hv = np.empty(hv_d, np.int16)
    lib().hgo_decompress_hd_sketch(p, hv_d, b, hv)
    return hv


def dot_i16(r, q) -> int:
    r = np.ascontiguousarray(r, dtype=np.int16)
    q = np.ascontiguousarray(q, dtype=np.int16)
    return int(lib().hgo_dot_i16(r, q, r.size))


def ani_from_dot(dot: int, n_r: int, n_q: int, k: int = 21) -> np.float32:
    return np.float32(lib().hgo_ani_from_dot(dot, n_r, n_q, k))


def logf_glibc(x) -> np.float32:
    return np.float32(lib().hgo_logf_glibc(C.c_float(float(np.float32(x)))))


def logf_selfcheck(first: int, last: int, stride: int = 1) -> int:
    return int(lib().hgo_logf_selfcheck(first, last, stride))


def set_threads(n: int) -> int:
    return int(lib().hgo_set_threads(n))


def sketch_batch(seq, seg_off, k=21, scaled=1500, seed=123, canonical=True, hv_d=4096, want_hv=True):
    """sketch.rs:35-52 over a batch: returns dict(hv, packed, quant_bits, norm2, n_hashes)."""
    s = _bytes(seq)
    off = np.ascontiguousarray(seg_off, dtype=np.uint64)
    n = off.size - 1
    hv = np.empty((n, hv_d), np.int16) if want_hv else None
    packed = np.zeros((n, 2 * hv_d), np.uint8)
    qb = np.zeros(n, np.uint8)
    norm2 = np.zeros(n, np.int32)
    nh = np.zeros(n, np.uint32)
    lib().hgo_sketch_batch(_pad(s), off, n, k, scaled, seed, int(canonical), hv_d,
                           hv.ctypes.data if want_hv else None, packed.ctypes.data, qb.ctypes.data,
                           norm2.ctypes.data, nh.ctypes.data)
    return dict(hv=hv, packed=packed, quant_bits=qb, norm2=norm2, n_hashes=nh)


def kmer_count_batch(seq, seg_off, k=21, scaled=1500, seed=123, canonical=True) -> int:
    s = _bytes(seq)
    off = np.ascontiguousarray(seg_off, dtype=np.uint64)
    return int(lib().hgo_kmer_count_batch(_pad(s), off, off.size - 1, k, scaled, seed, int(canonical)))


def dist_all(ref_hv, ref_norm, qry_hv, qry_norm, k=21, symmetric=False, want_dot=True):
    """dist.rs:231-294: ANI (and i32 dot) of every enumerated pair, in the reference's order."""
    r = np.ascontiguousarray(ref_hv, dtype=np.int16)
    q = np.ascontiguousarray(qry_hv, dtype=np.int16)
    nr = np.ascontiguousarray(ref_norm, dtype=np.int32)
    nq = np.ascontiguousarray(qry_norm, dtype=np.int32)
    R, D = r.shape
    Q = q.shape[0]
    n_pairs = R * (Q - 1) // 2 if symmetric else R * Q
    ani = np.zeros(max(n_pairs, 1), np.float32)
    dot = np.zeros(max(n_pairs, 1), np.int32) if want_dot else None
    lib().hgo_dist_all(r, nr, R, q, nq, Q, D, k, int(symmetric), ani, dot.ctypes.data if want_dot else None)
    return ani[:n_pairs], (dot[:n_pairs] if want_dot else None)


def pair_indices(R: int, Q: int, symmetric: bool) -> np.ndarray:
    """dist.rs:251-265: the (i, j) of every pair index, row-major."""
    if symmetric:
        i, j = np.triu_indices(R, 1, Q)
    else:
        i, j = np.divmod(np.arange(R * Q), Q)
    return np.stack([i, j], 1).astype(np.int64)


def ani_output_order(ani, ani_th=85.0) -> np.ndarray:
    """utils.rs:260-286: pair indices in emission order (ANI desc, ties by desc pair index)."""
    a = np.ascontiguousarray(ani, dtype=np.float32)
    order = np.zeros(max(a.size, 1), np.uint64)
    n = int(lib().hgo_ani_output_order(_pad(a), a.size, ani_th, order))
    return order[:n].astype(np.int64)


def format_ani_tsv(ref_names, qry_names, pairs, ani, order) -> str:
    """utils.rs:274-281: `{ref}\\t{query}\\t{:.3}\\n` per emitted pair."""
    out = []
    for p in order:
        i, j = pairs[p]
        out.append("%s\t%s\t%.3f\n" % (ref_names[i], qry_names[j], float(ani[p])))
    return "".join(out)
