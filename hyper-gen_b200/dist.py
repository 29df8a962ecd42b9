"""Host mirror of dist::dist (reference src/dist.rs:11-63): load two sketch files, unpack the
HVs on the GPU, run the fused all-pairs kernel, order and write the survivors."""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import fileio
from .ffi import Context


@dataclass
class SketchDist:
    """types.rs:237-245"""
    path_ref_sketch: str = ""
    path_query_sketch: str = ""
    out_file: str = ""
    ksize: int = 21
    hv_d: int = 1024
    ani_threshold: float = 85.0


def pair_index(i, j, n_qry: int, symmetric: bool):
    """Index of (i, j) in the reference's enumeration (dist.rs:251-265)."""
    i = np.asarray(i, np.int64)
    j = np.asarray(j, np.int64)
    if symmetric:
        return i * (n_qry - 1) - i * (i - 1) // 2 + (j - i - 1)
    return i * n_qry + j


def reference_output_order(hits, n_ref: int, n_qry: int, symmetric: bool) -> np.ndarray:
    """Permutation of `hits` into dump_ani_file's order: stable ascending sort by ANI over the
    pair enumeration, reversed (utils.rs:262-269) => ANI descending, ties by descending pair index."""
    idx = pair_index(hits["i"], hits["j"], n_qry, symmetric)
    asc = np.lexsort((idx, hits["ani"]))
    return asc[::-1]


def _stack_packed(sketches):
    hv_d = sketches[0].hv_d
    n = len(sketches)
    packed = np.zeros((n, 2 * hv_d), np.uint8)
    bits = np.zeros(n, np.uint8)
    norm = np.zeros(n, np.int32)
    for t, s in enumerate(sketches):
        b = s.hv.view(np.uint8)
        packed[t, :b.size] = b
        bits[t] = s.hv_quant_bits
        norm[t] = s.hv_norm_2
    return packed, bits, norm


def dist(sd: SketchDist, ctx: Context | None = None, path: int = 0) -> str:
    """Returns the TSV text written to sd.out_file."""
    own = ctx is None
    ctx = ctx or Context(0)
    try:
        if_sym = sd.path_ref_sketch == sd.path_query_sketch  # dist.rs:13 (path equality)
        ref = fileio.load_sketch(sd.path_ref_sketch)
        qry = ref if if_sym else fileio.load_sketch(sd.path_query_sketch)
        if ref[0].ksize != qry[0].ksize:
            raise ValueError("Ref and query sketches use different kmer sizes!")  # dist.rs:29-32
        if ref[0].hv_d != qry[0].hv_d:
            raise ValueError("Ref and query sketches use different HV dimensions!")  # dist.rs:36-39
        hv_d, ksize = ref[0].hv_d, ref[0].ksize
        # packed rows as the sketch file holds them -> GPU: decompress (hd.rs:171-232), dist, and the output
        # stage's sort all on the device (hg_dist_packed); `milli` is the `{:.3}` field in thousandths
        rp, rb, rn = _stack_packed(ref)
        qp, qb, qn = (rp, rb, rn) if if_sym else _stack_packed(qry)
        hits, milli = ctx.dist_packed(rp, rb, rn, qp, qb, qn, hv_d, ksize=ksize, ani_th=sd.ani_threshold,
                                      symmetric=if_sym, path=path, sorted_output=True)
        text = fileio.format_ani_lines_milli([s.file_str for s in ref], [s.file_str for s in qry], hits, milli)
        if sd.out_file:
            with open(sd.out_file, "w") as f:
                f.write(text)
        return text
    finally:
        if own:
            ctx.close()
