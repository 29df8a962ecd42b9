"""Host mirror of sketch_cuda::sketch_cuda (reference src/sketch_cuda.rs:43-117): discover
FASTA files, read them, sketch the whole batch through one C-ABI call, write the sketch file."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import fileio
from .ffi import Context, make_params


@dataclass
class SketchParams:
    """types.rs:83-113 (defaults included)"""
    path: str = ""
    out_file: str = ""
    canonical: bool = True
    ksize: int = 21
    seed: int = 123
    scaled: int = 1500
    hv_d: int = 4096


def sketch_sequences(ctx: Context, seqs: list[np.ndarray], names: list[str], p: SketchParams,
                     batch_bytes: int = 1 << 30) -> list[fileio.FileSketch]:
    """Sketch in-memory sequences in batches of at most `batch_bytes` of sequence."""
    out: list[fileio.FileSketch] = []
    params = make_params(k=p.ksize, scaled=p.scaled, seed=p.seed, canonical=p.canonical, hv_d=p.hv_d)
    start = 0
    while start < len(seqs):
        end, tot = start, 0
        while end < len(seqs) and (end == start or tot + seqs[end].size <= batch_bytes):
            tot += seqs[end].size
            end += 1
        sizes = np.array([0] + [s.size for s in seqs[start:end]], np.uint64)
        seg_off = np.cumsum(sizes).astype(np.uint64)
        seq = np.concatenate(seqs[start:end]) if tot else np.zeros(0, np.uint8)
        r = ctx.sketch_batch(seq, seg_off, params, want_hv=False)
        for t in range(end - start):
            b = int(r["quant_bits"][t])
            nbytes = b * p.hv_d // 8
            out.append(fileio.FileSketch(p.ksize, p.scaled, p.canonical, p.seed, p.hv_d, b, int(r["norm2"][t]),
                                         names[start + t], r["packed"][t, :nbytes].copy().view("<i2")))
        start = end
    return out


def sketch(p: SketchParams, ctx: Context | None = None) -> list[fileio.FileSketch]:
    own = ctx is None
    ctx = ctx or Context(0)
    try:
        files = fileio.get_fasta_files(p.path)
        seqs = [fileio.read_merge_seq(f) for f in files]
        sk = sketch_sequences(ctx, seqs, files, p)
        if p.out_file:
            fileio.dump_sketch(sk, p.out_file)
        return sk
    finally:
        if own:
            ctx.close()
