"""Host glue around the hot path: the reference's file formats, byte for byte.

  * FASTA discovery / reading  — utils::get_fasta_files (src/utils.rs:208-221) and
    fastx_reader::read_merge_seq (src/fastx_reader.rs:6-29);
  * sketch file                — bincode 1.x default encoding of Vec<FileSketch>
    (src/types.rs:224-235, src/utils.rs:234-258): LE fixed-width ints, u64 length prefixes;
  * ANI output                 — utils::dump_ani_file (src/utils.rs:260-308).
"""
from __future__ import annotations

import glob
import os
import struct
from dataclasses import dataclass, field

import numpy as np


@dataclass
class FileSketch:
    """types.rs:224-235 — one on-disk record."""
    ksize: int = 21
    scaled: int = 1500
    canonical: bool = True
    seed: int = 123
    hv_d: int = 4096
    hv_quant_bits: int = 16
    hv_norm_2: int = 0
    file_str: str = ""
    hv: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int16))  # packed bytes viewed as i16


def get_fasta_files(path: str) -> list[str]:
    """*.fna, then *.fa, then *.fasta under `path`, each group in glob (sorted) order; like the glob crate's default
    MatchOptions (utils.rs get_fasta_files) `*` also matches a leading dot."""
    out = []
    for pat in ("*.fna", "*.fa", "*.fasta"):
        out += sorted(glob.glob(os.path.join(path, pat), include_hidden=True))
    return out


def read_merge_seq(file_name: str) -> np.ndarray:
    """Sequence lines concatenated, one 'N' per header line, trailing \\n / \\r stripped."""
    with open(file_name, "rb") as f:
        data = f.read()
    return merge_seq_bytes(data)


def merge_seq_bytes(data: bytes) -> np.ndarray:
    parts = []
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            parts.append(b"N")
        else:
            parts.append(line[:-1] if line.endswith(b"\r") else line)
    # a trailing empty piece after the final '\n' contributes nothing
    return np.frombuffer(b"".join(parts), dtype=np.uint8)


def dump_sketch(sketches: list[FileSketch], out_path: str) -> int:
    """bincode::serialize::<Vec<FileSketch>> + fs::write; returns the byte count."""
    buf = bytearray(struct.pack("<Q", len(sketches)))
    for s in sketches:
        name = s.file_str.encode("utf-8")
        hv = np.ascontiguousarray(s.hv, dtype="<i2")
        buf += struct.pack("<BQBQQBi", s.ksize, s.scaled, 1 if s.canonical else 0, s.seed, s.hv_d, s.hv_quant_bits,
                           int(s.hv_norm_2))
        buf += struct.pack("<Q", len(name)) + name
        buf += struct.pack("<Q", hv.size) + hv.tobytes()
    with open(out_path, "wb") as f:
        f.write(buf)
    return len(buf)


def load_sketch(path: str) -> list[FileSketch]:
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    (n,) = struct.unpack_from("<Q", data, pos)
    pos += 8
    out = []
    hdr = struct.Struct("<BQBQQBi")
    for _ in range(n):
        ksize, scaled, canonical, seed, hv_d, qb, norm = hdr.unpack_from(data, pos)
        pos += hdr.size
        (ln,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        name = data[pos:pos + ln].decode("utf-8")
        pos += ln
        (hn,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        hv = np.frombuffer(data, dtype="<i2", count=hn, offset=pos).copy()
        pos += 2 * hn
        out.append(FileSketch(ksize, scaled, bool(canonical), seed, hv_d, qb, norm, name, hv))
    if pos != len(data):
        raise ValueError("trailing bytes in sketch file %s" % path)
    return out


def format_ani_lines(ref_names, qry_names, hits, order) -> str:
    """`{ref}\\t{query}\\t{:.3}\\n` per reported pair, in `order` (utils.rs:274-281)."""
    return "".join("%s\t%s\t%.3f\n" % (ref_names[int(hits["i"][t])], qry_names[int(hits["j"][t])],
                                       float(hits["ani"][t])) for t in order)


def format_ani_lines_milli(ref_names, qry_names, hits, milli) -> str:
    """The same lines from hits already in output order and their ANI in thousandths (hg_dist_sorted)."""
    return "".join("%s\t%s\t%d.%03d\n" % (ref_names[int(i)], qry_names[int(j)], m // 1000, m % 1000)
                   for i, j, m in zip(hits["i"].tolist(), hits["j"].tolist(), milli.tolist()))
