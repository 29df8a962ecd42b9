"""Host glue mirrored from the reference (no GPU): FASTA merge, sketch-file framing, output
order/format, multi-GPU partitioning.  Reference: src/fastx_reader.rs:6-29, src/utils.rs:208-308,
src/types.rs:224-235."""
import os
import struct

import numpy as np
import pytest


def test_read_merge_seq_matches_oracle(oracle, hg, tmp_path):
    from hypergen_b200 import fileio
    cases = [b">a\nACGT\nAC\n>b\r\nGG\r\nTT", b"", b"ACGT", b">x", b">x\n", b"\n\nAC\n", b">r1 desc\nacgtNN\n>r2\nTTTT\n",
             b">test\nAGCTCTTANNAGCCCNTTacgttacagccctgaaaacttt"]
    for c in cases:
        assert bytes(fileio.merge_seq_bytes(c)) == bytes(oracle.read_merge_seq(c)), c
    p = tmp_path / "g.fna"
    p.write_bytes(cases[0])
    assert bytes(fileio.read_merge_seq(str(p))) == b"NACGTACNGGTT"


def test_get_fasta_files_order(hg, tmp_path):
    from hypergen_b200 import fileio
    for n in ("b.fna", "a.fna", "z.fa", "c.fasta", "skip.txt", ".hidden.fna"):
        (tmp_path / n).write_bytes(b">x\nACGT\n")
    got = [os.path.basename(f) for f in fileio.get_fasta_files(str(tmp_path))]
    # *.fna, then *.fa, then *.fasta; the glob crate's default options let * match a leading dot (utils.rs get_fasta_files)
    assert got == [".hidden.fna", "a.fna", "b.fna", "z.fa", "c.fasta"]


def test_sketch_file_is_bincode_of_vec_filesketch(hg, oracle, tmp_path):
    from hypergen_b200 import fileio
    hv = (np.arange(4096) % 37 - 18).astype(np.int16)
    b, packed = oracle.compress_hd_sketch(hv)
    sk = fileio.FileSketch(21, 1500, True, 123, 4096, b, oracle.hv_l2_norm_sq(hv), "dir/g1.fna", packed.view("<i2"))
    path = str(tmp_path / "s.sketch")
    n = fileio.dump_sketch([sk, sk], path)
    raw = open(path, "rb").read()
    assert n == len(raw)
    # bincode 1.x default: u64 count, then fields in declaration order, fixed-width LE
    assert struct.unpack_from("<Q", raw, 0)[0] == 2
    ksize, scaled, canonical, seed, hv_d, qb, norm = struct.unpack_from("<BQBQQBi", raw, 8)
    assert (ksize, scaled, canonical, seed, hv_d, qb, norm) == (21, 1500, 1, 123, 4096, b, oracle.hv_l2_norm_sq(hv))
    off = 8 + 31
    assert struct.unpack_from("<Q", raw, off)[0] == len("dir/g1.fna") and raw[off + 8:off + 18] == b"dir/g1.fna"
    assert struct.unpack_from("<Q", raw, off + 18)[0] == b * 4096 // 16  # Vec<i16> length (hd.rs:155-157)
    back = fileio.load_sketch(path)
    assert len(back) == 2 and back[1].file_str == "dir/g1.fna" and back[0].hv_quant_bits == b
    assert np.array_equal(oracle.decompress_hd_sketch(back[0].hv.view(np.uint8), 4096, b), hv)


def test_reference_output_order_and_tsv(hg, oracle):
    from hypergen_b200 import dist as hdist, fileio
    rng = np.random.default_rng(2)
    R = Q = 9
    pairs = oracle.pair_indices(R, Q, True)
    ani = rng.choice(np.array([84.9, 85.0, 90.5, 99.123, 100.0], np.float32), size=len(pairs))
    keep = np.nonzero(ani >= np.float32(85.0))[0]
    perm = rng.permutation(keep)  # the device appends in arbitrary order
    hits = np.zeros(len(perm), hg.ffi.HIT_DTYPE)
    hits["i"], hits["j"], hits["ani"] = pairs[perm, 0], pairs[perm, 1], ani[perm]
    order = hdist.reference_output_order(hits, R, Q, True)
    want = oracle.ani_output_order(ani, 85.0)
    assert np.array_equal(hdist.pair_index(hits["i"], hits["j"], Q, True)[order], want)
    names = ["g%d.fna" % i for i in range(R)]
    assert fileio.format_ani_lines(names, names, hits, order) == oracle.format_ani_tsv(names, names, pairs, ani, want)


def test_partitioning(hg):
    from hypergen_b200 import multigpu as mg
    sizes = [5, 9, 1, 7, 3, 8, 2, 6]
    parts = mg.partition_greedy(sizes, 3)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes)
    for n, w in ((10000, 8), (1000, 3), (7, 4), (129, 2)):
        b = mg.triangle_rows(n, w)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        cnt = [mg.pairs_in_rows(n, b[r], b[r + 1]) for r in range(w)]
        assert sum(cnt) == n * (n - 1) // 2
        if n >= 1000:
            assert max(cnt) / (sum(cnt) / w) < 1.05  # balanced triangle
        assert mg.even_rows(n, w)[-1] == n
