"""End to end through the host mirrors of the reference drivers (sketch_cuda::sketch_cuda,
dist::dist): FASTA files -> sketch file -> ANI TSV, compared with the oracle's output text."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_fasta(path, seq, name, width=80, records=1):
    with open(path, "wb") as f:
        per = len(seq) // records
        for r in range(records):
            part = seq[r * per:(r + 1) * per if r < records - 1 else len(seq)]
            f.write(b">%s_%d some description\n" % (name.encode(), r))
            for i in range(0, len(part), width):
                f.write(bytes(part[i:i + width]) + b"\n")


def test_sketch_then_dist_files(ctx, hg, oracle, tmp_path):
    from hypergen_b200 import synth, sketch, dist, fileio
    d = tmp_path / "genomes"
    d.mkdir()
    n = 14
    for g in range(n):
        seq = synth.family_member(g + 6, 150_000).numpy()
        ext = ".fna" if g % 3 else ".fa"
        _write_fasta(str(d / ("g%02d%s" % (g, ext))), seq, "g%d" % g, records=1 + g % 3)
    out = str(tmp_path / "db.sketch")
    p = sketch.SketchParams(path=str(d), out_file=out, scaled=400, hv_d=2048)
    sk = sketch.sketch(p, ctx=ctx)
    files = fileio.get_fasta_files(str(d))
    assert [s.file_str for s in sk] == files and len(sk) == n
    # oracle: same reader semantics (one N per header line), same parameters
    seqs = [oracle.read_merge_seq(open(f, "rb").read()) for f in files]
    off = np.cumsum([0] + [s.size for s in seqs]).astype(np.uint64)
    want = oracle.sketch_batch(np.concatenate(seqs), off, scaled=400, hv_d=2048)
    back = fileio.load_sketch(out)
    for t in range(n):
        b = int(want["quant_bits"][t])
        assert back[t].hv_quant_bits == b and back[t].hv_norm_2 == int(want["norm2"][t])
        assert np.array_equal(back[t].hv.view(np.uint8), want["packed"][t, :b * 2048 // 8])
    # dist self-vs-self (same path => symmetric) and ref-vs-query through a second file
    tsv = dist.dist(dist.SketchDist(out, out, str(tmp_path / "ani.tsv"), ani_threshold=85.0), ctx=ctx)
    ani, _ = oracle.dist_all(want["hv"], want["norm2"], want["hv"], want["norm2"], symmetric=True)
    pairs = oracle.pair_indices(n, n, True)
    assert tsv == oracle.format_ani_tsv(files, files, pairs, ani, oracle.ani_output_order(ani, 85.0))
    assert tsv == open(tmp_path / "ani.tsv").read() and tsv.count("\n") > 0
    out2 = str(tmp_path / "q.sketch")
    fileio.dump_sketch(back[3:9], out2)
    tsv2 = dist.dist(dist.SketchDist(out, out2, "", ani_threshold=0.0), ctx=ctx)
    ani2, _ = oracle.dist_all(want["hv"], want["norm2"], want["hv"][3:9], want["norm2"][3:9], symmetric=False)
    pairs2 = oracle.pair_indices(n, 6, False)
    assert tsv2 == oracle.format_ani_tsv(files, files[3:9], pairs2, ani2, oracle.ani_output_order(ani2, 0.0))


def test_encode_sets_hook(ctx, hg, oracle):
    rng = np.random.default_rng(4)
    sets = [np.unique(rng.integers(0, 1 << 53, n, dtype=np.uint64)) for n in (0, 1, 5, 333, 4097, 9000)]
    for hv_d in (256, 4096):
        got = ctx.encode_sets(sets, hv_d)
        for t, s in enumerate(sets):
            hv = oracle.encode_hd(s, hv_d)
            assert np.array_equal(got["hv"][t], hv)
            b, packed = oracle.compress_hd_sketch(hv)
            assert got["quant_bits"][t] == b and got["norm2"][t] == oracle.hv_l2_norm_sq(hv)
            assert np.array_equal(got["packed"][t, :packed.size], packed)


def test_cpp_host_cli_matches_python_host(ctx, hg, oracle, tmp_path):
    """The C++ host (hyper-gen sketch / dist) writes the same sketch file and TSV."""
    import subprocess
    from hypergen_b200 import synth, sketch, dist, _build
    exe = _build.build_host()
    d = tmp_path / "genomes"
    d.mkdir()
    for g in range(9):
        _write_fasta(str(d / ("s%02d.fna" % g)), synth.family_member(g + 20, 120_000).numpy(), "s%d" % g, records=1 + g % 2)
    out_py, out_cc = str(tmp_path / "py.sketch"), str(tmp_path / "cc.sketch")
    sketch.sketch(sketch.SketchParams(path=str(d), out_file=out_py, scaled=500, hv_d=1024), ctx=ctx)
    subprocess.check_call([exe, "sketch", "-p", str(d), "-o", out_cc, "-s", "500", "-d", "1024", "-t", "4"])
    assert open(out_py, "rb").read() == open(out_cc, "rb").read()          # raw FASTA parsed on the GPU
    out_hp = str(tmp_path / "hp.sketch")
    subprocess.check_call([exe, "sketch", "-p", str(d), "-o", out_hp, "-s", "500", "-d", "1024", "-t", "4"],
                          env=dict(os.environ, HG_HOST_PARSE="1"))
    assert open(out_py, "rb").read() == open(out_hp, "rb").read()          # host-side reader, same bytes
    out_ts = str(tmp_path / "ts.sketch")  # a trailing separator does not change the recorded file names (PathBuf::join)
    subprocess.check_call([exe, "sketch", "-p", str(d) + "/", "-o", out_ts, "-s", "500", "-d", "1024", "-t", "4"])
    assert open(out_py, "rb").read() == open(out_ts, "rb").read()
    tsv_py = dist.dist(dist.SketchDist(out_py, out_py, str(tmp_path / "py.tsv"), ani_threshold=80.0), ctx=ctx)
    subprocess.check_call([exe, "dist", "-r", out_cc, "-q", out_cc, "-o", str(tmp_path / "cc.tsv"), "-a", "80.0"])
    assert open(tmp_path / "cc.tsv").read() == tsv_py and tsv_py
    # ref != query path => full R x Q grid (dist.rs:13)
    out2 = str(tmp_path / "copy.sketch")
    open(out2, "wb").write(open(out_cc, "rb").read())
    tsv2 = dist.dist(dist.SketchDist(out_py, out2, "", ani_threshold=80.0), ctx=ctx)
    subprocess.check_call([exe, "dist", "-r", out_cc, "-q", out2, "-o", str(tmp_path / "cc2.tsv"), "-a", "80.0"])
    assert open(tmp_path / "cc2.tsv").read() == tsv2
