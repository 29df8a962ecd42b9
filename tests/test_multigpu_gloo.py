"""The N > 1 path on CPU: world_size 2 over gloo.  The collective logic of
hypergen_b200.multigpu (query broadcast, row shards, hit gather) is device independent; the
per-shard compute is injected, here the oracle (test infrastructure), on the GPU the CUDA call."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, symmetric, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hypergen_b200 as hg
    from hypergen_b200 import multigpu as mg
    import oracle as O

    rng = np.random.default_rng(11)  # same data on both ranks; only rank 0's copy is used as the source
    n, D = 37, 256
    sets = [np.unique(rng.integers(0, 1 << 50, 120, dtype=np.uint64)) for _ in range(n)]
    for t in range(1, n, 3):
        sets[t] = np.unique(np.concatenate([sets[t - 1][:100], sets[t][:15]]))
    hv = np.stack([O.encode_hd(s, D) for s in sets])
    norm = np.array([O.hv_l2_norm_sq(v) for v in hv], np.int32)
    n_ref = n if symmetric else 23

    def compute(ref, ref_norm, i0, qry, qry_norm):
        r, q = ref.numpy(), qry.numpy()
        ani, dot = O.dist_all(r, ref_norm.numpy(), q, qry_norm.numpy(), symmetric=False)
        R, Q = r.shape[0], q.shape[0]
        ii, jj = np.divmod(np.arange(R * Q), Q)
        gi = ii + i0
        keep = (ani >= np.float32(85.0)) & ((jj > gi) if symmetric else True)
        h = np.zeros(int(keep.sum()), hg.ffi.HIT_DTYPE)
        h["i"], h["j"], h["dot"], h["ani"] = gi[keep], jj[keep], dot[keep], ani[keep]
        return h

    q_t = torch.from_numpy(hv) if rank == 0 else None
    qn_t = torch.from_numpy(norm) if rank == 0 else None
    if symmetric:
        hits = mg.dist_sharded(compute, None, None, q_t, qn_t, n, n, D, True, "cpu")
    else:
        # every rank holds the ref matrix (a database shard would be loaded locally)
        hits = mg.dist_sharded(compute, torch.from_numpy(hv[:n_ref]), torch.from_numpy(norm[:n_ref]), q_t, qn_t, n_ref,
                               n, D, False, "cpu")
    if rank == 0:
        ani, dot = O.dist_all(hv[:n_ref], norm[:n_ref], hv, norm, symmetric=symmetric)
        pairs = O.pair_indices(n_ref, n, symmetric)
        want = np.nonzero(ani >= np.float32(85.0))[0]
        got = mg.__dict__  # noqa: F841
        from hypergen_b200 import dist as hdist
        idx = hdist.pair_index(hits["i"], hits["j"], n, symmetric)
        ok = np.array_equal(np.sort(idx), want) and np.array_equal(hits["dot"], dot[idx]) and \
            np.array_equal(hits["ani"].view(np.uint32), ani[idx].view(np.uint32)) and len(want) > 0
        open(os.path.join(out_dir, "ok_%d" % int(symmetric)), "w").write("1" if ok else "0")
    else:
        assert hits is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("symmetric", [True, False])
def test_dist_sharded_world2_gloo(symmetric, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, symmetric, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / ("ok_%d" % int(symmetric))).read() == "1"


def _sketch_worker(rank, world, port, d, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hypergen_b200 as hg
    from hypergen_b200 import multigpu as mg, fileio
    import oracle as O

    def sketch_fn(paths):  # the GPU call on a real run; the oracle here
        out = []
        for f in paths:
            seq = fileio.read_merge_seq(f)
            w = O.sketch_batch(seq, np.array([0, seq.size], np.uint64), scaled=200, hv_d=512, want_hv=False)
            b = int(w["quant_bits"][0])
            out.append(fileio.FileSketch(21, 200, True, 123, 512, b, int(w["norm2"][0]), f,
                                         w["packed"][0, :b * 512 // 8].copy().view("<i2")))
        return out

    files = fileio.get_fasta_files(d)
    res = mg.sketch_files_distributed(files, sketch_fn, os.path.join(out_dir, "db.sketch") if rank == 0 else None)
    if rank == 0:
        want = sketch_fn(files)
        back = fileio.load_sketch(os.path.join(out_dir, "db.sketch"))
        ok = len(back) == len(files) and all(a.file_str == b.file_str and a.hv_norm_2 == b.hv_norm_2 and
                                             np.array_equal(a.hv, b.hv) for a, b in zip(back, want))
        open(os.path.join(out_dir, "ok_sketch"), "w").write("1" if ok else "0")
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_sketch_files_distributed_world2_gloo(tmp_path):
    from hypergen_b200 import synth
    d = tmp_path / "g"
    d.mkdir()
    for g in range(7):
        seq = synth.family_member(g, 20_000 + 7_000 * g).numpy()
        with open(d / ("f%d.fna" % g), "wb") as f:
            f.write(b">g%d\n" % g + bytes(seq) + b"\n")
    port = _free_port()
    mp.spawn(_sketch_worker, args=(2, port, str(d), str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok_sketch").read() == "1"


def _fixed_gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hypergen_b200 as hg
    from hypergen_b200 import multigpu as mg
    ok = True
    for n_local, cap in ((5 + 3 * rank, 16), (0, 4), (9 + rank, 8)):
        h = np.zeros(max(n_local, cap), hg.ffi.HIT_DTYPE)
        h["i"][:n_local] = 1000 * rank + np.arange(n_local)
        h["j"][:n_local] = 7
        raw = torch.from_numpy(np.frombuffer(h.tobytes(), dtype=np.uint8).copy())
        got, overflow = mg.gather_hits_fixed(raw, torch.tensor([n_local], dtype=torch.int64), cap)
        # the same through a hit_block (counter + hits in the layout the gather sends: no staging copy)
        blk, cnt_v, hits_v = mg.hit_block(cap, "cpu")
        cnt_v[0] = n_local
        hits_v[: min(n_local, cap) * 16] = raw[: min(n_local, cap) * 16]
        got2, overflow2 = mg.gather_hits_fixed(None, None, cap, block=blk, recv=torch.empty(world * (16 + cap * 16), dtype=torch.uint8))
        ok = ok and overflow2 == overflow and ((got is None) == (got2 is None)) and (got is None or np.array_equal(got, got2))
        want_overflow = any((9 + r if cap == 8 else (5 + 3 * r if cap == 16 else 0)) > cap for r in range(world))
        ok = ok and (overflow == want_overflow)
        if rank == 0 and not overflow:
            exp = np.concatenate([1000 * r + np.arange((5 + 3 * r) if cap == 16 else 0) for r in range(world)])
            ok = ok and np.array_equal(np.sort(got["i"]), np.sort(exp.astype(np.uint32)))
        if overflow:
            full = mg.gather_hits(h[:n_local], "cpu")
            if rank == 0:
                ok = ok and full.size == sum(9 + r for r in range(world))
    if rank == 0:
        open(os.path.join(out_dir, "ok_fixed"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_gather_hits_fixed_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_fixed_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok_fixed").read() == "1"
