"""Stress of the sharded dist (run under torchrun): two data sets alternate in runs of `run_len` steps, so a member that
reads a peer's rows before they arrived computes on the OTHER data set's rows and is caught; every step's gathered hit
list is compared, record by record, with the single-GPU result of the same data set (itself pinned to the oracle by
tests/test_gpu_dist.py).  A mismatch is reported with the tile, the member it was dealt to and the chunks it reads.
    torchrun ... tools/peer_stress.py [cfg3|cfg4|cfg5] [steps] [run_len]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("gloo")
import hypergen_b200 as hg
from hypergen_b200 import multigpu, synth
import bench as B
which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
run_len = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = {"cfg3": B.DIST_CONFIGS[0], "cfg4": B.DIST_CONFIGS[1], "cfg5": B.DIST_CONFIGS[2]}[which]
ctx = hg.Context(rank)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
D, sym = cfg["hv_d"], cfg["symmetric"]
cap = 4_000_000
sets = []
for seed in (cfg["seed"], cfg["seed"] + 77):
    if sym:
        n_ref = n_qry = cfg["n"]
        hv, norm, _, _ = B.encode_family(ctx, synth, dev, n_ref, D, cfg["n_per"], cfg["scaled"], seed)
        sets.append((hv, norm, hv, norm))
    else:
        n_ref, n_qry = cfg["n_ref"], cfg["n_qry"]
        rhv, rnorm, _, _ = B.encode_family(ctx, synth, dev, n_ref, D, cfg["n_per"], cfg["scaled"], seed)
        q_idx = torch.arange(5, n_ref, n_ref // n_qry, device=dev)[:n_qry]
        sets.append((rhv, rnorm, rhv[q_idx].contiguous(), rnorm[q_idx].contiguous()))


def key(h):
    return h["i"].astype(np.int64) << 32 | h["j"].astype(np.int64)


want = []
if rank == 0:  # the single-GPU answer of both data sets
    hp = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
    d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    for (rhv, rn, qhv, qn) in sets:
        with torch.cuda.stream(ext):
            ctx.dist_dev(rhv.data_ptr(), rn.data_ptr(), n_ref, 0, qhv.data_ptr(), qn.data_ptr(), n_qry, 0, D, 21, 85.0, sym, 0,
                         hp.data_ptr(), cap, d_cnt.data_ptr())
        ctx.sync()
        torch.cuda.synchronize()
        c = int(d_cnt.item())
        h = hp.numpy().view(hg.ffi.HIT_DTYPE)[:c].copy()
        want.append(h[np.argsort(key(h), kind="stable")])
    print("single-GPU: %s path %d, hits %s" % (which, ctx.dist_last_path, [w.size for w in want]), flush=True)
pg = multigpu.PeerGroup(ctx, hg.ffi.peer_window_need(n_qry, D, cap))
qb = multigpu.block_rows(n_qry, world); rb = multigpu.block_rows(n_ref, world)
a, b = qb[rank], qb[rank + 1]; ra, rbb = rb[rank], rb[rank + 1]
from multiprocessing import shared_memory
name = "hg_stress_hits_%s" % os.environ.get("MASTER_PORT", "0")
shm = shared_memory.SharedMemory(name=name, create=True, size=cap * 16) if rank == 0 else None
if world > 1: dist.barrier()
if rank != 0:
    shm = shared_memory.SharedMemory(name=name)
hits_np = np.ndarray((cap,), dtype=hg.ffi.HIT_DTYPE, buffer=shm.buf)
mapped = hg.ffi.host_register(hits_np)


def step(s, path):
    rhv, rn, qhv, qn = sets[s]
    if sym:
        pg.peer.dist_sharded_dev(None, None, 0, 0, qhv[a:b].data_ptr(), qn[a:b].data_ptr(), qb, D, 21, 85.0, True, path, 0, cap, mapped)
    else:
        pg.peer.dist_sharded_dev(rhv[ra:rbb].data_ptr(), rn[ra:rbb].data_ptr(), rbb - ra, ra, qhv[a:b].data_ptr(), qn[a:b].data_ptr(),
                                 qb, D, 21, 85.0, False, path, 0, cap, mapped)
    return pg.peer.dist_sharded_hits(cap, hits=hits_np)[0]


with torch.cuda.stream(ext):
    step(0, 0)
path = ctx.dist_last_path
tr, tc = (256, 256) if path == 3 else (256, 128)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bad = 0
for it in range(steps):
    s = (it // run_len) & 1
    flush.fill_(1); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    with torch.cuda.stream(ext):
        h = step(s, path)
    if rank == 0:
        g = h.copy()
        g = g[np.argsort(key(g), kind="stable")]
        w = want[s]
        if g.size == w.size and np.array_equal(g, w):
            continue
        bad += 1
        kg, kw = key(g), key(w)
        miss, extra = w[~np.isin(kw, kg)], g[~np.isin(kg, kw)]
        dup = g.size - np.unique(kg).size
        common_g, common_w = g[np.isin(kg, kw)], w[np.isin(kw, kg)]
        wrong = 0
        if common_g.size == common_w.size:
            wrong = int(np.count_nonzero((common_g["dot"] != common_w["dot"]) | (common_g["ani"].view(np.uint32) != common_w["ani"].view(np.uint32))))
        print("step %d (set %d, pos %d in run): %d hits vs %d; missing %d extra %d duplicate keys %d wrong values %d" % (
            it, s, it % run_len, g.size, w.size, miss.size, extra.size, dup, wrong), flush=True)
        for nm, arr in (("missing", miss), ("extra", extra)):
            if arr.size:
                tiles = {}
                for r in arr[:2000]:
                    tiles.setdefault((int(r["i"]) // tr, int(r["j"]) // tc), 0)
                    tiles[(int(r["i"]) // tr, int(r["j"]) // tc)] += 1
                print("   %s: first %s; tiles (R, C): count %s" % (nm, arr[:3], dict(list(tiles.items())[:12])), flush=True)
if rank == 0:
    print("peer_stress %s world %d path %d: %d steps, %d bad" % (which, world, path, steps, bad), flush=True)
if world > 1: dist.barrier()
hg.ffi.host_unregister(hits_np); del hits_np, h; shm.close()
if rank == 0: shm.unlink()
pg.close(); ctx.close()
if world > 1: dist.destroy_process_group()
