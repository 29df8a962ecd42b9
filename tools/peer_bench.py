"""Stage timeline of the sharded dist on every rank (run under torchrun).
    torchrun ... tools/peer_bench.py [cfg3|cfg4|cfg5] [mapped|window]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("gloo")
import hypergen_b200 as hg
from hypergen_b200 import multigpu, synth
import bench as B
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
mode = sys.argv[2] if len(sys.argv) > 2 else "mapped"
cfg = {"cfg3": B.DIST_CONFIGS[0], "cfg4": B.DIST_CONFIGS[1], "cfg5": B.DIST_CONFIGS[2]}[which]
prof = os.environ.get("PB_NOPROF") is None
ctx = hg.Context(rank); ctx.set_profiling(prof)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
D, sym = cfg["hv_d"], cfg["symmetric"]
if sym:
    n_ref = n_qry = cfg["n"]
    hv, norm, bits, _ = B.encode_family(ctx, synth, dev, n_ref, D, cfg["n_per"], cfg["scaled"], cfg["seed"])
    ref_hv, ref_norm, qry_hv, qry_norm = hv, norm, hv, norm
else:
    n_ref, n_qry = cfg["n_ref"], cfg["n_qry"]
    ref_hv, ref_norm, bits, _ = B.encode_family(ctx, synth, dev, n_ref, D, cfg["n_per"], cfg["scaled"], cfg["seed"])
    q_idx = torch.arange(5, n_ref, n_ref // n_qry, device=dev)[:n_qry]
    qry_hv, qry_norm = ref_hv[q_idx].contiguous(), ref_norm[q_idx].contiguous()
cap = 4_000_000
pg = multigpu.PeerGroup(ctx, hg.ffi.peer_window_need(n_qry, D, cap))
qb = multigpu.block_rows(n_qry, world); rb = multigpu.block_rows(n_ref, world)
a, b = qb[rank], qb[rank + 1]; ra, rbb = rb[rank], rb[rank + 1]
mapped = None
hits_np = np.empty(cap, hg.ffi.HIT_DTYPE)
if mode == "registered":  # plain process memory, page-locked by cudaHostRegister
    hits_np = np.zeros(cap, hg.ffi.HIT_DTYPE)
    mapped = hg.ffi.host_register(hits_np)
elif mode in ("mapped", "shm"):
    if world > 1 or mode == "shm":
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(name="hg_pb_hits", create=True, size=cap * 16) if rank == 0 else None
        if world > 1: dist.barrier()
        if rank != 0:
            shm = shared_memory.SharedMemory(name="hg_pb_hits")
        hits_np = np.ndarray((cap,), dtype=hg.ffi.HIT_DTYPE, buffer=shm.buf)
    else:
        hits_t = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True); hits_np = hits_t.numpy().view(hg.ffi.HIT_DTYPE)
    mapped = hg.ffi.host_register(hits_np) if (world > 1 or mode == "shm") else hits_t.data_ptr()
def step(path):
    if sym:
        pg.peer.dist_sharded_dev(None, None, 0, 0, qry_hv[a:b].data_ptr(), qry_norm[a:b].data_ptr(), qb, D, 21, 85.0, True, path, 0, cap, mapped)
    else:
        pg.peer.dist_sharded_dev(ref_hv[ra:rbb].data_ptr(), ref_norm[ra:rbb].data_ptr(), rbb - ra, ra, qry_hv[a:b].data_ptr(),
                                 qry_norm[a:b].data_ptr(), qb, D, 21, 85.0, False, path, 0, cap, mapped)
    return pg.peer.dist_sharded_hits(cap, hits=hits_np)[0]
with torch.cuda.stream(ext):
    step(0)
path = ctx.dist_last_path
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st, tot = [], []
for it in range(12):
    flush.fill_(1); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(ext):
        h = step(path)
    tot.append((time.perf_counter() - t0) * 1e3)
    st.append(pg.peer.stage_ms() if prof else [0, 0, 0, 0])
st = np.array(st[4:]).mean(0); tot = tot[2:]
line = "rank %d/%d %s %s path %d: host wall %.3f ms | operands %.3f push %.3f kernel %.3f final_barrier %.3f | hits %d" % (
    rank, world, which, mode, path, np.mean(tot[2:]), st[0], st[1], st[2], st[3], h.size)
if world > 1:
    lines = [None] * world
    dist.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines), flush=True)
    dist.barrier()
else:
    print(line, flush=True)
if mode == "registered":
    hg.ffi.host_unregister(hits_np)
if (mode == "mapped" and world > 1) or mode == "shm":
    hg.ffi.host_unregister(hits_np); del hits_np, h; shm.close()
    if rank == 0: shm.unlink()
pg.close(); ctx.close()
if world > 1: dist.destroy_process_group()
