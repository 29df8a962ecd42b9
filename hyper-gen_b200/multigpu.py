"""One process per GPU: how the two stages shard.

  * sketch — genomes are independent (the reference already treats files as independent
    rayon tasks, src/sketch.rs:35): greedy longest-first partition of the files over the
    ranks, no collective on the data path.
  * dist   — the ref sketch matrix is row-sharded, the query HVs + norms are broadcast, every
    rank runs the fused kernel on its R_r x Q block and the (sparse) hit lists are gathered.
    For the symmetric all-vs-all case the row boundaries are chosen so every rank owns the
    same number of upper-triangle pairs.

Two transports exist for the dist exchange:
  * the product path, `PeerGroup`: NVLink windows inside the library (csrc/peer.cu) - every rank turns its
    own rows into the kernel's operand form and pushes them into the windows of the ranks that compute with
    them (block pairs owned along the ring), hits are collected per rank and moved to rank 0's list in one
    piece.  torch.distributed only carries the 64-byte window handles once, at start-up;
  * `dist_sharded`, collectives of torch.distributed around a `compute` callback (NCCL on GPUs, gloo on
    CPU tensors): the reference formulation of the same sharding, kept for the CPU tests of the host logic.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from .ffi import HIT_DTYPE



def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, so that pinned staging
    buffers are first-touched on that node and the H2D stream does not cross the socket link.
    Returns what was done (for the bench line); a no-op where sysfs does not say."""
    import os
    info = {"bound": False}
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open(path + "/numa_node").read())
        cpulist = open(path + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info.update(numa_node=node, cpus=len(cpus))
        if node >= 0 and cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:  # noqa: BLE001 - diagnostic only
        info["error"] = repr(e)[:80]
    return info

def partition_greedy(sizes, n_ranks: int) -> list[list[int]]:
    """Longest-first assignment of items (genome files by byte size) to the least-loaded rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * n_ranks
    out: list[list[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda t: (load[t], t))
        out[r].append(i)
        load[r] += int(sizes[i])
    for lst in out:
        lst.sort()
    return out


def even_rows(n: int, n_ranks: int) -> list[int]:
    """Boundaries a_0..a_N of a contiguous row split with sizes differing by at most one."""
    return [(n * r) // n_ranks for r in range(n_ranks + 1)]


def triangle_rows(n: int, n_ranks: int, align: int = 1) -> list[int]:
    """Row boundaries such that each shard [a_r, a_{r+1}) holds ~1/N of the pairs (i, j > i)."""
    total = n * (n - 1) // 2
    bounds = [0]
    for r in range(1, n_ranks):
        target = total * r / n_ranks
        # pairs in rows [0, a) = a*(n-1) - a*(a-1)/2  -> solve the quadratic for a
        a = (2 * n - 1 - math.sqrt(max((2 * n - 1) ** 2 - 8 * target, 0.0))) / 2
        a = int(round(a / align)) * align
        bounds.append(min(max(a, bounds[-1]), n))
    bounds.append(n)
    return bounds


def pairs_in_rows(n: int, a: int, b: int) -> int:
    f = lambda x: x * (n - 1) - x * (x - 1) // 2
    return f(b) - f(a)


def broadcast_queries_fused(buf: torch.Tensor, n: int, d: int, src: int = 0):
    """One broadcast for HVs + norms: `buf` is a uint8 tensor of n*d*2 + n*4 bytes (filled on `src`).
    Returns (hv int16 [n, d], norm int32 [n]) views into it."""
    dist.broadcast(buf, src=src)
    hv = buf[: n * d * 2].view(torch.int16).view(n, d)
    norm = buf[n * d * 2: n * d * 2 + n * 4].view(torch.int32)
    return hv, norm


def broadcast_queries(qry_hv: torch.Tensor | None, qry_norm: torch.Tensor | None, shape, device, src: int = 0):
    """Rank `src` owns the query HVs (n x D int16) and norms (n int32); everyone gets a copy."""
    rank = dist.get_rank()
    n, d = shape
    if rank != src:
        qry_hv = torch.empty((n, d), dtype=torch.int16, device=device)
        qry_norm = torch.empty((n,), dtype=torch.int32, device=device)
    # moved as raw bytes: NCCL does not care, and gloo has no int16 kernels
    dist.broadcast(qry_hv.view(torch.uint8), src=src)
    dist.broadcast(qry_norm.view(torch.uint8), src=src)
    return qry_hv, qry_norm


def gather_hits(local_hits, device, dst: int = 0, count: int | None = None) -> np.ndarray | None:
    """Variable-length hit lists -> one array on rank `dst` (rank order, unsorted — the output
    stage orders hits, dist.reference_output_order); None elsewhere.

    `local_hits` is either a numpy HIT_DTYPE array or a uint8 torch tensor already on `device`
    holding `count` packed hg_hit records (the CUDA path: nothing touches the host until rank
    `dst` reads the gathered buffer once)."""
    world = dist.get_world_size()
    isz = HIT_DTYPE.itemsize
    if isinstance(local_hits, np.ndarray):
        count = int(local_hits.size)
        raw = torch.from_numpy(np.frombuffer(local_hits.tobytes(), dtype=np.uint8).copy()).to(device)
    else:
        raw = local_hits[: count * isz]
    cnt = torch.tensor([count], dtype=torch.int64, device=device)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, cnt)
    counts = counts.tolist()  # one host read: sizes the second collective
    mx = max(max(counts), 1)
    buf = torch.zeros(mx * isz, dtype=torch.uint8, device=device)
    buf[: raw.numel()] = raw
    allbuf = torch.empty(world * mx * isz, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allbuf, buf)
    if dist.get_rank() != dst:
        return None
    host = allbuf.cpu().numpy()
    parts = [np.frombuffer(host[r * mx * isz:(r * mx + counts[r]) * isz].tobytes(), dtype=HIT_DTYPE) for r in range(world)]
    return np.concatenate(parts) if parts else np.zeros(0, HIT_DTYPE)


def hit_block(cap_per_rank: int, device) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One device buffer laid out as the gather sends it: [count:int64 | pad | cap_per_rank x hg_hit].
    Returns (block, count view (1 x int64), hits view (uint8)): point the dist kernel's counter and hit array at the
    views and gather_hits_fixed(block=...) sends the block as it is - no staging copy."""
    block = torch.zeros(16 + cap_per_rank * HIT_DTYPE.itemsize, dtype=torch.uint8, device=device)
    return block, block[:8].view(torch.int64), block[16:]


def gather_hits_fixed(local_hits: torch.Tensor | None, count_t: torch.Tensor | None, cap_per_rank: int, dst: int = 0,
                      host_buf: torch.Tensor | None = None, block: torch.Tensor | None = None,
                      recv: torch.Tensor | None = None):
    """One-collective gather for the common (sparse) case: every rank contributes a fixed-size
    record block [count:int64 | cap_per_rank x hg_hit] so that no host round trip is needed to size
    the collective.  Returns (hits or None, overflowed: bool); on overflow (some rank had more than
    cap_per_rank hits) the caller falls back to gather_hits().  `local_hits` is a uint8 device tensor,
    `count_t` a 1-element int64 device tensor (the kernel's hit counter) - or pass `block` from hit_block(),
    which already is the record block; `recv` an optional reusable device buffer of world * block bytes;
    `host_buf` an optional pinned uint8 tensor of at least world * cap_per_rank * 16 bytes owned by the caller."""
    world, rank = dist.get_world_size(), dist.get_rank()
    isz = HIT_DTYPE.itemsize
    nbytes = 16 + cap_per_rank * isz
    if block is None:
        block = torch.empty(nbytes, dtype=torch.uint8, device=local_hits.device)
        block[:8] = count_t.view(torch.uint8)
        block[16:] = local_hits[: cap_per_rank * isz]
    send = block[:nbytes]
    if recv is None or recv.numel() < world * nbytes:
        recv = torch.empty(world * nbytes, dtype=torch.uint8, device=send.device)
    recv = recv[: world * nbytes]
    dist.all_gather_into_tensor(recv, send)
    counts = recv.view(world, nbytes)[:, :8].contiguous().view(torch.int64).cpu().numpy().ravel()  # 8 B per rank
    overflow = bool((counts > cap_per_rank).any())
    if overflow or rank != dst:
        return None, overflow
    total = int(counts.sum())
    host_t = host_buf[: total * isz] if host_buf is not None and host_buf.numel() >= total * isz else \
        torch.empty(total * isz, dtype=torch.uint8)
    pos = 0
    for r in range(world):  # only the used part of every rank's block crosses PCIe
        nb = int(counts[r]) * isz
        if nb:
            host_t[pos:pos + nb].copy_(recv[r * nbytes + 16: r * nbytes + 16 + nb], non_blocking=True)
            pos += nb
    if recv.is_cuda:
        torch.cuda.current_stream().synchronize()
    return host_t.numpy()[: total * isz].view(HIT_DTYPE), False


def dist_sharded(compute, ref_hv, ref_norm, qry_hv, qry_norm, n_ref: int, n_qry: int, hv_d: int, symmetric: bool,
                 device, bounds=None) -> np.ndarray | None:
    """Row-shard the refs, broadcast the queries, compute locally, gather the hits on rank 0.

    ref_hv / ref_norm: this rank's view of the FULL ref matrix or None (then the shard is taken
    from the broadcast queries — the all-vs-all case where ref == query).
    compute(ref_block, ref_norm_block, i0, qry_hv, qry_norm) -> np.ndarray[HIT_DTYPE]
    """
    rank, world = dist.get_rank(), dist.get_world_size()
    qry_hv, qry_norm = broadcast_queries(qry_hv, qry_norm, (n_qry, hv_d), device)
    if bounds is None:
        bounds = triangle_rows(n_ref, world) if symmetric else even_rows(n_ref, world)
    a, b = bounds[rank], bounds[rank + 1]
    if ref_hv is None:
        ref_hv, ref_norm = qry_hv, qry_norm
    local = compute(ref_hv[a:b], ref_norm[a:b], a, qry_hv, qry_norm) if b > a else np.zeros(0, HIT_DTYPE)
    return gather_hits(local, device)


# ---------------------------------------------------------------------------------------------
# file-level drivers (the multi-GPU form of sketch_cuda::sketch_cuda and dist::dist)
# ---------------------------------------------------------------------------------------------
def gather_objects(obj, dst: int = 0):
    """Small python objects (FileSketch records) to rank `dst`."""
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def sketch_files_distributed(files, sketch_fn, out_file: str | None = None):
    """Every rank sketches its greedy (by file size) share of `files` with `sketch_fn(list_of_paths)
    -> list[FileSketch]`; rank 0 reassembles the records in get_fasta_files order and writes the
    sketch file.  No collective touches sequence data."""
    import os
    from . import fileio
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [os.path.getsize(f) for f in files]
    mine = partition_greedy(sizes, world)[rank]
    local = sketch_fn([files[i] for i in mine])
    parts = gather_objects((mine, local))
    if rank != 0:
        return None
    allsk = [None] * len(files)
    for idx, sks in parts:
        for i, s in zip(idx, sks):
            allsk[i] = s
    if out_file:
        fileio.dump_sketch(allsk, out_file)
    return allsk


# ---------------------------------------------------------------------------------------------
# NVLink-window transport (the product path): hg_peer behind torch.distributed's bootstrap
# ---------------------------------------------------------------------------------------------
def block_rows(n: int, n_ranks: int, align: int | None = None) -> list[int]:
    """Boundaries of the contiguous row blocks the ranks hold (as hg_group_dist_packed deals them): multiples of 256
    rows when there are at least 512 rows per rank - what the ring ownership of the all-vs-all needs (csrc/peer.cu) -
    else of 4 rows."""
    if align is None:
        if n >= 512 * n_ranks:
            t = (n + 255) // 256  # the 256-row tile rows, dealt evenly
            return [n if r == n_ranks else min(n, 256 * ((r * t + n_ranks // 2) // n_ranks)) for r in range(n_ranks + 1)]
        align = 4
    return [n if r == n_ranks else (n * r // n_ranks) // align * align for r in range(n_ranks + 1)]


def exchange_handles(handle: bytes, device=None) -> list[bytes]:
    """All-gather of one fixed-size byte string per rank over the default process group."""
    world = dist.get_world_size()
    t = torch.tensor(list(handle), dtype=torch.uint8, device=device if device is not None else "cpu")
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return [bytes(x.cpu().tolist()) for x in parts]


class PeerGroup:
    """This rank's GPU as a member of the box's GPUs (one process per GPU, launched by torchrun)."""

    def __init__(self, ctx, window_bytes: int, device=None):
        from . import ffi
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.peer = ffi.Peer(ctx, self.rank, self.world, window_bytes,
                             exchange=(lambda h: exchange_handles(h, device)) if self.world > 1 else None)
        if self.world > 1:
            dist.barrier()  # every rank has mapped every window before anyone pushes into one

    def close(self):
        if self.world > 1:
            dist.barrier()  # nobody unmaps a window a peer may still push into
        self.peer.close()
