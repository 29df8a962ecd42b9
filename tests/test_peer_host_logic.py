"""Host logic of the multi-GPU dist without a GPU: the tile plan (which member computes which output tile, which
arrival flags a tile waits for, in which order), the row blocks, and the handle exchange over gloo (world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plans(hg, world, sym, path, qb, hv_d):
    """every member's chunk rows, push units and ring flag (hg_peer_plan_push)"""
    out = [hg.ffi.peer_plan_push(world, r, sym, path, qb, hv_d) for r in range(world)]
    assert len({o[2] for o in out}) == 1
    return [o[0] for o in out], [o[1] for o in out], out[0][2]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("path", [2, 3])
@pytest.mark.parametrize("n,hv_d", [(5003, 4096), (1500, 4096), (20000, 8192)])
def test_all_vs_all_tiles_are_dealt_once_and_wait_for_the_right_rows(hg, world, path, n, hv_d):
    from hypergen_b200 import multigpu
    qb = multigpu.block_rows(n, world)
    rows, units, ring = _plans(hg, world, True, path, qb, hv_d)
    assert ring == (world >= 2 and n >= 512 * world)  # blocks are multiples of 256 rows from 512 rows per member on
    tr, tc = 256, (256 if path == 3 else 128)
    seen = {}
    counts = []
    for rank in range(world):
        # what reaches this member, in which order: sender -> position of (chunk -> me) in its unit list
        pos = {}
        for m in range(world):
            if m == rank:
                continue
            chunk_units = [(st, dest) for st, dest in units[m] if st < 4]
            assert units[m][0] == (4, sum(1 << t for t in range(world) if t != m))  # the start set goes to everybody, first
            for u, (st, dest) in enumerate(chunk_units):
                if dest >> rank & 1:
                    pos[(m, st)] = u
        R, Cc, need = hg.ffi.peer_plan_tiles(world, rank, True, path, 0, qb, hv_d)
        counts.append(R.size)
        keys = []
        for r, c, nd in zip(R.tolist(), Cc.tolist(), need.tolist()):
            assert (r, c) not in seen
            seen[(r, c)] = rank
            # the mask names exactly the other members' chunks that intersect the tile's rows or columns ...
            want = 0
            for lo, hi in ((r * tr, min((r + 1) * tr, n)), (c * tc, min((c + 1) * tc, n))):
                for m in range(world):
                    if m == rank:
                        continue
                    for ch in range(4):
                        a, b = rows[m][ch], rows[m][ch + 1]
                        if a < b and a < hi and b > lo:
                            want |= 1 << (4 * m + ch)
            assert nd == want
            # ... and every one of them is actually sent here
            k = 0
            for b in range(32):
                if nd >> b & 1:
                    assert (b // 4, b % 4) in pos, "member %d waits for chunk %d of member %d, which is never sent to it" % (rank, b % 4, b // 4)
                    k = max(k, 1 + b % 4)
            keys.append(k)
        assert keys == sorted(keys)  # own rows first, then in the order the chunks arrive
        for m in range(world):
            assert rows[m][0] == qb[m] and rows[m][4] == qb[m + 1] and rows[m] == sorted(rows[m])
    # every non-empty tile of the upper triangle exactly once (a tile is empty when its largest j <= its smallest i)
    gx, gy = -(-n // tc), -(-n // tr)
    want = {(r, c) for r in range(gy) for c in range(gx) if min((c + 1) * tc, n) - 1 > r * tr}
    assert set(seen) == want
    if ring:
        # a member's rows go to floor(N / 2) members only; the block pairs are spread evenly
        for m in range(world):
            dests = 0
            for st, dest in units[m][1:]:
                dests |= dest
            assert bin(dests).count("1") == world // 2
        if n >= 20000:
            assert max(counts) <= 1.25 * (sum(counts) / world)
    else:
        assert max(counts) - min(counts) <= 1  # dealt round-robin: balanced to within one tile


def test_chunks_are_at_least_about_two_megabytes(hg):
    from hypergen_b200 import multigpu
    # config 4 at 8 GPUs: 125 query rows per member -> one chunk; config 5: 2560 rows x 8192 x 2 planes -> four
    rows, units, ring = _plans(hg, 8, False, 3, multigpu.block_rows(1000, 8), 4096)
    assert all(r[1] == r[4] for r in rows) and not ring
    assert all(len(u) == 2 and u[1][0] == 0 for u in units)  # start set + one chunk, both to everybody
    rows, units, ring = _plans(hg, 8, True, 2, multigpu.block_rows(20000, 8), 8192)
    assert ring and all(len(set(r)) == 5 for r in rows) and all(len(u) == 1 + 4 for u in units)
    # member 3's rows go to the four members behind it on the ring (2, 1, 0, 7), chunk by chunk
    assert units[3][1:] == [(c, 1 << 2 | 1 << 1 | 1 << 0 | 1 << 7) for c in range(4)]


def test_ref_x_query_tiles_cover_the_members_rows_row_major(hg):
    from hypergen_b200 import multigpu
    world, n_qry, n_ref_local = 4, 1000, 700
    qb = multigpu.block_rows(n_qry, world)
    R, Cc, need = hg.ffi.peer_plan_tiles(world, 2, False, 3, n_ref_local, qb)
    assert list(zip(R.tolist(), Cc.tolist())) == [(r, c) for r in range(3) for c in range(4)]
    assert need[0] == (1 << 0) | (1 << 4)  # columns 0..255: member 0's and member 1's rows (one chunk each: 1 MB blocks)


def test_block_rows():
    from hypergen_b200 import multigpu
    assert multigpu.block_rows(10000, 8) == [0, 1280, 2560, 3840, 5120, 6400, 7680, 8960, 10000]  # multiples of 256 rows
    assert multigpu.block_rows(1000, 8) == [0, 124, 248, 372, 500, 624, 748, 872, 1000]            # too few rows: of 4
    assert multigpu.block_rows(5, 2) == [0, 0, 5]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypergen_b200 import multigpu
    got = multigpu.exchange_handles(bytes([17 * (rank + 1)] * 64))
    with open(os.path.join(out_dir, "r%d" % rank), "wb") as f:
        f.write(b"".join(got))
    dist.destroy_process_group()


def test_handle_exchange_world2_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    want = bytes([17] * 64) + bytes([34] * 64)
    assert open(tmp_path / "r0", "rb").read() == want and open(tmp_path / "r1", "rb").read() == want
