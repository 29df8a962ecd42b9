"""Host logic of the multi-GPU dist without a GPU: the tile plan (which member computes which output tile, which
arrival flags a tile waits for, in which order), the row blocks, and the handle exchange over gloo (world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chunk_lo(qb, m, c):
    return qb[m + 1] if c >= 4 else qb[m] + (((qb[m + 1] - qb[m]) * c // 4) & ~3)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("path", [2, 3])
def test_all_vs_all_tiles_are_dealt_once_and_wait_for_the_right_rows(hg, world, path):
    from hypergen_b200 import multigpu
    n = 5003
    qb = multigpu.block_rows(n, world)
    tr, tc = 256, (256 if path == 3 else 128)
    seen = {}
    counts = []
    for rank in range(world):
        R, Cc, need = hg.ffi.peer_plan_tiles(world, rank, True, path, 0, qb)
        counts.append(R.size)
        keys = []
        for r, c, nd in zip(R.tolist(), Cc.tolist(), need.tolist()):
            assert (r, c) not in seen
            seen[(r, c)] = rank
            # the mask names exactly the other members' chunks that intersect the tile's rows or columns
            want = 0
            for lo, hi in ((r * tr, min((r + 1) * tr, n)), (c * tc, min((c + 1) * tc, n))):
                for m in range(world):
                    if m == rank:
                        continue
                    for ch in range(4):
                        a, b = _chunk_lo(qb, m, ch), _chunk_lo(qb, m, ch + 1)
                        if a < b and a < hi and b > lo:
                            want |= 1 << (4 * m + ch)
            assert nd == want
            keys.append(0 if nd == 0 else 1 + max(b % 4 for b in range(32) if nd >> b & 1))
        assert keys == sorted(keys)  # own rows first, then in the order the chunks arrive
    # every non-empty tile of the upper triangle exactly once (a tile is empty when its largest j <= its smallest i)
    gx, gy = -(-n // tc), -(-n // tr)
    want = {(r, c) for r in range(gy) for c in range(gx) if min((c + 1) * tc, n) - 1 > r * tr}
    assert set(seen) == want
    assert max(counts) - min(counts) <= 1  # dealt round-robin: balanced to within one tile


def test_ref_x_query_tiles_cover_the_members_rows_row_major(hg):
    from hypergen_b200 import multigpu
    world, n_qry, n_ref_local = 4, 1000, 700
    qb = multigpu.block_rows(n_qry, world)
    R, Cc, need = hg.ffi.peer_plan_tiles(world, 2, False, 3, n_ref_local, qb)
    assert list(zip(R.tolist(), Cc.tolist())) == [(r, c) for r in range(3) for c in range(4)]
    assert need[0] == sum(1 << (4 * 0 + ch) for ch in range(4)) | (1 << 4)  # columns 0..255: all of member 0, first chunk of member 1


def test_block_rows():
    from hypergen_b200 import multigpu
    assert multigpu.block_rows(10000, 8) == [0, 1248, 2500, 3748, 5000, 6248, 7500, 8748, 10000]
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
