"""Fixed vs per-K cost of the tensor dist kernel: time all-vs-all over n sketches for several D."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg
dev = torch.device("cuda", 0)
ctx = hg.Context(0); ctx.set_profiling(True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
hits = torch.empty(16 * 1_000_000, dtype=torch.uint8, device=dev); cnt = torch.zeros(1, dtype=torch.int64, device=dev)
for D in (256, 1024, 2048, 4096, 8192, 16384):
    hv = torch.randint(-300, 300, (n, D), dtype=torch.int16, device=dev)
    norm = (hv.to(torch.int64) ** 2).sum(1).to(torch.int32)
    ts = []
    for it in range(6):
        ctx.dist_dev(hv.data_ptr(), norm.data_ptr(), n, 0, hv.data_ptr(), norm.data_ptr(), n, 0, D, 21, 85.0, True, 2,
                     hits.data_ptr(), 1_000_000, cnt.data_ptr())
        ts.append(ctx.stage_ms()[3])
    tiles = (n // 128) * (n // 128 + 1) // 2
    t = min(ts[2:])
    print("D=%5d  stage %.3f ms  tiles %d  per-tile-per-SM %.2f us  kblocks %d" % (D, t, tiles, t * 1e3 * 148 / tiles, D // 128))
