"""Fixed vs per-K cost of the single-plane dist path: all-vs-all over n narrow rows for several D (stage = pre-pass + kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg
dev = torch.device("cuda", 0)
ctx = hg.Context(0); ctx.set_profiling(True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
hits = torch.empty(16 * 1_000_000, dtype=torch.uint8, device=dev); cnt = torch.zeros(1, dtype=torch.int64, device=dev)
for D in (256, 1024, 2048, 4096, 8192):
    hv = (2 * torch.randint(-110, 110, (n, D), device=dev) + torch.randint(0, 2, (n, 1), device=dev)).to(torch.int16)
    norm = (hv.to(torch.int64) ** 2).sum(1).to(torch.int32)
    ts = []
    for it in range(6):
        ctx.dist_dev(hv.data_ptr(), norm.data_ptr(), n, 0, hv.data_ptr(), norm.data_ptr(), n, 0, D, 21, 85.0, True, 3,
                     hits.data_ptr(), 1_000_000, cnt.data_ptr())
        ctx.sync()
        ts.append(ctx.stage_ms()[3])
    t = min(ts[2:])
    tiles = sum(((n + 255) // 256) - r for r in range((n + 255) // 256))
    print("D=%5d  stage %.3f ms  tiles %d  waves %.2f  per-wave %.2f us  kblocks %d" % (D, t, tiles, tiles / 74, t * 1e3 / -(-tiles // 74), D // 128))
