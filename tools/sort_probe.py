import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import hypergen_b200 as hg
ctx = hg.Context(0)
n = 185057
rng = np.random.default_rng(1)
hits = np.zeros(n, hg.ffi.HIT_DTYPE)
pair = rng.choice(10000 * 10000, size=n, replace=False)
hits["i"], hits["j"] = pair // 10000, pair % 10000
hits["ani"] = (85 + 15 * rng.random(n)).astype(np.float32)
base = torch.from_numpy(hits.view(np.uint8).reshape(-1).copy()).cuda()
d = base.clone(); m = torch.zeros(n, dtype=torch.int32, device="cuda")
for it in range(8):
    d.copy_(base); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.sort_hits_dev(d.data_ptr(), n, m.data_ptr()); ctx.sync()
    dt = time.perf_counter() - t0
print(os.environ.get("HG_SORT_MULTI"), os.environ.get("HG_SORT_ROUNDS"), "last call %.1f us" % (dt * 1e6))
