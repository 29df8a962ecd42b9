"""End-to-end `dist` for BASELINE config 4 (100,000 ref sketches x 1,000 queries, D=4096) through hg_dist_packed:
packed sketch rows in pinned host memory -> sorted hits in host memory.  HG_DIST_STREAM=0 disables the chunked
H2D / compute overlap for comparison.

    python tools/measure_e2e_cfg4.py [n_ref=100000]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg  # noqa: E402
from hypergen_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
ctx = hg.Context(0)
lib = hg.ffi.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
D, K = 4096, 21
packed_d = torch.empty((n, 2 * D), dtype=torch.uint8, device=dev)
bits_d = torch.empty(n, dtype=torch.uint8, device=dev)
norm_d = torch.empty(n, dtype=torch.int32, device=dev)
hv_d = torch.empty((10000, D), dtype=torch.int16, device=dev)
for s0 in range(0, n, 10000):
    m = min(10000, n - s0)
    sets = synth.hash_sets_family(m, seed=0xD157 + s0) if s0 else synth.hash_sets_family(m)
    off = np.zeros(m + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in sets])
    hashes = torch.from_numpy(np.concatenate(sets).view(np.int64)).to(dev)
    ctx.encode_sets_dev(hashes.data_ptr(), off, D, hv_d.data_ptr(), packed_d[s0:].data_ptr(), bits_d[s0:].data_ptr(),
                        norm_d[s0:].data_ptr())
    ctx.sync()
width = int(bits_d.max().item()) * D // 8
rp = torch.empty((n, width), dtype=torch.uint8, pin_memory=True)
rp.copy_(packed_d[:, :width])
rb = torch.empty(n, dtype=torch.uint8, pin_memory=True); rb.copy_(bits_d)
rn = torch.empty(n, dtype=torch.int32, pin_memory=True); rn.copy_(norm_d)
q = torch.arange(5, n, 100)
qp = torch.empty((q.numel(), width), dtype=torch.uint8, pin_memory=True); qp.copy_(rp[q])
qb = torch.empty(q.numel(), dtype=torch.uint8, pin_memory=True); qb.copy_(rb[q])
qn = torch.empty(q.numel(), dtype=torch.int32, pin_memory=True); qn.copy_(rn[q])
del packed_d, hv_d
cap = 4_000_000
hits = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
milli = torch.empty(cap, dtype=torch.int32, pin_memory=True)
nh = C.c_uint64(0)
torch.cuda.synchronize()


def call():
    rc = lib.hg_dist_packed(ctx._h, rp.data_ptr(), width, rb.data_ptr(), rn.data_ptr(), n, qp.data_ptr(), width, qb.data_ptr(),
                            qn.data_ptr(), q.numel(), D, K, 85.0, 0, 0, 1, hits.data_ptr(), milli.data_ptr(), cap, C.byref(nh))
    if rc:
        raise RuntimeError(lib.hg_last_error().decode())


call()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    call()
    ts.append(time.perf_counter() - t0)
dt = min(ts)
pairs = n * q.numel()
print(json.dumps({"config": 4, "n_ref": n, "n_qry": int(q.numel()), "stream": os.environ.get("HG_DIST_STREAM", "1"),
                  "ms": dt * 1e3, "pairs_per_s": pairs / dt, "h2d_bytes": int(rp.numel() + qp.numel()),
                  "h2d_gbs_equiv": (rp.numel() + qp.numel()) / dt / 1e9, "hits": int(nh.value), "path": ctx.dist_last_path,
                  "reason": ctx.dist_last_reason}))
