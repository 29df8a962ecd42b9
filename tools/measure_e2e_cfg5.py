"""End-to-end `dist` for BASELINE config 5's dist half (20,000 sketches, D=8192, scaled=500: rows too wide for one
s8 plane) through hg_dist_packed: packed rows in pinned host memory -> sorted hits in host memory.
HG_DIST_STREAM=0 disables the chunked H2D / compute overlap for comparison.

    python tools/measure_e2e_cfg5.py [n=20000]
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
D, K = 8192, 21
packed_d = torch.empty((n, 2 * D), dtype=torch.uint8, device=dev)
bits_d = torch.empty(n, dtype=torch.uint8, device=dev)
norm_d = torch.empty(n, dtype=torch.int32, device=dev)
hv_d = torch.empty((5000, D), dtype=torch.int16, device=dev)
for s0 in range(0, n, 5000):
    m = min(5000, n - s0)
    sets = synth.hash_sets_family(m, n_per=10000, scaled=500, seed=0xD157 + s0)
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
del packed_d, hv_d
cap = 8_000_000
hits = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
milli = torch.empty(cap, dtype=torch.int32, pin_memory=True)
nh = C.c_uint64(0)
torch.cuda.synchronize()


def call():
    rc = lib.hg_dist_packed(ctx._h, rp.data_ptr(), width, rb.data_ptr(), rn.data_ptr(), n, rp.data_ptr(), width, rb.data_ptr(),
                            rn.data_ptr(), n, D, K, 85.0, 1, 0, 1, hits.data_ptr(), milli.data_ptr(), cap, C.byref(nh))
    if rc:
        raise RuntimeError(lib.hg_last_error().decode())


call()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    call()
    ts.append(time.perf_counter() - t0)
dt = min(ts)
pairs = n * (n - 1) // 2
h = np.frombuffer(hits[: nh.value * 16].numpy().tobytes(), dtype=hg.ffi.HIT_DTYPE)
print(json.dumps({"config": 5, "n": n, "bits_hist": np.bincount(rb.numpy()).tolist(), "stream": os.environ.get("HG_DIST_STREAM", "1"),
                  "ms": dt * 1e3, "pairs_per_s": pairs / dt, "h2d_bytes": int(rp.numel()), "hits": int(nh.value),
                  "hits_checksum": int(h["i"].astype(np.int64).sum() * 31 + h["j"].astype(np.int64).sum() + h["dot"].astype(np.int64).sum()),
                  "path": ctx.dist_last_path, "reason": ctx.dist_last_reason}))
