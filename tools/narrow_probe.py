"""Single-plane (path 3) vs two-limb (path 2) tensor dist kernels on config-3 / config-4 shaped inputs:
same hits, stage time (pre-pass + kernel), outlier statistics.

    python tools/narrow_probe.py [n_sketches=10000] [--cfg4]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg  # noqa: E402
from hypergen_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
ctx = hg.Context(0)
ctx.set_profiling(True)
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10000
D = 4096


def encode(n, chunk=10000):
    hv = torch.empty((n, D), dtype=torch.int16, device=dev)
    norm = torch.empty(n, dtype=torch.int32, device=dev)
    bits = torch.empty(n, dtype=torch.uint8, device=dev)
    for s0 in range(0, n, chunk):
        m = min(chunk, n - s0)
        sets = synth.hash_sets_family(m, seed=0xD157 + s0) if s0 else synth.hash_sets_family(m)
        off = np.zeros(m + 1, np.uint64)
        off[1:] = np.cumsum([len(s) for s in sets])
        hashes = torch.from_numpy(np.concatenate(sets).view(np.int64)).to(dev)
        ctx.encode_sets_dev(hashes.data_ptr(), off, D, hv[s0:].data_ptr(), None, bits[s0:].data_ptr(), norm[s0:].data_ptr())
        ctx.sync()
    return hv, norm, bits


def run(r_hv, r_norm, q_hv, q_norm, sym, path, cap=8_000_000, reps=6):
    hits = torch.empty(cap * 16, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ts = []
    for _ in range(reps):
        ctx.dist_dev(r_hv.data_ptr(), r_norm.data_ptr(), r_hv.shape[0], 0, q_hv.data_ptr(), q_norm.data_ptr(), q_hv.shape[0], 0,
                     D, 21, 85.0, sym, path, hits.data_ptr(), cap, cnt.data_ptr())
        ctx.sync()
        ts.append(ctx.stage_ms()[3])
    c = int(cnt.item())
    h = np.frombuffer(hits[: c * 16].cpu().numpy().tobytes(), dtype=hg.ffi.HIT_DTYPE)
    h = np.sort(h, order=["i", "j"])
    return min(ts[2:]), h, ctx.dist_last_path, ctx.dist_last_reason


hv, norm, bits = encode(n)
rng = hv.to(torch.int32)
span = (rng.max(1).values - rng.min(1).values).cpu().numpy()
print(json.dumps({"n": n, "bits_hist": np.bincount(bits.cpu().numpy()).tolist(), "rows_span_gt_510": int((span > 510).sum()),
                  "span_max": int(span.max())}))
if "--cfg4" in sys.argv:
    q_hv, q_norm = hv[5:n:100].contiguous(), norm[5:n:100].contiguous()
    cases = [("cfg4 %d x %d" % (n, q_hv.shape[0]), hv, norm, q_hv, q_norm, False)]
else:
    cases = [("cfg3 all-vs-all %d" % n, hv, norm, hv, norm, True)]
for name, r, rn, q, qn, sym in cases:
    pairs = r.shape[0] * (r.shape[0] - 1) // 2 if sym else r.shape[0] * q.shape[0]
    res = {}
    for path in (2, 3, 0):
        ms, h, used, why = run(r, rn, q, qn, sym, path)
        res[path] = h
        print(json.dumps({"case": name, "path_req": path, "path": used, "stage_ms": ms, "pairs_per_s": pairs / ms * 1e3,
                          "alg_TOPs": 2 * D * pairs / ms * 1e9 / 1e12 * 1e-6, "hits": int(h.size), "reason": why}))
    same = res[2].size == res[3].size and np.array_equal(res[2], res[3])
    print(json.dumps({"case": name, "path3_equals_path2": bool(same)}))
    assert same
