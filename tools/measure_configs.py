"""One-off measurements of BASELINE configs 4 and 5 on one B200 (for DESIGN.md; not bench lines).

    python tools/measure_configs.py 4     # 100,000 ref x 1,000 queries, D=4096, ani_th=85
    python tools/measure_configs.py 5     # scaled=500, D=8192: sketch 20,000 genomes + all-vs-all
"""
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
ctx.set_profiling(True)
which = sys.argv[1] if len(sys.argv) > 1 else "4"


def encode_family_sets(n, D, n_per=3333, scaled=1500, chunk=10000):
    hv = torch.empty((n, D), dtype=torch.int16, device=dev)
    norm = torch.empty(n, dtype=torch.int32, device=dev)
    bits = torch.empty(n, dtype=torch.uint8, device=dev)
    for s0 in range(0, n, chunk):
        m = min(chunk, n - s0)
        sets = synth.hash_sets_family(m, n_per=n_per, scaled=scaled, seed=0xD157 + s0)
        off = np.zeros(m + 1, np.uint64)
        off[1:] = np.cumsum([len(s) for s in sets])
        hashes = torch.from_numpy(np.concatenate(sets).view(np.int64)).to(dev)
        ctx.encode_sets_dev(hashes.data_ptr(), off, D, hv[s0:].data_ptr(), None, bits[s0:].data_ptr(), norm[s0:].data_ptr())
        ctx.sync()
    return hv, norm, bits


def time_dist(r_hv, r_norm, q_hv, q_norm, D, symmetric, reps=5, cap=8_000_000):
    hits = torch.empty(cap * 16, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    best, path = 1e9, 0
    for it in range(reps + 2):
        ctx.dist_dev(r_hv.data_ptr(), r_norm.data_ptr(), r_hv.shape[0], 0, q_hv.data_ptr(), q_norm.data_ptr(), q_hv.shape[0], 0,
                     D, 21, 85.0, symmetric, path, hits.data_ptr(), cap, cnt.data_ptr())
        ms = ctx.stage_ms()[3]
        path = ctx.dist_last_path
        if it >= 2:
            best = min(best, ms)
    return best, int(cnt.item()), path, ctx.dist_last_reason


if which == "4":
    D = 4096
    r_hv, r_norm, _ = encode_family_sets(100_000, D)
    q_hv, q_norm = r_hv[5:100_000:100].contiguous(), r_norm[5:100_000:100].contiguous()  # 1,000 queries drawn from the families
    ms, nh, path, why = time_dist(r_hv, r_norm, q_hv, q_norm, D, False)
    pairs = r_hv.shape[0] * q_hv.shape[0]
    print(json.dumps({"config": 4, "pairs": pairs, "kernel_ms": ms, "pairs_per_s": pairs / ms * 1e3, "hits": nh, "path": path,
                      "reason": why, "alg_TOPs": 2 * D * pairs / ms * 1e3 / 1e12}))
else:
    D, scaled, n, L, batch = 8192, 500, 20_000, 5_000_000, 500
    p = hg.make_params(scaled=scaled, hv_d=D)
    hv = torch.empty((n, D), dtype=torch.int16, device=dev)
    packed = torch.empty((batch, 2 * D), dtype=torch.uint8, device=dev)
    bits = torch.empty(n, dtype=torch.uint8, device=dev)
    norm = torch.empty(n, dtype=torch.int32, device=dev)
    nh = torch.empty(n, dtype=torch.int32, device=dev)
    t_sk = 0.0
    for b0 in range(0, n, batch):
        seq, off = synth.family_batch(batch, L, device=dev, first=b0)
        torch.cuda.synchronize()
        ctx.sketch_batch_dev(seq.data_ptr(), off, p, hv[b0:].data_ptr(), packed.data_ptr(), bits[b0:].data_ptr(),
                             norm[b0:].data_ptr(), nh[b0:].data_ptr())
        st = ctx.stage_ms()
        ctx.sketch_status()
        t_sk += st[0] + st[1] + st[2]
        if b0 == 0:
            first = st
        del seq
    ms, nhits, path, why = time_dist(hv, norm, hv, norm, D, True)
    pairs = n * (n - 1) // 2
    print(json.dumps({"config": 5, "genomes": n, "sketch_ms_total": t_sk, "genomes_per_s": n / t_sk * 1e3, "first_batch_stage_ms": first,
                      "quant_bits_hist": np.bincount(bits.cpu().numpy()).tolist(), "mean_hashes": float(nh.float().mean()),
                      "pairs": pairs, "dist_kernel_ms": ms, "pairs_per_s": pairs / ms * 1e3, "hits": nhits, "path": path, "reason": why,
                      "alg_TOPs": 2 * D * pairs / ms * 1e3 / 1e12}))
