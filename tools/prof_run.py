"""Small driver for ncu captures: a few launches of each hot kernel on resident inputs.

    ncu --set full --clock-control none --import-source on -k regex:kmer_hash -s 1 -c 1 \
        -o gpurun_out/prof_kmer python tools/prof_run.py --genomes 200
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg  # noqa: E402
from hypergen_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=200)
ap.add_argument("--dist-n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--path", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = hg.Context(0)
p = hg.make_params()
n, L, D = a.genomes, 5_000_000, 4096
seq, off = synth.family_batch(n, L, device=dev)
hv = torch.empty((n, D), dtype=torch.int16, device=dev)
packed = torch.empty((n, 2 * D), dtype=torch.uint8, device=dev)
bits = torch.empty(n, dtype=torch.uint8, device=dev)
norm = torch.empty(n, dtype=torch.int32, device=dev)
nh = torch.empty(n, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for _ in range(a.reps):
    ctx.sketch_batch_dev(seq.data_ptr(), off, p, hv.data_ptr(), packed.data_ptr(), bits.data_ptr(), norm.data_ptr(),
                         nh.data_ptr())
ctx.sync()
if a.dist_n:
    m = a.dist_n
    sets = synth.hash_sets_family(m)
    o = np.zeros(m + 1, np.uint64)
    o[1:] = np.cumsum([len(s) for s in sets])
    hashes = torch.from_numpy(np.concatenate(sets).view(np.int64)).to(dev)
    qhv = torch.empty((m, D), dtype=torch.int16, device=dev)
    qb = torch.empty(m, dtype=torch.uint8, device=dev)
    qn = torch.empty(m, dtype=torch.int32, device=dev)
    ctx.encode_sets_dev(hashes.data_ptr(), o, D, qhv.data_ptr(), None, qb.data_ptr(), qn.data_ptr())
    hits = torch.empty(16 * 4_000_000, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    for _ in range(a.reps):
        ctx.dist_dev(qhv.data_ptr(), qn.data_ptr(), m, 0, qhv.data_ptr(), qn.data_ptr(), m, 0, D, 21, 85.0, True, a.path,
                     hits.data_ptr(), 4_000_000, cnt.data_ptr())
    ctx.sync()
    print("dist path", ctx.dist_last_path, ctx.dist_last_reason, "hits", int(cnt.item()))
print("done", ctx.launches, "launches")
