"""debug: group (one process, 2 GPUs) symmetric dist vs single GPU - which hits differ"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hypergen_b200 as hg
from hypergen_b200 import synth
n, D, n_per, scaled = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 1500
with hg.Context(0) as c0:
    hashes, off = synth.hash_sets_family_dev(n, n_per=n_per, scaled=scaled)
    sets = [hashes.numpy().view(np.uint64)[int(off[g]):int(off[g + 1])] for g in range(n)]
    r = c0.encode_sets(sets, hv_d=D)
    packed, bits, norm = r["packed"], r["quant_bits"], r["norm2"]
    want, _ = c0.dist_packed(packed, bits, norm, packed, bits, norm, D, ksize=21, ani_th=70.0, symmetric=True, sorted_output=True)
with hg.Group(2) as g:
    for rep in range(2):
        got, _ = g.dist_packed(packed, bits, norm, packed, bits, norm, D, ksize=21, ani_th=70.0, symmetric=True, sorted_output=True)
        ws = np.sort(want, order=["i", "j"]); gs = np.sort(got, order=["i", "j"])
        wk = ws["i"].astype(np.int64) * n + ws["j"]; gk = gs["i"].astype(np.int64) * n + gs["j"]
        extra = gs[~np.isin(gk, wk)]; missing = ws[~np.isin(wk, gk)]
        dup = gk.size - np.unique(gk).size
        print("rep", rep, "want", want.size, "got", got.size, "extra", extra.size, "missing", missing.size, "dups", dup)
        if extra.size:
            print(" extra i range", extra["i"].min(), extra["i"].max(), "j range", extra["j"].min(), extra["j"].max())
            print(" extra tile rows", np.unique(extra["i"] // 256)[:20], "tile cols", np.unique(extra["j"] // 256)[:20])
            print(extra[:5])
        both = np.isin(gk, wk)
        common_w = ws[np.isin(wk, gk)]
        common_g = gs[both]
        if common_g.size == common_w.size:
            bad = (common_g["dot"] != common_w["dot"])
            print(" common", common_g.size, "dot mismatches", int(bad.sum()))
            if bad.any():
                b = common_g[bad]
                print("  bad tile rows", np.unique(b["i"] // 256)[:20], "cols", np.unique(b["j"] // 256)[:20])
