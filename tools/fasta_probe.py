"""Launch the FASTA-merge kernels on 100 x 5 Mbp files (80-column lines) for an ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg
from hypergen_b200 import synth
ctx = hg.Context(0)
seq = synth.genome(0xB200, 5_000_000).numpy()
lines = seq.reshape(-1, 80)
body = np.concatenate([lines, np.full((lines.shape[0], 1), 10, np.uint8)], axis=1).tobytes()
files = [b">g%d\n" % i + body for i in range(100)]
for _ in range(3):
    r = ctx.sketch_fasta_batch(files, hg.make_params(), want_hv=False)
print(r["n_hashes"][:4], len(files[0]))
