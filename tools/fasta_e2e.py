"""Time the raw-FASTA entry (hg_sketch_fasta_batch) end to end from pinned host memory: 1000 files of
5 Mbp in 80-column lines.  Prints genomes/s and the stage times."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg
from hypergen_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = hg.Context(0)
seq = synth.genome(0xB200, 5_000_000).numpy()
lines = seq.reshape(-1, 80)
body = np.concatenate([lines, np.full((lines.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
one = np.concatenate([np.frombuffer(b">genome\n", np.uint8), body])
raw = torch.empty(one.size * n, dtype=torch.uint8).pin_memory()
rn = raw.numpy()
for i in range(n):
    rn[i * one.size:(i + 1) * one.size] = one
off = (np.arange(n + 1, dtype=np.uint64) * one.size)
p = hg.make_params()
lib = hg.ffi.load()
import ctypes as C
packed = torch.empty((n, 2 * 4096), dtype=torch.uint8).pin_memory()
bits = torch.empty(n, dtype=torch.uint8).pin_memory()
norm = torch.empty(n, dtype=torch.int32).pin_memory()
nh = torch.empty(n, dtype=torch.int32).pin_memory()
def run():
    rc = lib.hg_sketch_fasta_batch(ctx._h, C.c_void_p(raw.data_ptr()), C.c_void_p(off.ctypes.data), n, C.byref(p), None,
                                   C.c_void_p(packed.data_ptr()), C.c_void_p(bits.data_ptr()), C.c_void_p(norm.data_ptr()),
                                   C.c_void_p(nh.data_ptr()))
    assert rc == 0, lib.hg_last_error()
for _ in range(2):
    run()
ts = []
for _ in range(3):
    t = time.perf_counter(); run(); ts.append(time.perf_counter() - t)
t = min(ts)
print("raw-FASTA e2e: %.1f ms / %d files = %.0f genomes/s, %.1f GB/s of file bytes; n_hashes[0]=%d" %
      (t * 1e3, n, n / t, raw.numel() / t / 1e9, int(nh[0])))
