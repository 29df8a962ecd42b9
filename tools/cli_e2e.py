"""`hyper-gen sketch` / `hyper-gen dist` (the C++ host above the C ABI) end to end on FASTA files in /dev/shm:
wall-clock files/s as the reference logs them (src/sketch.rs:60-65).

    python tools/cli_e2e.py [n_files=512]
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypergen_b200 as hg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
exe = hg._build.build_host()
d = "/dev/shm/hg_cli_e2e"
os.makedirs(d, exist_ok=True)
rng = np.random.default_rng(1)
acgt = np.frombuffer(b"ACGT", np.uint8)
t0 = time.time()
base = acgt[rng.integers(0, 4, 5_000_000)]
for g in range(n):
    s = base.copy()
    m = rng.random(5_000_000) < 0.02 * (g % 10)
    s[m] = acgt[rng.integers(0, 4, int(m.sum()))]
    lines = np.concatenate([s.reshape(-1, 80), np.full((62500, 1), 10, np.uint8)], axis=1)
    with open("%s/g%05d.fna" % (d, g), "wb") as f:
        f.write(b">g%d\n" % g)
        f.write(lines.tobytes())
print("wrote %d files in %.1f s" % (n, time.time() - t0))
for threads in (16, 32):
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run([exe, "sketch", "-p", d, "-o", d + "/db.sketch", "-t", str(threads), "-D", "gpu"], capture_output=True, text=True, env=dict(os.environ, HG_CLI_TIMING="1"))
        dt = time.time() - t0
        assert r.returncode == 0, r.stderr[-500:]
    print(r.stderr.strip()[-300:])
    print("sketch -t %d: %.2f s wall, %.0f files/s (%.2f GB/s of FASTA)" % (threads, dt, n / dt, n * 5.0625e-3 / dt))
t0 = time.time()
r = subprocess.run([exe, "dist", "-r", d + "/db.sketch", "-q", d + "/db.sketch", "-o", d + "/ani.tsv", "-a", "85"], capture_output=True, text=True)
dt = time.time() - t0
assert r.returncode == 0, r.stderr[-500:]
print("dist: %.2f s wall, %d lines" % (dt, sum(1 for _ in open(d + "/ani.tsv"))))
subprocess.run(["rm", "-rf", d])
