"""Timing-only experiments on the dist tensor kernel (HG_DIST_DBG variants give wrong results on purpose).
usage: dist_dbg.py [n] [real|rand]"""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypergen_b200 as hg
from hypergen_b200 import synth
n, D = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(os.environ.get("HVD", "4096"))
kind = sys.argv[2] if len(sys.argv) > 2 else "real"
ctx = hg.Context(0)
if kind == "rand":
    g = torch.Generator(device="cuda").manual_seed(1)
    hv = torch.randint(-200, 200, (n, D), dtype=torch.int16, device="cuda", generator=g)
    norm = (hv.int() ** 2).sum(1).int()
else:
    sets = synth.hash_sets_family(n)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in sets])
    hashes = torch.from_numpy(np.concatenate(sets).view(np.int64)).cuda()
    hv = torch.empty((n, D), dtype=torch.int16, device="cuda")
    bits = torch.empty(n, dtype=torch.uint8, device="cuda")
    norm = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.encode_sets_dev(hashes.data_ptr(), off, D, hv.data_ptr(), None, bits.data_ptr(), norm.data_ptr())
    ctx.sync()
cap = 1 << 22
hits = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk, pw, stop = [], [], [False]
def sample():
    while not stop[0]:
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
        time.sleep(0.01)
ctx.set_profiling(True)
ms = []
th = threading.Thread(target=sample); th.start()
for it in range(int(os.environ.get("ITERS", "200"))):
    ctx.dist_dev(hv.data_ptr(), norm.data_ptr(), n, 0, hv.data_ptr(), norm.data_ptr(), n, 0, D, 21, float(os.environ.get("TH", "85.0")), True, 2,
                 hits.data_ptr(), cap, cnt.data_ptr())
    ctx.sync()
    ms.append(ctx.stage_ms()[3])
stop[0] = True; th.join()
ms = ms[5:]
print("%s DBG=%s KERNEL=%s: ms first %.4f min %.4f median %.4f last %.4f | SM MHz median %d min %d | W max %.0f | hits %d" % (
    kind, os.environ.get("HG_DIST_DBG"), os.environ.get("HG_DIST_KERNEL"), ms[0], min(ms), sorted(ms)[len(ms) // 2], ms[-1],
    sorted(clk)[len(clk) // 2], min(clk), max(pw), int(cnt.item())))
