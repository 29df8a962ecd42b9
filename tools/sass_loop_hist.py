"""Opcode histogram of the biggest backward-branch loop of a kernel (the 8-position group of kmer_hash_kernel).
    python tools/sass_loop_hist.py <lib.so> <mangled-name-substring>"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.+?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and (best is None or a - tgt > best[1] - best[0]):
            best = (tgt, a)
body = [t for a, t in ins if best[0] <= a <= best[1]]
# the rare paths (survivor handling) are the long forward-branch spans inside the loop: report the main path without them
skips = []
for a, t in ins:
    if best[0] <= a <= best[1]:
        m = re.search(r"BRA(?:\.U)?(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m and t.startswith("@"):
            tgt = int(m.group(1), 16)
            if a < tgt <= best[1] and tgt - a > 40 * 16 and not any(lo < a < hi for lo, hi in skips):
                skips.append((a, tgt))
if skips:
    main = [t for a, t in ins if best[0] <= a <= best[1] and not any(lo < a < hi for lo, hi in skips)]
    print("main path (without the survivor branches %s): %d instructions, %.1f per k-mer" % (
        ", ".join("0x%x..0x%x" % x for x in skips), len(main), len(main) / 8))
    body = main
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
wide = sum(1 for t in body if "IMAD.WIDE" in t)
print("loop 0x%x..0x%x: %d instructions (%.1f per k-mer at 8 positions per trip), IMAD.WIDE %d" % (best[0], best[1], len(body), len(body) / 8, wide))
for k, v in ops.most_common(14):
    print("  %-8s %4d  %.1f / k-mer" % (k, v, v / 8))
