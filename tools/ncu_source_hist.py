"""Opcode / stall histogram from `ncu --page source --csv` of a report: python tools/ncu_source_hist.py rep"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
stall_by_op = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "MUFU", "SHFL", "LDG", "BAR")) and "." in op else "")
    try:
        n = int(r[iex] or 0); s = int(r[ismp] or 0)
    except ValueError:
        continue   # header row of the next launch in the report
    ops[op] += n; smp[op] += s; tot += n; tots += s
    for i, h in stall_cols:
        v = int(r[i] or 0) if (r[i] or "0").lstrip("-").isdigit() else 0
        if v: stall_by_op[op][h] += v
print("total warp-instructions", tot, "samples", tots)
for op, n in ops.most_common(28):
    top = ", ".join("%s %d" % (h.replace("stall_", ""), v) for h, v in stall_by_op[op].most_common(3))
    print("%-14s %11d %5.1f%%  samples %5.1f%%   %s" % (op, n, 100.0 * n / tot, 100.0 * smp[op] / max(tots, 1), top))
