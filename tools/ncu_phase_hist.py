"""Samples / executed instructions along the SASS address order, in blocks, with the source lines each block maps to:
   python tools/ncu_phase_hist.py rep [block]"""
import csv, subprocess, sys, io, collections
def I(v):
    try: return int(v)
    except Exception: return 0
rep = sys.argv[1]; blk = int(sys.argv[2]) if len(sys.argv) > 2 else 150
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []; fname = "?"; hdr = None; cur_line = None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0].isdigit(): cur_line = (fname, int(r[0])); continue
    if r[2].startswith("0x"):
        d = {k: v for k, v in zip(hdr[4:], r[4:])}
        st = {k.replace("stall_", ""): I(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
        rows.append((int(r[2], 16), r[3].strip(), cur_line, I(d["Instructions Executed"]), I(d["# Samples"]), st))
rows.sort(key=lambda r: r[0])
tot_s = sum(r[4] for r in rows); tot_i = sum(r[3] for r in rows)
print("instructions in kernel: %d, executed %d, samples %d" % (len(rows), tot_i, tot_s))
for b in range(0, len(rows), blk):
    ch = rows[b:b + blk]
    s = sum(r[4] for r in ch); n = sum(r[3] for r in ch)
    lines = collections.Counter()
    for r in ch:
        if r[2] and r[2][0] == "pinn_step_tc.cu": lines[r[2][1]] += 1
    st = collections.Counter()
    for r in ch:
        for k, v in r[5].items(): st[k] += v
    ops = collections.Counter(r[1].split()[1].split(".")[0] if r[1].startswith("@") else r[1].split()[0].split(".")[0] for r in ch)
    lo = min(lines) if lines else 0; hi = max(lines) if lines else 0
    print("sass %5d-%5d  exec %5.1f%%  smp %5.1f%% (%.2f smp/kinst) lines %3d-%3d | %s | %s" % (
        b, b + len(ch), 100.0 * n / tot_i, 100.0 * s / tot_s, 1000.0 * s / max(n, 1), lo, hi,
        ", ".join("%s %d" % kv for kv in st.most_common(4)), " ".join("%s:%d" % kv for kv in ops.most_common(4))))
