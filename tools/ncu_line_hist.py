"""Per-source-line instruction/stall histogram from an .ncu-rep captured with --import-source on (-lineinfo build):
   python tools/ncu_line_hist.py rep [top]"""
import csv, subprocess, sys, io, collections
def I(v):
    try: return int(v)
    except Exception: return 0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; hdr = None; agg = []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    d = {k: v for k, v in zip(hdr[4:], r[4:])}
    stalls = {k.replace("stall_", ""): I(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
    agg.append((fname, int(r[0]), r[1].strip()[:90], I(d["Instructions Executed"]), I(d["# Samples"]), stalls))
tot_i = sum(a[3] for a in agg); tot_s = sum(a[4] for a in agg)
print("total instr %d samples %d" % (tot_i, tot_s))
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    st = ", ".join("%s %d" % kv for kv in sorted(a[5].items(), key=lambda kv: -kv[1])[:3] if kv[1])
    print("%-16s:%4d inst %5.1f%% smp %5.1f%% | %-28s | %s" % (a[0], a[1], 100.0 * a[3] / tot_i, 100.0 * a[4] / tot_s, st, a[2]))
