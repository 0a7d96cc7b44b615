"""Executed instructions and stall samples per source function (and per line for the top ones) of the first launch in a report:
   python tools/ncu_func_hist.py rep [lines]"""
import csv, subprocess, io, collections, re, sys, os
rep = sys.argv[1]; want_lines = len(sys.argv) > 2
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def I(v):
    try: return int(v)
    except Exception: return 0
per_line = collections.OrderedDict(); fname = None; seen = set(); skip = False
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path":
        fname = r[1]; skip = fname in seen; seen.add(fname); continue   # the second launch repeats the files
    if skip or len(r) < 8 or not r[0].isdigit(): continue
    per_line[(fname, int(r[0]))] = (I(r[7]), I(r[6]), r[1].strip())
funcs = {}
def fn(path, line):
    if path not in funcs:
        fl = []
        try:
            for i, l in enumerate(open(path).read().split("\n"), 1):
                m = re.match(r"(?:template.*>\s*)?__(?:device|global)__ .*?\b(\w+)\(", l)
                if m: fl.append((i, m.group(1)))
        except OSError: pass
        funcs[path] = fl
    name = "?"
    for i, n in funcs[path]:
        if i <= line: name = n
    return name
agg = collections.Counter(); smp = collections.Counter()
tot = sum(v[0] for v in per_line.values()); ts = sum(v[1] for v in per_line.values())
for (path, line), (n, s, _) in per_line.items():
    k = (os.path.basename(path), fn(path, line)); agg[k] += n; smp[k] += s
print("warp-instructions %d, samples %d" % (tot, ts))
for k, n in agg.most_common(30):
    print("%-18s %-24s inst %5.1f%%  samples %5.1f%%" % (k[0], k[1], 100.0 * n / tot, 100.0 * smp[k] / max(ts, 1)))
if want_lines:
    print("---- hottest lines by samples ----")
    for (path, line), (n, s, txt) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-16s %4d inst %4.1f%% smp %4.1f%%  %s" % (os.path.basename(path), line, 100.0 * n / tot, 100.0 * s / max(ts, 1), txt[:90]))
