"""Static SASS mnemonic counts per kernel of the built library (evidence of tcgen05 / TMA / packed-FP32 code):
   python tools/sass_mnemonics.py > profiles/<tag>_sass_mnemonics.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pinn_for_quantum_wavefunction_surfaces_b200", "libpinn_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "LDGSTS", "HMMA", "USETMAXREG", "FFMA2", "FMUL2", "FADD2", "ELECT", "SYNCS",
        "BAR", "MUFU", "FFMA", "FMUL", "FADD", "LOP3", "LDS", "STS", "ACQBULK", "UTCATOMSWS", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("SASS mnemonic counts (cuobjdump -sass libpinn_b200.so), static, per kernel")
name, ops = None, collections.Counter()
def flush():
    if name and sum(ops.values()):
        print("%s  (%d SASS instructions)" % (name, sum(ops.values())))
        print("   " + "  ".join("%s %d" % (k, ops[k]) for k in WANT if ops[k]))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, ops = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        ops[m.group(1)] += 1
flush()
