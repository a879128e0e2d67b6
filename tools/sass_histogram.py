"""SASS opcode histogram per kernel of the built library (evidence for profiles/):
python tools/sass_histogram.py sequnwinder_b200/libsequnwinder_b200.so > profiles/r02_sass_opcodes.txt
Counts the mnemonics that say how a kernel moves and multiplies: DMMA (fp64 tensor MMA), UBLKCP
(cp.async.bulk, the TMA engine's 1-D copy), SYNCS (mbarrier), I2F (byte -> double), LDS/STS, LDG/STG,
ATOMS, RED, BAR, plus spills (LDL/STL)."""
import collections, re, subprocess, sys

lib = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
keys = ["DMMA", "UBLKCP", "UTMALDG", "SYNCS", "I2F", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG",
        "ATOMS", "ATOMG", "RED", "REDUX", "BAR", "SHFL", "LDL", "STL"]
print("kernel | total | " + " | ".join(keys))
for f in sorted(hist):
    h = hist[f]
    print(f"{f} | {sum(h.values())} | " + " | ".join(str(h.get(k, 0)) for k in keys))
