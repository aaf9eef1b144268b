"""Per-kernel SASS mnemonic counts of machupx_b200/liblls.so (cuobjdump -sass): the evidence that the FP64 tensor path (DMMA), the
asynchronous copies (LDGSTS = cp.async; UBLKCP / UTMALDG / UTMASTG = TMA bulk / tensor copies) and the system-scope flag
instructions of the peer-memory exchange are in the shipped binary.  usage: python scripts/summarise_sass.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "machupx_b200", "liblls.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(re.findall(r"Function : (\S+)", out), demangle))
WATCH = ("DMMA", "DFMA", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU.RCP64H", "SHFL", "BAR", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS",
         "ATOM", "RED", "MEMBAR", "ERRBAR", "NANOSLEEP")
arch = re.search(r"arch = (\S+)", out)
print("# %s   (%s)" % (os.path.relpath(lib, ROOT), arch.group(1) if arch else "?"))
print("# columns: total SASS instructions, then counts of the mnemonics that matter")
cur, counts = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for w in WATCH:
            if op.startswith(w) or (w in ("ST.E.STRONG.SYS", "LD.E.STRONG.SYS") and op.startswith(w.split(".")[0]) and "STRONG.SYS" in op):
                counts[cur][w] += 1
tot = collections.Counter()
for fn in sorted(counts, key=lambda f: -counts[f]["total"]):
    c = counts[fn]
    full = names.get(fn, fn)
    short = re.sub(r">\(.*", ">", full) if ">(" in full else re.sub(r"\(.*", "", full)
    short = short.replace("(bool)", "").replace("(int)", "").replace("void ", "")
    print("%-70s total %6d  " % (short[:70], c["total"]) + "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]))
    tot.update(c)
print("%-70s total %6d  " % ("ALL KERNELS", tot["total"]) + "  ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w]))
