"""SASS opcode histogram of the built library (run here, no GPU needed): per kernel family the counts of the Blackwell-specific
opcodes that prove tcgen05 / TMEM / TMA use, plus the top opcodes overall.
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import os, re, subprocess, sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "diff-reg_b200", "libdiffreg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fam = None
per = defaultdict(Counter)
op_re = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fam = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "").replace("drg::", "")
        continue
    m = op_re.match(line)
    if m and fam:
        per[fam][m.group(1).split(".")[0]] += 1
        if m.group(1).split(".")[0] in ("UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "UTCATOMSWS"):
            per[fam]["*" + m.group(1)] += 1
special = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCATOMSWS", "MUFU", "REDG", "ATOMG", "HMMA", "DFMA")
print(f"# SASS opcode histogram of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; all template instantiations of a kernel summed)")
print("# UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTCATOMSWS = tcgen05.alloc / dealloc")
for f in sorted(per):
    c = per[f]
    tot = sum(v for k, v in c.items() if not k.startswith("*"))
    sp = "  ".join(f"{k}={c[k]}" for k in special if c.get(k))
    print(f"\n{f}: {tot} instructions\n  special: {sp or '-'}")
    det = "  ".join(f"{k[1:]}={v}" for k, v in sorted(c.items()) if k.startswith("*"))
    if det:
        print(f"  detail:  {det}")
    print("  top:     " + "  ".join(f"{k}={v}" for k, v in c.most_common(14) if not k.startswith("*")))
