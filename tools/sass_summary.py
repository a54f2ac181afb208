"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md): tcgen05.mma = UTC*MMA,
tcgen05.ld / st = LDTM / STTM, TMA = UTMALDG / UTMASTG / UBLKCP / UBLKRED, tcgen05.commit = UTCBAR, legacy tensor path = HMMA.
usage: python tools/sass_summary.py [library.so] > profiles/sass_summary.txt"""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "scarf_b200/csrc/libscarf_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = [("UTCxMMA", r"UTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG/STG", r"UTMA(LDG|STG)"),
        ("UBLKCP/RED", r"UBLK(CP|RED)"), ("UTCBAR", r"UTCBAR"), ("HMMA", r"\bHMMA"), ("DMMA", r"\bDMMA"), ("DFMA", r"\bDFMA"),
        ("FMNMX3", r"\bFMNMX3")]
rows, name, cnt = [], None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, cnt))
        name, cnt = m.group(1), [0] * len(pats)
        continue
    if name and "/*" in line:
        for i, (_, p) in enumerate(pats):
            if re.search(p, line):
                cnt[i] += 1
if name:
    rows.append((name, cnt))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {so}: static instruction counts per kernel (kernels that use a tensor / TMA instruction, plus "
      "the FP64 kernels of the eigensolver)")
print(f"{'kernel':58s} " + " ".join(f"{p[0]:>11s}" for p in pats))
for (raw, c), dn in zip(rows, names):
    dn = dn.replace("void ", "").replace("(anonymous namespace)::", "")
    dn = re.sub(r"\((?!anonymous).*", "", dn)
    if sum(c[:7]) > 0 or dn.startswith("eig_") or dn.startswith("jacobi"):
        print(f"{dn[:58]:58s} " + " ".join(f"{x:11d}" for x in c))
