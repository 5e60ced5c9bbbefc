"""SASS evidence: per-kernel counts of the Blackwell-specific instructions in libvmlp_b200.so (no GPU needed).
python tools/sass_counts.py > profiles/rNN_sass_tcgen05_tma_counts.csv"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "jittor-mlp_b200", "libvmlp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "SYNCS", "FFMA2", "FMUL2", "MUFU", "HMMA", "NANOSLEEP",
        "BAR.SYNC", "BAR.ARV", "UCGABAR", "ATOMS.CAST"]
print("# SASS evidence: cuobjdump -sass jittor-mlp_b200/libvmlp_b200.so, instruction counts per kernel that uses tcgen05 / TMA")
print("# (tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG/UTMASTG/UTMAPF, "
      "mbarrier -> SYNCS; HMMA = legacy mma.sync, must be 0; ATOMS.CAST = shared-memory CAS spin loops, must be 0 in the token kernels)")
print("kernel," + ",".join(KEYS) + ",sass_instructions")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
rows = []
for blk, name in zip(blocks, names):
    body = [l for l in blk.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    cnt = {k: sum(1 for l in body if re.search(r"\b" + re.escape(k), l)) for k in KEYS}
    if cnt["UTCHMMA"] or cnt["UTMALDG"] or cnt["UTMASTG"]:
        short = name.split("(CUtensorMap")[0].split("(const ")[0].replace("(int)", "")
        rows.append((short, cnt, len(body)))
for short, cnt, n in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    print('"%s",' % short + ",".join(str(cnt[k]) for k in KEYS) + f",{n}")
