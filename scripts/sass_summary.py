#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS mnemonics in libdewi_b200.so (sm_100a): tcgen05 MMAs (UTCHMMA; .2CTA =
cta_group::2), TMA loads (UTMALDG), TMEM loads (LDTM), tcgen05.commit barriers (UTCBAR).
Usage: python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "dewi-design-for-an-entropy-weighted-index-for-text-image-corpora_b200" / "libdewi_b200.so"
elf = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
print(f"# {LIB.name}: {LIB.stat().st_size} bytes; {elf.count('sm_100a')} cubin(s), all sm_100a; "
      f"sm_90-only mnemonics (HGMMA / wgmma): {len(re.findall(r'HGMMA', sass))}")
print("# kernel | UTCHMMA | UTCHMMA.2CTA | UTMALDG | LDTM | UTCBAR | SASS instructions")
rows, cur, it = [], None, iter(names)
for ln in sass.splitlines():
    if "Function :" in ln:
        cur = [next(it), 0, 0, 0, 0, 0, 0]
        rows.append(cur)
    elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        cur[6] += 1
        if "UTCHMMA.2CTA" in ln:
            cur[2] += 1
        elif "UTCHMMA" in ln:
            cur[1] += 1
        cur[3] += "UTMALDG" in ln
        cur[4] += "LDTM" in ln
        cur[5] += "UTCBAR" in ln
for r in sorted(rows):
    name = re.sub(r"\(anonymous namespace\)::", "", r[0])
    name = re.sub(r"\(.*", "", name)
    print(f"{name:70s} {r[1]:4d} {r[2]:4d} {r[3]:4d} {r[4]:4d} {r[5]:4d} {r[6]:6d}")
tot = [sum(r[i] for r in rows) for i in range(1, 7)]
print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':70s} {tot[0]:4d} {tot[1]:4d} {tot[2]:4d} {tot[3]:4d} {tot[4]:4d} {tot[5]:6d}")
