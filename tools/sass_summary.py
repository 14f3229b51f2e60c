#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from the built library.

    python tools/sass_summary.py [path/to/libafa_sm100.so] > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = tensor-map TMA, UBLKCP = 1-D bulk TMA,
FFMA2 / FMUL2 = packed f32x2 math, HMMA = legacy mma.sync, MUFU = special-function unit (B200_PROFILING.md, "What proves a
Blackwell-native kernel").  Runs on the CPU build box (cuobjdump, no GPU)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200", "afa_b200", "libafa_sm100.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "HMMA", "MUFU", "LDS", "STS", "LDG", "STG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn:
                counts[name][mn] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, REPO)} ({os.path.getsize(lib)} bytes), cuobjdump -sass, sm_100a")
print("# " + " ".join(f"{m:>8s}" for m in ["instrs"] + MNEMONICS) + "  kernel")
agg = collections.Counter()
for k, c in counts.items():
    if c["_total"] == 0:
        continue
    print("  " + " ".join(f"{c.get(m, 0):8d}" for m in ["_total"] + MNEMONICS) + "  " + k)
    agg.update(c)
print("  " + " ".join(f"{agg.get(m, 0):8d}" for m in ["_total"] + MNEMONICS) + "  TOTAL (" + str(len(counts)) + " kernels)")
