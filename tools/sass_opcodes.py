#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell-native paths (tcgen05 MMA, TMEM loads, TMA tensor and
bulk copies, tensor-core barriers) in the in-tree libvsgpu.so -> profiles/r2_sass_opcodes.txt.

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops.
"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTCBAR", "UTMALDG", "UTMALDG.2CTA", "UBLKCP", "UBLKPF", "SYNCS", "HMMA", "FFMA.SAT", "FMNMX3"]


def main() -> None:
    objs = sorted((ROOT / "vectorsearch_b200" / "csrc" / "build").glob("*.o"))
    if not objs:
        sys.exit("build the library first (python -c 'import __graft_entry__ as g; g.build()')")
    per = defaultdict(Counter)
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", str(o)], capture_output=True, text=True).stdout
        fn = None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(anonymous namespace\)::", "", name)
                name = re.sub(r"^void ", "", name)
                name = re.sub(r"vs::", "", name)
                fn = re.sub(r"\(.*$", "", name)
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
            if m and fn:
                op = m.group(1)
                base = op.split(".")[0]
                per[fn][base] += 1
                if base == "UTCHMMA" and ".2CTA" in op:
                    per[fn]["UTCHMMA.2CTA"] += 1
                if base == "UTMALDG" and ".2CTA" in op:
                    per[fn]["UTMALDG.2CTA"] += 1
                if op.startswith("FFMA.SAT"):
                    per[fn]["FFMA.SAT"] += 1
    print("# SASS opcode counts per kernel of vectorsearch_b200/libvsgpu.so (cuobjdump -sass of csrc/build/*.o, sm_100a)")
    print("# made by tools/sass_opcodes.py; kernels without any of these opcodes are listed at the end")
    print(f"{'kernel':<78}" + "".join(f"{w:>13}" for w in WANT))
    rest = []
    for fn in sorted(per):
        c = per[fn]
        if any(c.get(w, 0) for w in WANT[:8] + ["HMMA"]):
            print(f"{fn[:77]:<78}" + "".join(f"{c.get(w, 0):>13}" for w in WANT))
        else:
            rest.append(fn)
    print("\n# no tensor-core / TMA opcodes (plain LDG/STG + ALU kernels): " + ", ".join(sorted(set(r[:60] for r in rest))))


if __name__ == "__main__":
    main()
