"""SASS opcode counts of the built library per kernel family (the Blackwell evidence, tracked under profiles/).

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "olmoasr_b200" / "csrc" / "liboasr_b200.so"
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "HMMA", "LDGSTS", "UTCBAR", "SYNCS", "REDG",
        "ATOMG", "MUFU.EX2", "USETMAXREG"]


def family(mangled: str) -> str:
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout
    m = re.search(r"(\w+_kernel)\b", name)
    return m.group(1) if m else "other"


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts = defaultdict(Counter)
    n_fn = Counter()
    fam = None
    for line in sass.splitlines():
        if "Function :" in line:
            fam = family(line.split("Function :")[1].strip())
            n_fn[fam] += 1
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m or fam is None:
            continue
        op = m.group(1)
        for c in COLS:
            if c == "UTCHMMA":
                hit = op.startswith("UTCHMMA") and ".2CTA" not in op
            elif c == "UTCHMMA.2CTA":
                hit = op.startswith("UTCHMMA") and ".2CTA" in op
            else:
                hit = op.startswith(c)
            if hit:
                counts[fam][c] += 1
    print(f"# SASS opcode counts of {LIB.relative_to(LIB.parents[2])} (cuobjdump -sass, sm_100a), per kernel family (all template instances summed)")
    print("# tcgen05.mma -> UTCHMMA[.2CTA]; tcgen05.ld/st -> LDTM/STTM; TMA -> UTMALDG/UTMASTG/UTMAREDG, L2 prefetch UTMAPF; mma.sync -> HMMA;")
    print("# cp.async -> LDGSTS; tcgen05.commit -> UTCBAR; mbarrier -> SYNCS; setmaxnreg -> USETMAXREG")
    print(f"{'kernel':44s}" + "".join(f"{c:>13s}" for c in COLS))
    for fam in sorted(counts, key=lambda f: -sum(counts[f].values())):
        print(f"{(fam + ' x' + str(n_fn[fam])):44s}" + "".join(f"{counts[fam][c]:13d}" for c in COLS))


if __name__ == "__main__":
    sys.exit(main())
