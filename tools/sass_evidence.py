"""Instruction counts per kernel from `cuobjdump -sass` of the built library (mnemonics: /opt/skills/guides/B200_PROFILING.md).

    python tools/sass_evidence.py > profiles/rNN_sass_evidence.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "mlx-audio-swift_b200" / "lib" / "libb200audio.so"
COLS = [("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM/STTM (tcgen05.ld/st)", r"\b(LDTM|STTM)"),
        ("UTMALDG/UTMASTG (TMA tensor)", r"\bUTMA(LDG|STG)"), ("UBLKCP (bulk copy)", r"\bUBLKCP"),
        ("UCGABAR/cluster barrier", r"\bUCGABAR"), ("ACQBULK/PDL (griddepcontrol)", r"\bACQBULK|\bPREEXIT"),
        ("HMMA (legacy mma.sync)", r"\bHMMA"), ("MUFU.SIN", r"MUFU\.SIN"), ("MUFU.EX2", r"MUFU\.EX2")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = [0] * len(COLS)
            continue
        if cur is None:
            continue
        for i, (_, pat) in enumerate(COLS):
            if re.search(pat, line):
                counts[cur][i] += 1
    names = list(counts)
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, dm):
        d = re.sub(r"\(.*$", "", d)
        demangle[n] = d
    print(f"# SASS evidence: `cuobjdump -sass {LIB.relative_to(LIB.parents[2])}`, instruction counts per kernel")
    print("# (mnemonic table: /opt/skills/guides/B200_PROFILING.md).  Regenerate with `python tools/sass_evidence.py`.\n")
    print("| kernel | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    rows = sorted(((demangle[n], c) for n, c in counts.items()), key=lambda r: r[0])
    tot = [0] * len(COLS)
    for name, c in rows:
        for i, v in enumerate(c):
            tot[i] += v
        if any(c):
            print(f"| `{name}` | " + " | ".join(str(v) for v in c) + " |")
    print("| **total over all kernels** | " + " | ".join(str(v) for v in tot) + " |")
    quiet = [n for n, c in rows if not any(c)]
    print(f"\n{len(rows)} kernels in the library; {len(quiet)} use none of the above (plain LDG/STG/FFMA kernels).")


if __name__ == "__main__":
    sys.exit(main())
