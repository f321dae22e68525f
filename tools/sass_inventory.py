#!/usr/bin/env python3
"""Per-kernel inventory of the built library: registers / shared memory / spills (cuobjdump --dump-resource-usage) and the
counts of the SASS mnemonics the design relies on (cuobjdump -sass).  Runs without a GPU.

    python tools/sass_inventory.py [math_audio_b200/lib/libbemb200.so] > profiles/<round>_sass_inventory.md
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "math_audio_b200" / "lib" / "libbemb200.so"

# column -> regular expression on the mnemonic (first token after the predicate)
COLS = OrderedDict([
    ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("DMMA", r"^DMMA"), ("MUFU64", r"^MUFU\.\w*64"),
    ("LDG128", r"^LDG\.E(\.\w+)*\.128"), ("STG128", r"^STG\.E(\.\w+)*\.128"), ("LDGSTS", r"^LDGSTS"), ("UBLKCP", r"^UBLKCP"),
    ("SYNCS", r"^SYNCS"), ("UCGABAR", r"^UCGABAR"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("ATOM/RED", r"^(ATOM|RED|ATOMG)"),
    ("LDL/STL", r"^(LDL|STL)"),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)  # template arguments stay, the parameter list goes


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(m.group(i)) for i in (2, 3, 4, 5))
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts, total, cur = {}, Counter(), None
    for line in sass.splitlines():
        f = re.match(r"\s+Function : (\S+)", line)
        if f:
            cur = f.group(1)
            counts[cur] = Counter()
            continue
        ins = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if ins and cur:
            total[cur] += 1
            for col, rx in COLS.items():
                if re.match(rx, ins.group(1)):
                    counts[cur][col] += 1
    names = demangle(sorted(counts))
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", res)))
    print(f"# SASS inventory of `{LIB.name}` ({LIB.stat().st_size} bytes; code objects: {', '.join(archs)}; {len(counts)} kernels)\n")
    print("Static counts per kernel (instructions in the binary, not executed counts).  `STACK` / `LDL/STL` > 0 = register spills or local arrays.\n")
    hdr = ["kernel", "REG", "STACK", "SHARED", "instr"] + list(COLS)
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for k in sorted(counts, key=lambda k: -total[k]):
        reg, stack, shared, _local = usage.get(k, (0, 0, 0, 0))
        row = [f"`{short(names[k])}`", reg, stack, shared, total[k]] + [counts[k][c] or "" for c in COLS]
        print("| " + " | ".join(str(v) for v in row) + " |")
    allc = Counter()
    for c in counts.values():
        allc.update(c)
    print("\nWhole library: " + ", ".join(f"{c} {allc[c]}" for c in COLS) + f"; tcgen05 (`UTC*MMA`) {len(re.findall(r'UTC[A-Z]*MMA', sass))} "
          "(none by design: every contraction on this path is FP64, tcgen05 has no f64 kind -- the block matvec uses `DMMA.8x8x4`).")


if __name__ == "__main__":
    main()
