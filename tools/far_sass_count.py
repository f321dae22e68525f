#!/usr/bin/env python3
"""Static instruction mix of the far kernel's row loop (the body executed once per (row, column) pair and thread).

    python tools/far_sass_count.py [object-or-library] [kernel-name-substring ...]

Disassembles with cuobjdump -sass, takes for each matching kernel the smallest backward-branch span holding >= 90 % of
its DFMAs (the loop over collocation rows) and counts the instructions in it by class.  DP = DFMA + DMUL + DADD (+ DSETP): what the FP64
pipe executes per pair; 924 algorithmic flops per Tri3 pair = 462 DFMA equivalents."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True, check=True).stdout
    cur, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if cur:
        yield cur, body


def row_loop(body):
    """The smallest backward-branch span that holds at least 90 % of the kernel's DFMAs: the loop over collocation rows."""
    dfma = [a for a, i in body if classify(i) == "DFMA"]
    best = None
    for addr, ins in body:
        m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and sum(1 for a in dfma if tgt <= a <= addr) >= 0.9 * len(dfma):
                if best is None or addr - tgt < best[1] - best[0]:
                    best = (tgt, addr)
    return best


def classify(ins):
    op = ins.split()[0]
    if op.startswith("@"):
        op = ins.split()[1]
    base = op.split(".")[0]
    return base


def main():
    path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "math_audio_b200" / "lib" / "libbemb200.so"
    pats = sys.argv[2:] or ["far_kernel"]
    for name, body in kernels(path):
        if not any(p in name for p in pats):
            continue
        lp = row_loop(body)
        if not lp:
            continue
        cnt = collections.Counter(classify(i) for a, i in body if lp[0] <= a <= lp[1])
        dp = sum(cnt[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
        tot = sum(cnt.values())
        short = re.sub(r"^.*?(far_kernel\w*?I[^E]*E).*$", r"\1", name)
        print(f"{short}: row loop 0x{lp[0]:x}-0x{lp[1]:x}: {tot} instructions, DP {dp} (DFMA {cnt['DFMA']} DMUL {cnt['DMUL']} DADD {cnt['DADD']} DSETP {cnt['DSETP']}), "
              f"MUFU {cnt['MUFU']}, LDS {cnt['LDS']}, LDL {cnt['LDL']}, STL {cnt['STL']}, LDG {cnt['LDG']}, STG {cnt['STG']}, UBLKCP {sum(1 for a, i in body if 'UBLKCP' in i)} (whole kernel)")


if __name__ == "__main__":
    main()
