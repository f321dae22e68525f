#!/usr/bin/env python3
"""Extract the quadrature tables the independent restatement needs straight from the reference's
`math-bem/src/core/integration/gauss.rs` (static arrays GL<n>_X / GL<n>_W and GAUCORWEI_TR<n>) into
`oracle/independent/tables.json`.  Independent of tools/gen_quad_tables.py (which feeds the C++ oracle and the
CUDA kernels): different parser, different output, so a transcription slip in one shows up against the other.

    python oracle/independent/extract_tables.py [/root/reference]
"""
import json
import re
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
src = (ref / "math-bem/src/core/integration/gauss.rs").read_text()
out = {"gl_x": {}, "gl_w": {}, "tri": {}}
for m in re.finditer(r"static\s+(GL(\d+)_([XW]))\s*:\s*\[f64;\s*(\d+)\]\s*=\s*\[(.*?)\];", src, re.S):
    _, order, kind, cnt, body = m.groups()
    vals = [float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?", body)]
    assert len(vals) == int(cnt) == int(order), (order, kind, len(vals))
    out["gl_x" if kind == "X" else "gl_w"][order] = vals
for m in re.finditer(r"static\s+GAUCORWEI_TR(\d+)\s*:\s*\[\[f64;\s*3\];\s*(\d+)\]\s*=\s*\[(.*?)\];", src, re.S):
    npts, cnt, body = m.groups()
    vals = [float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?", body)]
    assert len(vals) == 3 * int(cnt) and int(cnt) == int(npts)
    out["tri"][npts] = [vals[3 * i: 3 * i + 3] for i in range(int(cnt))]
assert set(out["gl_x"]) == set(out["gl_w"]) >= {"1", "2", "3", "4", "5", "6", "7", "8"}
assert set(out["tri"]) >= {"1", "4", "7", "13"}
dst = Path(__file__).resolve().parent / "tables.json"
dst.write_text(json.dumps(out, indent=0))
print("wrote", dst, {k: sorted(v, key=int) for k, v in out.items()})
