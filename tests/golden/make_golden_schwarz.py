#!/usr/bin/env python3
"""Golden fixture of the additive Schwarz preconditioner: the LINE-BY-LINE restatement of schwarz.rs (oracle/schwarz_oracle.py,
CsrSchwarz: from_csr on the CSR image of the dense operator, ILU(0) per subdomain, substitutions, weighted combination) applied
to the committed ico1_ka0p5 system (80 unknowns), plus the left-preconditioned GMRES run of the oracle on it.

    python tests/golden/make_golden_schwarz.py      -> tests/golden/schwarz_ico1.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402
from oracle import schwarz_oracle as so  # noqa: E402

OUT = Path(__file__).resolve().parent
g = np.load(OUT / "ico1_ka0p5.npz")
A, b = g["A"], g["b"]
n = A.shape[0]
vals, cols, ptr = so.dense_to_csr(A)
rng = np.random.default_rng(2024)
r = rng.standard_normal(n) + 1j * rng.standard_normal(n)
rec = {"r": r}
for S in (4, 7):
    pre = so.CsrSchwarz(vals, cols, ptr, n, S, 0)
    rec[f"z_S{S}"] = pre.apply(r)
    x, info = orc.gmres_preconditioned_cb(lambda v: A @ v, pre.apply, n, b, max_iterations=100, restart=20, tolerance=1e-10)
    rec[f"x_S{S}"] = x
    rec[f"info_S{S}"] = np.array([info["iterations"], info["restarts"], int(info["converged"])])
    print(S, pre.stats(), info)
np.savez_compressed(OUT / "schwarz_ico1.npz", **rec)
