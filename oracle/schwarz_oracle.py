"""CPU restatement of math-solvers/src/preconditioners/schwarz.rs (AdditiveSchwarzPreconditioner) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.

Two restatements of the same algorithm:
  * ``CsrSchwarz`` -- the reference line by line on CSR arrays (from_csr :66-125, build_adjacency :161-174,
    extend_partition :177-203, build_subdomain :205-251, ilu_factorize :253-345, Subdomain::solve :348-380,
    apply_sequential :399-417), pure Python loops: small cases only;
  * ``DenseSchwarz`` -- the same operations on dense blocks with numpy (what the reference computes when the CSR matrix is
    the image of a dense operator: ILU(0) of a full pattern IS the LU factorisation without pivoting), every entry updated
    in the reference's order.  tests/test_schwarz_cpu.py checks the two against each other and against numpy.linalg.

parity unpinned by reference vectors: schwarz.rs holds no numeric golden vector; its own tests (:462-572: a 1-D Laplacian is
reduced in norm, sizes / stats, "solves better than identity") are restated in tests/test_schwarz_cpu.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

TINY = 1e-30


def _norm(z: complex) -> float:  # ComplexField::norm (traits.rs:93-95)
    return float(np.sqrt(z.real * z.real + z.imag * z.imag))


def _inv(z: complex) -> complex:  # ComplexField::inv: conj / norm_sqr (traits.rs:150-153)
    ns = z.real * z.real + z.imag * z.imag
    return complex(z.real / ns, -z.imag / ns)


def contiguous_partition(n: int, num_subdomains: int) -> List[np.ndarray]:
    """schwarz.rs:67-83."""
    S = min(max(int(num_subdomains), 1), n)
    base, rem = divmod(n, S)
    out, start = [], 0
    for i in range(S):
        size = base + (1 if i < rem else 0)
        out.append(np.arange(start, start + size, dtype=np.int64))
        start += size
    return out


def build_adjacency(row_ptrs, col_indices, n: int) -> List[List[int]]:
    """schwarz.rs:161-174."""
    adj: List[List[int]] = [[] for _ in range(n)]
    for i in range(n):
        for idx in range(row_ptrs[i], row_ptrs[i + 1]):
            j = int(col_indices[idx])
            if i != j:
                adj[i].append(j)
    return adj


def extend_partition(partition, adjacency, overlap: int, n: int) -> np.ndarray:
    """schwarz.rs:177-203."""
    inside = [False] * n
    for i in partition:
        inside[int(i)] = True
    frontier = [int(i) for i in partition]
    for _ in range(overlap):
        new = []
        for i in frontier:
            for nb in adjacency[i]:
                if not inside[nb]:
                    inside[nb] = True
                    new.append(nb)
        frontier = new
    return np.array([i for i in range(n) if inside[i]], dtype=np.int64)


def dense_to_csr(A: np.ndarray, pattern: Optional[np.ndarray] = None):
    """CSR image of a dense matrix (all entries, or those of a boolean pattern), columns ascending."""
    n = A.shape[0]
    vals, cols, ptr = [], [], [0]
    for i in range(n):
        js = range(A.shape[1]) if pattern is None else np.nonzero(pattern[i])[0]
        for j in js:
            vals.append(complex(A[i, j]))
            cols.append(int(j))
        ptr.append(len(vals))
    return vals, cols, ptr


class CsrSchwarz:
    """The reference, line by line (pure Python; O(n^4) on dense patterns like the reference itself: n <= ~40)."""

    def __init__(self, values, col_indices, row_ptrs, n: int, num_subdomains: int, overlap: int):
        self.n = n
        parts = contiguous_partition(n, num_subdomains)
        adjacency = build_adjacency(row_ptrs, col_indices, n)
        ext = [extend_partition(p, adjacency, overlap, n) for p in parts]
        count = [0] * n
        for p in ext:
            for i in p:
                count[int(i)] += 1
        self.weights = [1.0 / c if c > 0 else 1.0 for c in count]
        self.subdomains = [self._build_subdomain(values, col_indices, row_ptrs, [int(i) for i in p]) for p in ext]

    @staticmethod
    def _build_subdomain(values, col_indices, row_ptrs, gidx):
        g2l = {g: l for l, g in enumerate(gidx)}
        lv, lc, lp = [], [], [0]
        for g in gidx:
            for idx in range(row_ptrs[g], row_ptrs[g + 1]):
                lcidx = g2l.get(int(col_indices[idx]))
                if lcidx is not None:
                    lv.append(complex(values[idx]))
                    lc.append(lcidx)
            lp.append(len(lv))
        n = len(gidx)
        v = list(lv)
        for i in range(n):  # ilu_factorize, schwarz.rs:271-305
            for idx in range(lp[i], lp[i + 1]):
                k = lc[idx]
                if k >= i:
                    break
                u_kk = 0j
                for k_idx in range(lp[k], lp[k + 1]):
                    if lc[k_idx] == k:
                        u_kk = v[k_idx]
                        break
                if _norm(u_kk) < TINY:
                    continue
                l_ik = v[idx] * _inv(u_kk)
                v[idx] = l_ik
                for j_idx in range(lp[i], lp[i + 1]):
                    j = lc[j_idx]
                    if j <= k:
                        continue
                    for k_j_idx in range(lp[k], lp[k + 1]):
                        if lc[k_j_idx] == j:
                            v[j_idx] = v[j_idx] - l_ik * v[k_j_idx]
                            break
        L = [[] for _ in range(n)]
        U = [[] for _ in range(n)]
        u_diag = [1 + 0j] * n
        for i in range(n):
            for idx in range(lp[i], lp[i + 1]):
                j = lc[idx]
                if j < i:
                    L[i].append((j, v[idx]))
                else:
                    U[i].append((j, v[idx]))
                    if j == i:
                        u_diag[i] = v[idx]
        return dict(gidx=gidx, L=L, U=U, u_diag=u_diag)

    @staticmethod
    def _solve(sd, rhs):  # schwarz.rs:348-380
        n = len(sd["gidx"])
        y = list(rhs)
        for i in range(n):
            for j, l in sd["L"][i]:
                y[i] = y[i] - l * y[j]
        x = y
        for i in range(n - 1, -1, -1):
            for j, u in sd["U"][i]:
                if j > i:
                    x[i] = x[i] - u * x[j]
            if _norm(sd["u_diag"][i]) > TINY:
                x[i] = x[i] * _inv(sd["u_diag"][i])
        return x

    def apply(self, r):  # apply_sequential, schwarz.rs:399-417
        out = [0j] * self.n
        for sd in self.subdomains:
            sol = self._solve(sd, [complex(r[g]) for g in sd["gidx"]])
            for l, g in enumerate(sd["gidx"]):
                out[g] += sol[l] * self.weights[g]
        return np.array(out, dtype=np.complex128)

    def stats(self):  # schwarz.rs:135-158
        sizes = [len(sd["gidx"]) for sd in self.subdomains]
        return len(sizes), min(sizes), max(sizes), sum(sizes) / len(sizes)


def lu_nopivot(block: np.ndarray) -> np.ndarray:
    """ilu_factorize (schwarz.rs:271-305) on a full pattern: row i, pivots k < i in ascending order,
    l_ik = a_ik * inv(u_kk), then a_ij -= l_ik * a_kj for j > k.  Returns L (strict lower, unit diagonal implied) and U in
    one array."""
    F = np.array(block, dtype=np.complex128)
    n = F.shape[0]
    for i in range(n):
        for k in range(i):
            u_kk = F[k, k]
            if _norm(u_kk) < TINY:
                continue
            l_ik = F[i, k] * _inv(u_kk)
            F[i, k] = l_ik
            F[i, k + 1:] = F[i, k + 1:] - l_ik * F[k, k + 1:]
    return F


def lu_solve_nopivot(F: np.ndarray, rhs: np.ndarray) -> np.ndarray:
    """Subdomain::solve (schwarz.rs:348-380); every y_i / x_i accumulates its terms in ascending j like the reference
    (np.cumsum adds strictly left to right)."""
    n = F.shape[0]
    y = np.array(rhs, dtype=np.complex128)
    for i in range(1, n):
        terms = np.empty(i + 1, dtype=np.complex128)
        terms[0] = y[i]
        terms[1:] = -(F[i, :i] * y[:i])
        y[i] = np.cumsum(terms)[-1]
    x = y
    for i in range(n - 1, -1, -1):
        if i + 1 < n:
            terms = np.empty(n - i, dtype=np.complex128)
            terms[0] = x[i]
            terms[1:] = -(F[i, i + 1:] * x[i + 1:])
            x[i] = np.cumsum(terms)[-1]
        if _norm(F[i, i]) > TINY:
            x[i] = x[i] * _inv(F[i, i])
    return x


class DenseSchwarz:
    """AdditiveSchwarzPreconditioner of a dense operator: ``subdomains`` (index arrays; the contiguous partition of
    ``num_subdomains`` blocks by default), local LU without pivoting, weights 1 / multiplicity."""

    def __init__(self, A: np.ndarray, num_subdomains: int = 0, subdomains: Optional[Sequence[np.ndarray]] = None):
        A = np.asarray(A)
        self.n = A.shape[0]
        subs = subdomains if subdomains is not None else contiguous_partition(self.n, num_subdomains)
        self.subs = [np.asarray(s, dtype=np.int64) for s in subs if len(s)]
        count = np.zeros(self.n, dtype=np.int64)
        for s in self.subs:
            count[s] += 1
        self.weights = np.where(count > 0, 1.0 / np.maximum(count, 1), 1.0)
        self.factors = [lu_nopivot(A[np.ix_(s, s)]) for s in self.subs]

    def apply(self, r: np.ndarray) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.complex128)
        for s, F in zip(self.subs, self.factors):
            out[s] += lu_solve_nopivot(F, r[s]) * self.weights[s]
        return out

    def stats(self):
        sizes = [len(s) for s in self.subs]
        return len(sizes), min(sizes), max(sizes), sum(sizes) / len(sizes)
