"""SECOND, INDEPENDENT CPU restatement of the hot path in numpy -- TEST INFRASTRUCTURE ONLY.

Written from the reference's Rust sources alone (math-bem/src/core/assembly/tbem.rs:96-345,
core/integration/regular.rs:33-260, core/integration/singular.rs:48-82,123-465,497-745,
core/integration/gauss.rs:15-105, core/types.rs:64-70, core/mesh/element.rs:124-131,
math-solvers/src/iterative/gmres.rs:105-277,589-621, blas_helpers.rs:21-73; for the neighbours of the path also
core/incident.rs:93-342 and core/postprocess/pressure.rs:81-259,438-478) WITHOUT consulting
oracle/bem_oracle.cpp: its purpose is to pin the C++ oracle (the reference cannot be compiled in this
image and holds no numeric golden vectors for this path -- SURVEY.md 8c).  Two restatements written
separately from the same source that agree to 1e-13 on every matrix entry and on every GMRES iteration
count are the strongest pin available without a Rust toolchain.

`assemble_rows_general` walks tbem.rs:126-345 for every boundary-condition class (free terms, matrix entry per field BC, the
integrators' right-hand-side terms with the unscaled beta).

Different structure on purpose: the un-subdivided pairs of a row are evaluated as one vectorised numpy
expression over (field element, quadrature point); only subdivided pairs and the self term walk the
reference's loops.  Tables come from oracle/independent/tables.json (extract_tables.py parses gauss.rs).
Only the product-independent mesh arrays (nodes, connectivity, centres, normals, areas) are shared inputs.
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np

_T = json.loads((Path(__file__).resolve().parent / "tables.json").read_text())

TRI, QUAD = 3, 4
CSI6 = [0.0, 1.0, 0.0, 0.5, 0.5, 0.0]
ETA6 = [0.0, 0.0, 1.0, 0.0, 0.5, 0.5]
CSI8 = [1.0, -1.0, -1.0, 1.0, 0.0, -1.0, 0.0, 1.0]
ETA8 = [1.0, 1.0, -1.0, -1.0, 1.0, 0.0, -1.0, 0.0]


# ---- gauss.rs -----------------------------------------------------------------------------------
def gauss_legendre(order: int):
    """gauss.rs:15-60: tabulated orders 1-8, 10, 12, 16, 20; anything else rounds UP to the next table."""
    assert 1 <= order <= 80
    if str(order) not in _T["gl_x"]:
        for cand in (2, 4, 6, 8, 12, 16, 20):
            if order <= cand:
                order = cand
                break
        else:
            order = 20
    return np.array(_T["gl_x"][str(order)]), np.array(_T["gl_w"][str(order)])


def triangle_quadrature(order: int) -> np.ndarray:
    """gauss.rs:67-89: orders 1, 2, 3 -> TR1, TR4, TR7, everything else TR13; weights x 0.5."""
    key = {1: "1", 2: "4", 3: "7"}.get(order, "13")
    t = np.array(_T["tri"][key])
    t[:, 2] = t[:, 2] * 0.5
    return t


def quad_quadrature(order: int) -> np.ndarray:
    x, w = gauss_legendre(order)
    return np.array([(xi, eta, w[i] * w[j]) for i, xi in enumerate(x) for j, eta in enumerate(x)])


def element_quadrature(etype: int, order: int) -> np.ndarray:
    return triangle_quadrature(order) if etype == TRI else quad_quadrature(order)


# ---- shape functions (regular.rs:193-260, singular.rs:398-465, :696-721) -------------------------
def shape(etype: int, s, t):
    """-> (N, dN/ds, dN/dt), each with a leading axis over the element's nodes; s, t scalars or arrays."""
    s = np.asarray(s, dtype=float)
    t = np.asarray(t, dtype=float)
    one = np.ones_like(s)
    if etype == TRI:
        return (np.stack([1.0 - s - t, s, t]), np.stack([-one, one, 0 * one]), np.stack([-one, 0 * one, one]))
    s1, s2, t1, t2 = 0.25 * (s + 1.0), 0.25 * (s - 1.0), t + 1.0, t - 1.0
    n = np.stack([s1 * t1, -s2 * t1, s2 * t2, -s1 * t2])
    ds = np.stack([0.25 * (t + 1.0), -0.25 * (t + 1.0), 0.25 * (t - 1.0), -0.25 * (t - 1.0)])
    dt = np.stack([0.25 * (s + 1.0), 0.25 * (1.0 - s), 0.25 * (s - 1.0), -0.25 * (s + 1.0)])
    return n, ds, dt


def geometry_at(coords: np.ndarray, etype: int, s, t):
    """compute_parameters: coords (nodes, 3); s, t arrays of any shape P -> N (nodes, P), jac (P), normal (P, 3), y (P, 3)."""
    n, ds, dt = shape(etype, s, t)
    y = np.tensordot(n, coords, axes=(0, 0))
    xs = np.tensordot(ds, coords, axes=(0, 0))
    xt = np.tensordot(dt, coords, axes=(0, 0))
    nrm = np.cross(xs, xt)
    jac = np.sqrt(np.sum(nrm * nrm, axis=-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        unit = np.where(jac[..., None] > 1e-15, nrm / jac[..., None], 0.0)
    return n, jac, unit, y


# ---- kernels at a cloud of points (regular.rs:106-154) ----------------------------------------------
def kernels(x, nx, y, ny, wga, k, harmonic=1.0):
    """zg, zhh, zht, ze at points y (P, 3) with unit normals ny (P, 3) and weights wga (P); points closer than 1e-15 are
    skipped (weight zero), as regular.rs:120-122."""
    wav = harmonic * k
    d = y - x
    r = np.sqrt(np.sum(d * d, axis=-1))
    ok = r >= 1e-15
    rs = np.where(ok, r, 1.0)
    u = d / rs[..., None]
    re1 = wav * rs
    re2 = wga / (4.0 * math.pi * rs)
    zg = (np.cos(re1) * re2 + 1j * (np.sin(re1) * re2)) * ok
    base = zg * (-1.0 / rs + 1j * wav)
    h1 = np.sum(u * ny, axis=-1)
    h2 = -np.sum(u * nx, axis=-1)
    rq = h1 * h2
    nn = np.sum(ny * nx, axis=-1)
    dq = rs * rs
    ze = zg * (((3.0 / dq - k * k) * rq + nn / dq) + 1j * (-wav / rs * (3.0 * rq + nn)))
    return zg, base * h1, base * h2, ze


# ---- adaptive subdivision (singular.rs:497-693) -------------------------------------------------------
def _gauss_order(disfac: float, gmin=4, gmax=7, acc=0.0005) -> int:
    for order in range(gmin, gmax + 1):
        q = disfac / (2.0 * order + 1.0)
        if q ** (2 * order + 1) < acc and q ** (2 * order + 2) < acc and q ** (2 * order + 3) < acc:
            return order
    return gmax


def generate_subelements(x: np.ndarray, coords: np.ndarray, etype: int, area: float):
    """-> list of (xi_centre, eta_centre, factor, gauss_order, tri_vertices | None).  Level l has edge factor 2^-l; a
    candidate is accepted when dist(centre, x) / sqrt(area * factor^2) >= 3; more than 15 splits in a level abandon the
    rest of that level; 60 working slots; at most 110 outputs."""
    nv = etype
    out = []
    cur = [(list((CSI8 if etype == QUAD else CSI6)[:nv]), list((ETA8 if etype == QUAD else ETA6)[:nv]))]
    faclin = 2.0
    while True:
        faclin *= 0.5
        arels = area * faclin * faclin
        nxt = []
        ndie = 0
        for xi, et in cur:
            sc = sum(xi[:nv]) / nv
            tc = sum(et[:nv]) / nv
            n, _, _ = shape(etype, sc, tc)
            p = n @ coords
            diff = p - x
            ratdis = math.sqrt(float(diff @ diff)) / math.sqrt(arels)
            if ratdis < 3.0:
                ndie += 1
                if ndie > 15:
                    break
                xs = list(xi) + [(xi[j] + xi[(j + 1) % nv]) / 2.0 for j in range(nv)]
                es = list(et) + [(et[j] + et[(j + 1) % nv]) / 2.0 for j in range(nv)]
                for j in range(nv):
                    j1 = j + nv
                    j2 = j1 - 1 if j1 > nv else j1 + nv - 1
                    if etype == QUAD:
                        nxt.append(([xs[j], xs[j1], sc, xs[j2]], [es[j], es[j1], tc, es[j2]]))
                    else:
                        nxt.append(([xs[j], xs[j1], xs[j2]], [es[j], es[j1], es[j2]]))
                if etype == TRI:
                    nxt.append(([xs[3], xs[4], xs[5]], [es[3], es[4], es[5]]))
            else:
                if etype == QUAD:
                    rec = (sum(xi) / 4.0, sum(et) / 4.0, faclin, _gauss_order(0.5 / ratdis), None)
                else:
                    rec = ((xi[0] + xi[1] + xi[2]) / 3.0, (et[0] + et[1] + et[2]) / 3.0, faclin, _gauss_order(0.5 / ratdis),
                           [(xi[0], et[0]), (xi[1], et[1]), (xi[2], et[2])])
                out.append(rec)
                if len(out) >= 110:
                    return out
        if ndie == 0:
            break
        # the reference keeps its 60-slot arrays between levels: after a `break` at the 16th split only the first
        # 15 * 4 children exist, which is exactly what `nxt` holds here
        cur = nxt[:60]
    return out


def unscaled_beta(k: float, harmonic: float, tau: float) -> complex:
    """PhysicsParams::burton_miller_beta (types.rs:64-70): i * harmonic_factor / k outside, 0 inside -- what the integrators
    use for their right-hand-side terms whatever beta the assembly was called with (regular.rs:166-168)."""
    return complex(0.0, harmonic / k) if tau > 0.0 else 0j


def regular_pair(x, nx, coords, etype, area, k, harmonic=1.0, bc=None, bc_type=0, tau=1.0, gamma=1.0):
    """regular_integration for one (source, field element) pair -> (G, H, Ht, E) and, when per-node boundary values `bc` are
    given (compute_rhs), a fifth entry: the right-hand-side contribution of regular.rs:157-177 (the values are interpolated
    with the first len(bc) shape functions only; velocity: (zg gamma tau + zht beta0) v, pressure: -(zhh gamma tau + ze beta0) p)."""
    acc = np.zeros(5 if bc is not None else 4, dtype=complex)
    b0 = unscaled_beta(k, harmonic, tau)
    for xc, ec, fac, order, tv in generate_subelements(x, coords, etype, area):
        q = element_quadrature(etype, order)
        cs, et, w = q[:, 0], q[:, 1], q[:, 2]
        if abs(abs(fac) - 1.0) < 1e-10:
            s, t, w2 = cs, et, w
        elif tv is not None:
            l0 = 1.0 - cs - et
            s = tv[0][0] * l0 + tv[1][0] * cs + tv[2][0] * et
            t = tv[0][1] * l0 + tv[1][1] * cs + tv[2][1] * et
            det = abs((tv[1][0] - tv[0][0]) * (tv[2][1] - tv[0][1]) - (tv[2][0] - tv[0][0]) * (tv[1][1] - tv[0][1]))
            w2 = w * det
        else:
            s, t, w2 = xc + cs * fac, ec + et * fac, w * (fac * fac)
        nsh, jac, ny, y = geometry_at(coords, etype, s, t)
        zg, zhh, zht, ze = kernels(x, nx, y, ny, w2 * jac, k, harmonic)
        terms = [(zg, 0), (zhh, 1), (zht, 2), (ze, 3)]
        if bc is not None:
            zb = sum(bc[i] * nsh[i] for i in range(min(len(bc), nsh.shape[0])))
            if bc_type == 0:
                terms.append(((zg * gamma * tau + zht * b0) * zb, 4))
            elif bc_type == 1:
                terms.append((-((zhh * gamma * tau + ze * b0) * zb), 4))
        # the reference accumulates point by point: a left-to-right running sum, not numpy's pairwise sum
        for z, i in terms:
            a = acc[i]
            for v in z:
                a = a + v
            acc[i] = a
    return acc


# ---- self term (singular.rs:48-82, 123-394, 730-745) -----------------------------------------------------
def for_ka(ka: float):
    if ka < 0.3:
        return 3, 4, 4, 2
    if ka < 1.0:
        return 4, 5, 6, 2
    if ka < 2.0:
        return 5, 6, 8, 3
    return 6, 7, 10, 4


def element_size(coords: np.ndarray, etype: int) -> float:
    tot = 0.0
    for i in range(etype):
        d = coords[(i + 1) % etype] - coords[i]
        tot += math.sqrt(float(d @ d))
    return tot / etype


def singular_self(x, nx, coords, etype, k, harmonic=1.0, bc=None, bc_type=0, tau=1.0, gamma=1.0):
    """-> (G, H, Ht, E) and with `bc` a fifth entry, the right-hand-side contribution: velocity values are interpolated at every
    Duffy point (singular.rs:359-375), pressure values enter as their mean times -(H gamma tau + E beta0) at the end (:381-391)."""
    wav = harmonic * k
    b0 = unscaled_beta(k, harmonic, tau)
    R = 0j
    ngpo1, ngausin, nsec1, nsec2 = for_ka(k * element_size(coords, etype))
    gx, gw = gauss_legendre(ngpo1)
    sx, sw = gauss_legendre(ngausin)
    G = H = HT = E = 0j
    nn = etype
    for ieg in range(nn):
        ig1 = (ieg + 1) % nn
        ig2 = ieg + nn
        dp = coords[ig1] - coords[ieg]
        leneg = math.sqrt(float(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]))
        tdir = dp / leneg
        wscale = leneg / (2.0 * nsec1)
        delsec = 2.0 / nsec1
        secmid = -1.0 - delsec / 2.0
        zre = 0j
        for _ in range(nsec1):
            secmid += delsec
            for ig in range(ngpo1):
                sga = secmid + gx[ig] / nsec1
                wga = gw[ig] * wscale
                p = coords[ieg] + dp * (sga + 1.0) / 2.0
                d = p - x
                r = math.sqrt(float(d @ d))
                if r < 1e-15:
                    continue
                u = d / r
                zg = complex(math.cos(wav * r) / (4.0 * math.pi * r), math.sin(wav * r) / (4.0 * math.pi * r))
                grad = zg * complex(-1.0 / r, wav) * u
                cr = np.array([grad[1] * tdir[2] - grad[2] * tdir[1], grad[2] * tdir[0] - grad[0] * tdir[2],
                               grad[0] * tdir[1] - grad[1] * tdir[0]])
                zre += (cr[0] * nx[0] + cr[1] * nx[1] + cr[2] * nx[2]) * wga
        E += zre
        cs_, et_ = (CSI6, ETA6) if etype == TRI else (CSI8, ETA8)
        for isec in range(nsec2):
            aresub = (1.0 / 24.0 / nsec2) if etype == TRI else (0.25 / nsec2)
            s0, t0 = (1.0 / 3.0, 1.0 / 3.0) if etype == TRI else (0.0, 0.0)
            if isec == 0:
                s1, s2, t1, t2 = cs_[ieg], cs_[ig2], et_[ieg], et_[ig2]
            else:  # every later section repeats the SECOND sub-triangle (singular.rs:268-278)
                s1, s2, t1, t2 = cs_[ig2], cs_[ig1], et_[ig2], et_[ig1]
            for i, sga in enumerate(sx):
                for j, tga in enumerate(sx):
                    wg = sw[i] * sw[j]
                    sgg = 0.5 * (1.0 - sga) * s0 + 0.25 * (1.0 + sga) * ((1.0 - tga) * s1 + (1.0 + tga) * s2)
                    tgg = 0.5 * (1.0 - sga) * t0 + 0.25 * (1.0 + sga) * ((1.0 - tga) * t1 + (1.0 + tga) * t2)
                    nsh, jac, ny, y = geometry_at(coords, etype, sgg, tgg)
                    w = wg * (1.0 + sga) * aresub * float(jac)
                    d = y - x
                    r = math.sqrt(float(d @ d))
                    if r < 1e-15:
                        continue
                    u = d / r
                    re2 = w / (4.0 * math.pi * r)
                    zg = complex(math.cos(wav * r) * re2, math.sin(wav * r) * re2)
                    base = zg * complex(-1.0 / r, wav)
                    zht = base * float(-(u @ nx))
                    G += zg
                    H += base * float(u @ ny)
                    HT += zht
                    E += zg * (k * k) * float(nx @ ny)
                    if bc is not None and bc_type == 0:
                        zb = sum(bc[i] * float(nsh[i]) for i in range(min(len(bc), nsh.shape[0])))
                        R += (zg * gamma * tau + zht * b0) * zb
    if bc is None:
        return np.array([G, H, HT, E])
    if bc_type == 1:
        R = -(H * gamma * tau + E * b0) * (sum(bc) / len(bc))
    return np.array([G, H, HT, E, R])


# ---- tbem.rs: rigid (zero-velocity) elements only -----------------------------------------------------------
def dg_dn_sign(centers: np.ndarray, k: float) -> float:
    n = min(len(centers), 100)
    avg = sum(math.sqrt(float(c @ c)) for c in centers[:n]) / n if n else 0.0
    return 1.0 if k * avg < 0.5 else -1.0


def assemble_rows(nodes, conn, etype, centers, normals, areas, k, beta, rows, harmonic=1.0, tau=1.0, gamma=1.0):
    """Rows `rows` of build_tbem_system_with_beta for an all-rigid mesh (Velocity([0]) on every element, dof = element
    index): A[i, j] = sign gamma tau H_ij + beta E_ij, A[i, i] += -gamma / 2 (tbem.rs:284-290, 323-331)."""
    n = len(conn)
    sign = dg_dn_sign(centers, k)
    out = np.zeros((len(rows), n), dtype=complex)
    coords_all = [nodes[conn[j][: etype[j]]] for j in range(n)]
    # level-0 decision for the whole row at once: the element centre through the shape functions at the mean of the
    # vertices' local coordinates, exactly as generate_subelements forms it
    def _centre(c):
        et = len(c)
        cs_, et_ = (CSI8, ETA8) if et == QUAD else (CSI6, ETA6)
        return shape(et, sum(cs_[:et]) / et, sum(et_[:et]) / et)[0] @ c
    cen = np.stack([_centre(c) for c in coords_all])
    for oi, i in enumerate(rows):
        x, nx = centers[i], normals[i]
        dist = np.sqrt(np.sum((cen - x) ** 2, axis=1))
        far = dist / np.sqrt(areas) >= 3.0  # ratio test of level 0 (faclin = 1)
        far[i] = False
        for et in (TRI, QUAD):
            idx = np.nonzero(far & (etype == et))[0]
            if not len(idx):
                continue
            q = element_quadrature(et, 4)  # accepted at level 0 => ratio >= 3 => disfac <= 1/6 => order 4
            assert all(_gauss_order(0.5 / (dist[j] / math.sqrt(areas[j]))) == 4 for j in idx[:3])
            C = np.stack([coords_all[j] for j in idx])                      # (E, nodes, 3)
            nsh, ds, dt = shape(et, q[:, 0], q[:, 1])                       # (nodes, Q)
            y = np.einsum("nq,enc->eqc", nsh, C)
            xs = np.einsum("nq,enc->eqc", ds, C)
            xt = np.einsum("nq,enc->eqc", dt, C)
            nrm = np.cross(xs, xt)
            jac = np.sqrt(np.sum(nrm * nrm, axis=-1))
            ny = nrm / jac[..., None]
            zg, zhh, zht, ze = kernels(x, nx, y, ny, q[:, 2][None, :] * jac, k, harmonic)
            Hs = np.zeros(len(idx), dtype=complex)
            Es = np.zeros(len(idx), dtype=complex)
            for qq in range(q.shape[0]):  # running sums in quadrature order, as the reference
                Hs = Hs + zhh[:, qq]
                Es = Es + ze[:, qq]
            out[oi, idx] = (Hs * sign) * gamma * tau + Es * beta
        for j in np.nonzero(~far)[0]:
            if j == i:
                g, h, ht, e = singular_self(x, nx, coords_all[j], int(etype[j]), k, harmonic)
            else:
                g, h, ht, e = regular_pair(x, nx, coords_all[j], int(etype[j]), float(areas[j]), k, harmonic)
            out[oi, j] = (h * sign) * gamma * tau + e * beta
        out[oi, i] += -gamma * 0.5
    return out


def assemble_rows_general(nodes, conn, etype, centers, normals, areas, bc_type, bc_values, dof, is_eval, k, beta, rows,
                          harmonic=1.0, tau=1.0, gamma=1.0):
    """Rows (by DOF address) of build_tbem_system_with_beta for ANY boundary conditions, element by element as tbem.rs:126-218
    walks them -> (A[rows, :], rhs[rows]).  bc_type 0 velocity / 1 pressure / 2 transfer (contributes nothing, :239-242);
    bc_values[e]: the per-node values of element e (length 1..4).  Free terms: add_free_terms (:273-304) with the MEAN of the
    values; matrix entry: assemble_tbem (:311-345); right-hand side: the integrators' own term with the unscaled beta."""
    n_el = len(conn)
    bnd = [e for e in range(n_el) if not is_eval[e]]
    ndof = len(bnd)
    sign = dg_dn_sign(centers, k)
    A = np.zeros((len(rows), ndof), dtype=complex)
    rhs = np.zeros(len(rows), dtype=complex)
    by_dof = {int(dof[e]): e for e in bnd}
    for oi, d in enumerate(rows):
        i = by_dof[int(d)]
        x, nx = centers[i], normals[i]
        vals = [complex(v) for v in bc_values[i]]
        avg = sum(vals) / len(vals)
        if bc_type[i] == 0:
            A[oi, int(dof[i])] -= gamma * 0.5
            rhs[oi] += avg * beta * tau * 0.5
        elif bc_type[i] == 1:
            A[oi, int(dof[i])] -= beta * tau * 0.5
            rhs[oi] += avg * tau * 0.5
        for j in bnd:
            c = nodes[conn[j][: etype[j]]]
            fv = [complex(v) for v in bc_values[j]]
            compute_rhs = any(abs(v) > 1e-15 for v in fv)
            kw = dict(bc=fv, bc_type=int(bc_type[j]), tau=tau, gamma=gamma) if compute_rhs else {}
            if j == i:
                res = singular_self(x, nx, c, int(etype[j]), k, harmonic, **kw)
            else:
                res = regular_pair(x, nx, c, int(etype[j]), float(areas[j]), k, harmonic, **kw)
            g, h, ht, e = res[0], res[1] * sign, res[2], res[3]
            if bc_type[j] == 0:
                coeff = h * gamma * tau + e * beta
            elif bc_type[j] == 1:
                coeff = -(g * gamma * tau + ht * beta)
            else:
                coeff = 0j
            A[oi, int(dof[j])] += coeff
            if compute_rhs:
                rhs[oi] += res[4]
    return A, rhs


# ---- gmres.rs:105-277 ------------------------------------------------------------------------------------------
def _inner(x, y):
    s = 0j
    for a, b in zip(x, y):
        s += a.conjugate() * b
    return s


def _norm(x):
    s = 0.0
    for a in x:
        s += a.real * a.real + a.imag * a.imag
    return math.sqrt(s)


def gmres(apply, b, restart, tol, max_cycles, x0=None, sequential_blas=False, precond=None):
    """gmres_with_guess (gmres.rs:105-277) and, with `precond` (a function r -> M^-1 r), gmres_preconditioned_with_guess
    (gmres.rs:434-585): left preconditioning -- M^-1 b for the reference norm, M^-1 (b - A x) at the start of every cycle,
    M^-1 (A v_j) in every Arnoldi step, residuals relative to ||M^-1 b||.  `sequential_blas` = the reference's
    element-by-element inner products (slow in Python); otherwise numpy's vdot / norm (different summation order: same counts
    unless a decision sits on a rounding edge)."""
    inner = _inner if sequential_blas else (lambda x, y: complex(np.vdot(x, y)))
    norm = _norm if sequential_blas else (lambda x: float(np.linalg.norm(x)))
    plain_apply = apply
    if precond is not None:
        apply = lambda v: precond(plain_apply(v))  # noqa: E731  (w = M^-1 (A v_j), gmres.rs:513-514)
    n = len(b)
    m = restart
    x = np.zeros(n, dtype=complex) if x0 is None else np.array(x0, dtype=complex)
    b_norm = norm(b if precond is None else precond(np.array(b, dtype=complex)))
    if b_norm < 1e-15:
        return x, dict(iterations=0, restarts=0, residual=0.0, converged=True)
    its = restarts = 0
    for _ in range(max_cycles):
        r = b - plain_apply(x)
        if precond is not None:
            r = precond(r)
        beta = norm(r)
        rel = beta / b_norm
        if rel < tol:
            return x, dict(iterations=its, restarts=restarts, residual=rel, converged=True)
        v = [r * complex(1.0 / beta)]
        h = np.zeros((m + 1, m), dtype=complex)
        cs, sn = [], []
        g = np.zeros(m + 1, dtype=complex)
        g[0] = beta
        for j in range(m):
            its += 1
            w = apply(v[j])
            for i in range(j + 1):
                h[i, j] = inner(v[i], w)
                w = w - h[i, j] * v[i]
            wn = norm(w)
            h[j + 1, j] = wn
            broke = wn < 1e-14
            if not broke:
                v.append(w + complex(1.0 / wn - 1.0) * w)
            for i in range(j):
                tmp = cs[i].conjugate() * h[i, j] + sn[i].conjugate() * h[i + 1, j]
                h[i + 1, j] = 0j - sn[i] * h[i, j] + cs[i] * h[i + 1, j]
                h[i, j] = tmp
            a_, b_ = h[j, j], h[j + 1, j]
            if abs(b_) < 1e-30:
                c, s = 1 + 0j, 0j
            elif abs(a_) < 1e-30:
                c, s = 0j, 1 + 0j
            else:
                rr = math.sqrt(a_.real ** 2 + a_.imag ** 2 + b_.real ** 2 + b_.imag ** 2)
                c, s = a_ / rr, b_ / rr
            cs.append(c)
            sn.append(s)
            h[j, j] = c.conjugate() * h[j, j] + s.conjugate() * h[j + 1, j]
            h[j + 1, j] = 0j
            tmp = c.conjugate() * g[j] + s.conjugate() * g[j + 1]
            g[j + 1] = 0j - s * g[j] + c * g[j + 1]
            g[j] = tmp
            rel = abs(g[j + 1]) / b_norm
            if rel < tol or broke:
                y = _back(h, g, j + 1)
                for i, yi in enumerate(y):
                    x = x + yi * v[i]
                return x, dict(iterations=its, restarts=restarts, residual=rel, converged=True)
        y = _back(h, g, m)
        for i, yi in enumerate(y):
            x = x + yi * v[i]
        restarts += 1
    r = b - plain_apply(x)
    rel = norm(r if precond is None else precond(r)) / b_norm
    return x, dict(iterations=its, restarts=restarts, residual=rel, converged=False)


def row_sum_correction(A: np.ndarray) -> float:
    """apply_row_sum_correction (tbem.rs:500-520), in place: A[i, i] -= sum_j A[i, j], row by row with a running sum;
    returns |sum of all row sums| / n."""
    n = A.shape[0]
    total = 0j
    for i in range(n):
        rs = 0j
        for v in A[i]:
            rs += v
        total += rs
        A[i, i] -= rs
    return abs(total) / n


def _back(h, g, k):
    y = np.zeros(k, dtype=complex)
    for i in range(k - 1, -1, -1):
        s = g[i]
        for j in range(i + 1, k):
            s -= h[i, j] * y[j]
        if abs(h[i, i]) > 1e-30:
            y[i] = s / h[i, i]
    return y


# ---- neighbours of the path, again from the Rust sources alone --------------------------------
# core/incident.rs:93-342 and core/postprocess/pressure.rs:81-259, 438-478.  Vectorised over collocation / evaluation points,
# so the summation order differs from the C++ oracle's loops on purpose (agreement is to rounding, not to the bit).
def incident_pressure(kind: str, vec, amplitude: complex, points: np.ndarray, k: float) -> np.ndarray:
    """incident.rs:93-166.  kind 'plane': p = A exp(i k d.x) (direction used as given, incident.rs:106-116);
    kind 'point': p = S exp(ikr)/(4 pi r) for r > 1e-10, else 0 (:119-132)."""
    pts = np.asarray(points, dtype=float)
    v = np.asarray(vec, dtype=float)
    if kind == "plane":
        kdotx = k * (pts @ v)
        return amplitude * (np.cos(kdotx) + 1j * np.sin(kdotx))
    d = pts - v
    r = np.sqrt(np.sum(d * d, axis=1))
    out = np.zeros(len(pts), dtype=complex)
    ok = r > 1e-10
    kr = k * r[ok]
    out[ok] = amplitude * ((np.cos(kr) + 1j * np.sin(kr)) / (4.0 * math.pi * r[ok]))
    return out


def incident_normal_derivative(kind: str, vec, amplitude: complex, points: np.ndarray, normals: np.ndarray, k: float) -> np.ndarray:
    """incident.rs:177-280.  plane: dp/dn = i k (d.n) p (:196-208); point: S (ik - 1/r) G (x - x0).n / r (:210-233)."""
    pts, nrm = np.asarray(points, dtype=float), np.asarray(normals, dtype=float)
    v = np.asarray(vec, dtype=float)
    if kind == "plane":
        return 1j * (k * (nrm @ v)) * incident_pressure(kind, v, amplitude, pts, k)
    d = pts - v
    r = np.sqrt(np.sum(d * d, axis=1))
    out = np.zeros(len(pts), dtype=complex)
    ok = r > 1e-10
    ro = r[ok]
    kr = k * ro
    g = (np.cos(kr) + 1j * np.sin(kr)) / (4.0 * math.pi * ro)
    dgdr = (1j * k - 1.0 / ro) * g
    drdn = np.sum(d[ok] * nrm[ok], axis=1) / ro
    out[ok] = amplitude * dgdr * drdn
    return out


def incident_rhs(sources, points, normals, k: float, beta: complex, tau: float = 1.0, gamma: float = 1.0) -> np.ndarray:
    """compute_rhs_with_beta (incident.rs:317-342): rhs = -(gamma p_inc + beta tau dp_inc/dn); `sources` is a list of
    (kind, vector, amplitude) -- one entry for PlaneWave / PointSource, several for the Multiple* variants (:136-165, :235-276)."""
    p = sum(incident_pressure(kd, v, a, points, k) for kd, v, a in sources)
    dp = sum(incident_normal_derivative(kd, v, a, points, normals, k) for kd, v, a in sources)
    return -(gamma * p + beta * tau * dp)


def scattered_field(nodes, conn, is_eval, eval_points, surface_pressure, surface_velocity, k: float, harmonic: float = 1.0) -> np.ndarray:
    """compute_scattered_field + integrate_element_field (pressure.rs:81-259): entry j of the surface vectors belongs to the
    j-th NON-EVALUATION element (:96-113); every element -- Quad4 too -- is integrated over the triangle of its first three
    nodes with the 7-point rule (order 3, :162-198); G = exp(i w r)/(4 pi r) with w = k * harmonic_factor (:92, :236-238);
    result += p dG/dn_y J w_q - [ |v| > 1e-15 ] v G J w_q (:248-256); quadrature points with J < 1e-15 or r < 1e-15 are skipped."""
    nodes = np.asarray(nodes, dtype=float)
    bnd = [e for e in range(len(conn)) if not is_eval[e]]
    tri = triangle_quadrature(3)
    xi, eta, wq = tri[:, 0], tri[:, 1], tri[:, 2]
    N = np.stack([1.0 - xi - eta, xi, eta], axis=1)                 # (7, 3)
    c = np.stack([nodes[np.asarray(conn)[bnd][:, v].astype(np.int64)] for v in range(3)], axis=1)   # (nb, 3 nodes, 3)
    y = np.einsum("qn,bnd->bqd", N, c)                              # quadrature points (nb, 7, 3)
    dxds = c[:, 1] - c[:, 0]
    dxdt = c[:, 2] - c[:, 0]
    nv = np.cross(dxds, dxdt)
    jac = np.sqrt(np.sum(nv * nv, axis=1))
    good = jac >= 1e-15
    en = np.zeros_like(nv)
    en[good] = nv[good] / jac[good][:, None]
    w = k * harmonic
    ps = np.asarray(surface_pressure, dtype=complex)
    vs = None if surface_velocity is None else np.asarray(surface_velocity, dtype=complex)
    out = np.zeros(len(eval_points), dtype=complex)
    for i, x in enumerate(np.asarray(eval_points, dtype=float)):
        rv = y - x                                                   # (nb, 7, 3)
        r = np.sqrt(np.sum(rv * rv, axis=2))
        use = good[:, None] & (r >= 1e-15)
        rs = np.where(use, r, 1.0)
        g = (np.cos(w * rs) + 1j * np.sin(w * rs)) / (4.0 * math.pi * rs)
        dgdn = g * (-1.0 / rs + 1j * w) * (np.einsum("bqd,bd->bq", rv, en) / rs)
        vjacwe = jac[:, None] * wq[None, :]
        contrib = ps[:, None] * dgdn * vjacwe
        if vs is not None:
            has_v = np.abs(vs) > 1e-15
            contrib = contrib - np.where(has_v[:, None], vs[:, None] * g * vjacwe, 0.0)
        out[i] = np.sum(np.where(use, contrib, 0.0))
    return out


def rcs(centers, normals, areas, is_eval, surface_pressure, directions, k: float) -> np.ndarray:
    """compute_rcs (pressure.rs:438-478): F(d) = sum_j p_j exp(-i k c_j.d) A_j (i k)(n_j.d) over the non-evaluation elements in
    enumeration order, RCS = 4 pi |F|^2."""
    keep = np.asarray(is_eval) == 0
    c, n, a = np.asarray(centers, dtype=float)[keep], np.asarray(normals, dtype=float)[keep], np.asarray(areas, dtype=float)[keep]
    ps = np.asarray(surface_pressure, dtype=complex)
    out = []
    for d in np.asarray(directions, dtype=float).reshape(-1, 3):
        phase = -k * (c @ d)
        far = np.sum(ps * (np.cos(phase) + 1j * np.sin(phase)) * a * (1j * k) * (n @ d))
        out.append(4.0 * math.pi * (far.real ** 2 + far.imag ** 2))
    return np.array(out)
