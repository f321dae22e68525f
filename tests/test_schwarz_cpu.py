"""CPU tests of the additive Schwarz / block-Jacobi restatement (oracle/schwarz_oracle.py) and of the host helpers of its
device twin: the reference's own tests (math-solvers/src/preconditioners/schwarz.rs:462-572) restated, the line-by-line CSR
restatement against the dense-block one and against LAPACK, the partition helpers."""
import math

import numpy as np

from math_audio_b200 import bem
from oracle import schwarz_oracle as so


def reference_test_matrix():
    """schwarz.rs:468-492 create_test_matrix."""
    n = 20
    d = np.zeros((n, n), dtype=np.complex128)
    for i in range(n):
        d[i, i] = 4.0
        if i > 0:
            d[i, i - 1] = -1.0
        if i < n - 1:
            d[i, i + 1] = -1.0
        if i >= 5:
            d[i, i - 5] = -0.5
        if i < n - 5:
            d[i, i + 5] = -0.5
    return d


def csr_of(d):
    return so.dense_to_csr(d, np.abs(d) > 1e-15)  # CsrMatrix::from_dense(&dense, 1e-15)


def test_reference_schwarz_basic_and_stats():
    d = reference_test_matrix()
    v, c, p = csr_of(d)
    r = np.array([math.sin(i) for i in range(20)], dtype=np.complex128)
    z = so.CsrSchwarz(v, c, p, 20, 4, 1).apply(r)  # test_schwarz_basic
    assert z.shape == (20,) and np.all(np.abs(z) < 100.0)
    ns, mn, mx, avg = so.CsrSchwarz(v, c, p, 20, 4, 2).stats()  # test_schwarz_stats
    assert ns == 4 and mn > 0 and mx >= mn and avg > 5.0


def test_reference_schwarz_with_gmres_and_overlap(orc):
    d = reference_test_matrix()
    v, c, p = csr_of(d)
    b = np.array([math.sin(i) for i in range(20)], dtype=np.complex128)
    its = []
    for overlap in (0, 1, 2):  # test_schwarz_with_gmres, test_schwarz_overlap_effect
        pre = so.CsrSchwarz(v, c, p, 20, 4, overlap)
        x, info = orc.gmres_preconditioned_cb(lambda y: d @ y, pre.apply, 20, b, max_iterations=100, restart=20, tolerance=1e-8)
        assert info["converged"]
        assert np.linalg.norm(d @ x - b) / np.linalg.norm(b) < 1e-6
        its.append(info["iterations"])
    x0, plain = orc.gmres(d, b, max_iterations=100, restart=20, tolerance=1e-8)
    assert min(its) <= plain["iterations"]


def test_csr_restatement_equals_dense_block_restatement_and_lapack():
    rng = np.random.default_rng(7)
    n = 26
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 9.0 * np.eye(n)
    r = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    v, c, p = so.dense_to_csr(A)
    for S in (1, 3, 4, 26):
        zc = so.CsrSchwarz(v, c, p, n, S, 0).apply(r)
        zd = so.DenseSchwarz(A, S).apply(r)
        assert np.max(np.abs(zc - zd)) < 1e-14
        zl = np.zeros(n, dtype=np.complex128)
        for s in so.contiguous_partition(n, S):
            zl[s] = np.linalg.solve(A[np.ix_(s, s)], r[s])
        assert np.max(np.abs(zd - zl)) < 1e-13
    # a dense pattern couples everything: one overlap layer makes every subdomain the whole domain, weights 1/S
    zc = so.CsrSchwarz(v, c, p, n, 3, 1).apply(r)
    zd = so.DenseSchwarz(A, subdomains=[np.arange(n)] * 3).apply(r)
    assert np.max(np.abs(zc - zd)) < 1e-14
    assert np.max(np.abs(zc - np.linalg.solve(A, r))) < 1e-12


def test_tiny_pivot_is_skipped_like_the_reference():
    A = np.array([[0.0, 2.0], [3.0, 4.0]], dtype=np.complex128)  # u_00 = 0: nothing is eliminated with it (schwarz.rs:283-285)
    v, c, p = so.dense_to_csr(A)
    r = np.array([1.0, 2.0], dtype=np.complex128)
    zc = so.CsrSchwarz(v, c, p, 2, 1, 0).apply(r)
    zd = so.DenseSchwarz(A, 1).apply(r)
    assert np.allclose(zc, zd, rtol=0, atol=0)


def test_golden_schwarz_fixture(orc):
    """tests/golden/schwarz_ico1.npz was produced by the line-by-line CSR restatement; the dense-block restatement (what the GPU
    tests compare against at larger sizes) must reproduce it, M^-1 r to rounding and the preconditioned GMRES counts exactly."""
    from pathlib import Path

    gold = Path(__file__).resolve().parent / "golden"
    g, f = np.load(gold / "ico1_ka0p5.npz"), np.load(gold / "schwarz_ico1.npz")
    A, b, n = g["A"], g["b"], g["A"].shape[0]
    for S in (4, 7):
        ds = so.DenseSchwarz(A, S)
        z = ds.apply(f["r"])
        assert np.max(np.abs(z - f[f"z_S{S}"])) < 1e-13 * np.max(np.abs(z))
        x, info = orc.gmres_preconditioned_cb(lambda v: A @ v, ds.apply, n, b, max_iterations=100, restart=20, tolerance=1e-10)
        assert [info["iterations"], info["restarts"], int(info["converged"])] == list(f[f"info_S{S}"])
        assert np.linalg.norm(x - f[f"x_S{S}"]) / np.linalg.norm(x) < 1e-10
        assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-9


def test_partition_helpers():
    for n, S in [(20, 4), (10, 3), (7, 7), (5, 9), (1280, 10), (20480, 160)]:
        a = bem.schwarz_partition(n, S)
        b = so.contiguous_partition(n, S)
        assert len(a) == len(b) and all(np.array_equal(x.astype(np.int64), y) for x, y in zip(a, b))
        assert np.array_equal(np.concatenate(a).astype(np.int64), np.arange(n))
    for n, nr, bs in [(1280, 2, 128), (20480, 8, 128), (121680, 8, 256), (50176, 4, 200), (11, 4, 2)]:
        chunk = (n + nr - 1) // nr
        parts = bem.schwarz_partition_aligned(n, nr, bs)
        assert np.array_equal(np.concatenate(parts).astype(np.int64), np.arange(n))
        for q in parts:
            assert int(q[0]) // chunk == int(q[-1]) // chunk  # never straddles two ranks' row blocks
            assert len(q) <= 2 * bs
    adj = [[j for j in (i - 1, i + 1) if 0 <= j < 12] for i in range(12)]
    for part, ov in [([0, 1, 2], 1), ([5, 6], 2), ([11], 3), ([3, 4], 0)]:
        assert np.array_equal(bem.extend_partition(part, adj, ov, 12).astype(np.int64), so.extend_partition(part, adj, ov, 12))


def test_geometric_subdomains_are_rank_aligned_partitions():
    """bem.spatial_subdomains / bem.voronoi_subdomains (host logic of the config-3 / config-4 block-Jacobi legs): every DOF in
    exactly one cluster, no cluster across two ranks' row blocks, sizes bounded, deterministic, indices ascending."""
    from math_audio_b200.mesh import generate_box_mesh_quad, generate_geodesic_sphere_mesh

    for mesh, nranks, bs in [(generate_geodesic_sphere_mesh(1.0, 9), 4, 40), (generate_box_mesh_quad(0.32, 0.44, 0.64, 8, 11, 16), 2, 64),
                             (generate_geodesic_sphere_mesh(0.1, 5), 8, 16)]:
        n = mesh.num_dofs
        chunk = (n + nranks - 1) // nranks
        for fn in (bem.spatial_subdomains, bem.voronoi_subdomains):
            parts = fn(mesh.center[:n], nranks, bs)
            again = fn(mesh.center[:n], nranks, bs)
            assert len(parts) == len(again) and all(np.array_equal(a, b) for a, b in zip(parts, again))
            allidx = np.concatenate(parts).astype(np.int64)
            assert np.array_equal(np.sort(allidx), np.arange(n))
            for q in parts:
                q = q.astype(np.int64)
                assert len(q) >= 1 and np.all(np.diff(q) > 0)
                assert q[0] // chunk == q[-1] // chunk
            sizes = np.array([len(q) for q in parts])
            assert sizes.max() <= (bs if fn is bem.spatial_subdomains else 3 * bs)
            # compact: the mean cluster radius is a small multiple of the radius of a disc of the cluster's share of the surface
            c = mesh.center[:n]
            rad = np.mean([np.linalg.norm(c[q.astype(np.int64)] - c[q.astype(np.int64)].mean(axis=0), axis=1).max() for q in parts])
            area = float(np.sum(mesh.area[:n])) if hasattr(mesh, "area") else None
            if area:
                assert rad < 3.0 * np.sqrt(area * sizes.mean() / n / np.pi)
