"""Boundary-element meshes in the SoA layout the C ABI consumes.

Host-side mirror of the reference's mesh data contract and generators:

* ``Element`` / ``Mesh``                    math-bem/src/core/types.rs:329-398
* ``generate_sphere_mesh`` (UV sphere)      math-bem/src/core/mesh/generators.rs:29-98
* ``generate_icosphere_mesh``               math-bem/src/core/mesh/generators.rs:110-228
* ``create_mesh_from_data`` / geometry      math-bem/src/core/mesh/generators.rs:435-602

plus two synthetic inputs SURVEY.md section 8d names for the benchmark configs
(class-I geodesic icosphere for ~120k elements, closed Quad4 "cabinet" box).

The reference stores elements as an array of structs with heap vectors; the
device wants structure-of-arrays, so :class:`Mesh` holds one contiguous numpy
array per field (see ``include/bemb200.h``, ``bemb200_mesh``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

PAD = np.uint32(0xFFFFFFFF)

BC_VELOCITY = 0   # BoundaryCondition::Velocity / VelocityWithAdmittance  (tbem.rs:236-238)
BC_PRESSURE = 1   # BoundaryCondition::Pressure                            (tbem.rs:237)
BC_TRANSFER = 2   # Transfer* variants: contribute nothing                 (tbem.rs:239-242)


@dataclass
class Mesh:
    """SoA boundary mesh. One DOF per non-evaluation element (types.rs:359-363)."""

    nodes: np.ndarray      # (n_nodes, 3) f64
    conn: np.ndarray       # (n_elem, 4) u32, Tri3 rows padded with 0xFFFFFFFF
    etype: np.ndarray      # (n_elem,) u8: 3 = Tri3, 4 = Quad4
    center: np.ndarray     # (n_elem, 3) f64  collocation points
    normal: np.ndarray     # (n_elem, 3) f64  n_x (outward-flipped, generators.rs:590-601)
    area: np.ndarray       # (n_elem,) f64    drives the subdivision ratio test
    bc_type: np.ndarray    # (n_elem,) i32
    bc_len: np.ndarray     # (n_elem,) u8     number of per-node BC values given (1..4)
    bc_val: np.ndarray     # (n_elem, 4) c128 per-node BC values (first bc_len used)
    dof: np.ndarray        # (n_elem,) u32    dof_addresses[0]
    is_eval: np.ndarray    # (n_elem,) u8     ElementProperty::Evaluation
    meta: dict = field(default_factory=dict)

    @property
    def n_elem(self) -> int:
        return int(self.etype.shape[0])

    @property
    def n_nodes(self) -> int:
        return int(self.nodes.shape[0])

    @property
    def num_dofs(self) -> int:
        """count_dofs(): tbem.rs:225-231."""
        return int((self.is_eval == 0).sum())

    def set_velocity_bc(self, values=None) -> None:
        """Rigid by default: ``Velocity(vec![0])`` on every element (qa_suite.rs:221-224)."""
        n = self.n_elem
        self.bc_type[:] = BC_VELOCITY
        self.bc_val[:] = 0.0
        if values is None:
            self.bc_len[:] = 1
        else:
            v = np.asarray(values, dtype=np.complex128)
            assert v.shape[0] == n
            if v.ndim == 1:
                self.bc_len[:] = 1
                self.bc_val[:, 0] = v
            else:
                self.bc_len[:] = v.shape[1]
                self.bc_val[:, : v.shape[1]] = v

    def validate(self) -> None:
        n = self.n_elem
        assert self.nodes.dtype == np.float64 and self.nodes.ndim == 2 and self.nodes.shape[1] == 3
        assert self.conn.dtype == np.uint32 and self.conn.shape == (n, 4)
        assert self.etype.dtype == np.uint8 and set(np.unique(self.etype)) <= {3, 4}
        for a in (self.center, self.normal):
            assert a.dtype == np.float64 and a.shape == (n, 3)
        assert self.area.dtype == np.float64 and self.area.shape == (n,)
        assert self.bc_type.dtype == np.int32 and self.bc_len.dtype == np.uint8
        assert self.bc_val.dtype == np.complex128 and self.bc_val.shape == (n, 4)
        assert self.dof.dtype == np.uint32 and self.is_eval.dtype == np.uint8
        assert (self.bc_len >= 1).all() and (self.bc_len <= 4).all()
        nd = self.num_dofs
        d = self.dof[self.is_eval == 0]
        assert d.size == 0 or (np.sort(d) == np.arange(nd, dtype=np.uint32)).all(), "dofs must be a permutation of 0..ndof"


def mesh_from_data(nodes, connectivity) -> Mesh:
    """create_mesh_from_data + compute_element_geometry: generators.rs:435-602.

    ``connectivity`` is a list/array of 3- or 4-node rows.  Every element gets the
    default ``Velocity([0])`` BC and ``dof = index`` (generators.rs:466-468).
    """
    nodes = np.ascontiguousarray(np.asarray(nodes, dtype=np.float64).reshape(-1, 3))
    n = len(connectivity)
    conn = np.full((n, 4), PAD, dtype=np.uint32)
    etype = np.empty(n, dtype=np.uint8)
    if isinstance(connectivity, np.ndarray) and connectivity.ndim == 2:
        w = connectivity.shape[1]
        conn[:, :w] = connectivity.astype(np.uint32)
        etype[:] = w
    else:
        for i, c in enumerate(connectivity):
            conn[i, : len(c)] = c
            etype[i] = len(c)
    center = np.zeros((n, 3))
    normal = np.zeros((n, 3))
    area = np.zeros(n)
    tri = etype == 3
    quad = etype == 4
    if tri.any():
        c = conn[tri]
        p0, p1, p2 = nodes[c[:, 0]], nodes[c[:, 1]], nodes[c[:, 2]]
        center[tri] = (p0 + p1 + p2) / 3.0
        cr = np.cross(p1 - p0, p2 - p0)
        ln = np.sqrt(cr[:, 0] * cr[:, 0] + cr[:, 1] * cr[:, 1] + cr[:, 2] * cr[:, 2])
        area[tri] = ln / 2.0
        ok = ln > 1e-15
        nrm = np.zeros_like(cr)
        nrm[ok] = cr[ok] / ln[ok, None]
        normal[tri] = nrm
    if quad.any():
        c = conn[quad]
        p0, p1, p2, p3 = nodes[c[:, 0]], nodes[c[:, 1]], nodes[c[:, 2]], nodes[c[:, 3]]
        center[quad] = (p0 + p1 + p2 + p3) / 4.0
        cr = np.cross(p2 - p0, p3 - p1)  # diagonal cross product (generators.rs:566-584)
        ln = np.sqrt(cr[:, 0] * cr[:, 0] + cr[:, 1] * cr[:, 1] + cr[:, 2] * cr[:, 2])
        area[quad] = ln / 2.0
        ok = ln > 1e-15
        nrm = np.zeros_like(cr)
        nrm[ok] = cr[ok] / ln[ok, None]
        normal[quad] = nrm
    # outward flip of the STORED normal only (connectivity untouched): generators.rs:590-601
    ndc = normal[:, 0] * center[:, 0] + normal[:, 1] * center[:, 1] + normal[:, 2] * center[:, 2]
    flip = ndc < 0.0
    normal[flip] = -normal[flip]
    m = Mesh(
        nodes=nodes,
        conn=conn,
        etype=etype,
        center=np.ascontiguousarray(center),
        normal=np.ascontiguousarray(normal),
        area=area,
        bc_type=np.zeros(n, dtype=np.int32),
        bc_len=np.ones(n, dtype=np.uint8),
        bc_val=np.zeros((n, 4), dtype=np.complex128),
        dof=np.arange(n, dtype=np.uint32),
        is_eval=np.zeros(n, dtype=np.uint8),
        meta={"n_flipped": int(flip.sum())},
    )
    return m


def generate_sphere_mesh(radius: float, n_theta: int, n_phi: int) -> Mesh:
    """UV sphere: generators.rs:29-98 (2*n_phi*(n_theta-1) triangles)."""
    nodes = [[0.0, 0.0, radius]]
    for i in range(1, n_theta):
        theta = math.pi * i / n_theta
        st, ct = math.sin(theta), math.cos(theta)
        for j in range(n_phi):
            phi = 2.0 * math.pi * j / n_phi
            nodes.append([radius * st * math.cos(phi), radius * st * math.sin(phi), radius * ct])
    nodes.append([0.0, 0.0, -radius])
    south = len(nodes) - 1
    el = []
    for j in range(n_phi):
        el.append([0, 1 + j, 1 + (j + 1) % n_phi])
    for i in range(n_theta - 2):
        r0 = 1 + i * n_phi
        r1 = 1 + (i + 1) * n_phi
        for j in range(n_phi):
            jn = (j + 1) % n_phi
            n0, n1, n2, n3 = r0 + j, r0 + jn, r1 + j, r1 + jn
            el.append([n0, n2, n1])
            el.append([n1, n2, n3])
    last = 1 + (n_theta - 2) * n_phi
    for j in range(n_phi):
        el.append([last + j, south, last + (j + 1) % n_phi])
    return mesh_from_data(np.array(nodes), np.array(el, dtype=np.uint32))


def generate_cylinder_mesh(radius: float, height: float, n_circumference: int, n_height: int) -> Mesh:
    """Open cylinder (lateral surface only), Quad4: generators.rs:242-285."""
    nodes, el = [], []
    z_min = -height / 2.0
    dz = height / n_height
    for i in range(n_height + 1):
        z = z_min + i * dz
        for j in range(n_circumference):
            phi = 2.0 * math.pi * j / n_circumference
            nodes.append([radius * math.cos(phi), radius * math.sin(phi), z])
    for i in range(n_height):
        r0, r1 = i * n_circumference, (i + 1) * n_circumference
        for j in range(n_circumference):
            jn = (j + 1) % n_circumference
            el.append([r0 + j, r0 + jn, r1 + jn, r1 + j])
    return mesh_from_data(np.array(nodes), np.array(el, dtype=np.uint32))


def generate_closed_cylinder_mesh(radius: float, height: float, n_circumference: int, n_height: int, n_cap_rings: int) -> Mesh:
    """Cylinder with end caps (Quad4 lateral surface and cap rings, Tri3 fans at the cap centres):
    generators.rs:287-432."""
    nodes, el = [], []
    z_min, z_max = -height / 2.0, height / 2.0
    dz = height / n_height
    nc = n_circumference
    for i in range(n_height + 1):
        z = z_min + i * dz
        for j in range(nc):
            phi = 2.0 * math.pi * j / nc
            nodes.append([radius * math.cos(phi), radius * math.sin(phi), z])
    bottom_center = len(nodes)
    nodes.append([0.0, 0.0, z_min])
    for ring in range(1, n_cap_rings + 1):
        r = radius * ring / n_cap_rings
        for j in range(nc):
            phi = 2.0 * math.pi * j / nc
            nodes.append([r * math.cos(phi), r * math.sin(phi), z_min])
    top_center = len(nodes)
    nodes.append([0.0, 0.0, z_max])
    for ring in range(1, n_cap_rings + 1):
        r = radius * ring / n_cap_rings
        for j in range(nc):
            phi = 2.0 * math.pi * j / nc
            nodes.append([r * math.cos(phi), r * math.sin(phi), z_max])
    for i in range(n_height):
        r0, r1 = i * nc, (i + 1) * nc
        for j in range(nc):
            jn = (j + 1) % nc
            el.append([r0 + j, r0 + jn, r1 + jn, r1 + j])
    b1 = bottom_center + 1
    for j in range(nc):
        jn = (j + 1) % nc
        el.append([bottom_center, b1 + jn, b1 + j])
    for ring in range(n_cap_rings - 1):
        rs = bottom_center + 1 + ring * nc
        ns = rs + nc
        for j in range(nc):
            jn = (j + 1) % nc
            el.append([rs + j, rs + jn, ns + jn, ns + j])
    outer_bottom = bottom_center + 1 + (n_cap_rings - 1) * nc
    for j in range(nc):
        jn = (j + 1) % nc
        el.append([outer_bottom + j, outer_bottom + jn, jn, j])
    t1 = top_center + 1
    for j in range(nc):
        jn = (j + 1) % nc
        el.append([top_center, t1 + j, t1 + jn])
    for ring in range(n_cap_rings - 1):
        rs = top_center + 1 + ring * nc
        ns = rs + nc
        for j in range(nc):
            jn = (j + 1) % nc
            el.append([rs + j, ns + j, ns + jn, rs + jn])
    outer_top = top_center + 1 + (n_cap_rings - 1) * nc
    top_row = n_height * nc
    for j in range(nc):
        jn = (j + 1) % nc
        el.append([top_row + j, top_row + jn, outer_top + jn, outer_top + j])
    return mesh_from_data(np.array(nodes), el)


_ICO_FACES = [
    [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
    [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
    [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
    [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
]


def _ico_vertices():
    phi = (1.0 + math.sqrt(5.0)) / 2.0
    v = [
        [-1.0, phi, 0.0], [1.0, phi, 0.0], [-1.0, -phi, 0.0], [1.0, -phi, 0.0],
        [0.0, -1.0, phi], [0.0, 1.0, phi], [0.0, -1.0, -phi], [0.0, 1.0, -phi],
        [phi, 0.0, -1.0], [phi, 0.0, 1.0], [-phi, 0.0, -1.0], [-phi, 0.0, 1.0],
    ]
    out = []
    for x, y, z in v:
        ln = math.sqrt(x * x + y * y + z * z)
        out.append([x / ln, y / ln, z / ln])
    return out


def generate_icosphere_mesh(radius: float, subdivisions: int) -> Mesh:
    """Recursive icosphere: generators.rs:110-228 (20*4^s triangles)."""
    verts = _ico_vertices()
    faces = [list(f) for f in _ICO_FACES]
    for _ in range(subdivisions):
        cache = {}
        new_faces = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            idx = cache.get(key)
            if idx is not None:
                return idx
            va, vb = verts[a], verts[b]
            m = [(va[0] + vb[0]) / 2.0, (va[1] + vb[1]) / 2.0, (va[2] + vb[2]) / 2.0]
            ln = math.sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2])
            verts.append([m[0] / ln, m[1] / ln, m[2] / ln])
            cache[key] = len(verts) - 1
            return len(verts) - 1

        for v0, v1, v2 in faces:
            m01 = mid(v0, v1)
            m12 = mid(v1, v2)
            m20 = mid(v2, v0)
            new_faces += [[v0, m01, m20], [v1, m12, m01], [v2, m20, m12], [m01, m12, m20]]
        faces = new_faces
    nodes = np.array(verts) * radius
    return mesh_from_data(nodes, np.array(faces, dtype=np.uint32))


def generate_geodesic_sphere_mesh(radius: float, nu: int) -> Mesh:
    """Class-I geodesic icosphere, 20*nu^2 triangles (nu=78 -> 121 680).

    Not a reference generator: same icosahedron, face winding and flip rule as
    generators.rs:110-160, but each face is split into nu^2 triangles directly so that
    element counts other than 20*4^s are reachable.
    """
    base = np.array(_ico_vertices())
    key_to_idx = {}
    verts = []

    def vid(p):
        k = (round(p[0] * 1e9), round(p[1] * 1e9), round(p[2] * 1e9))
        i = key_to_idx.get(k)
        if i is None:
            i = len(verts)
            key_to_idx[k] = i
            verts.append(p)
        return i

    faces = []
    for fa, fb, fc in _ICO_FACES:
        A, B, Cc = base[fa], base[fb], base[fc]
        idx = {}
        for i in range(nu + 1):
            for j in range(nu + 1 - i):
                p = (A * (nu - i - j) + B * i + Cc * j) / nu
                p = p / math.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2])
                idx[(i, j)] = vid((float(p[0]), float(p[1]), float(p[2])))
        for i in range(nu):
            for j in range(nu - i):
                faces.append([idx[(i, j)], idx[(i + 1, j)], idx[(i, j + 1)]])
                if i + j < nu - 1:
                    faces.append([idx[(i + 1, j)], idx[(i + 1, j + 1)], idx[(i, j + 1)]])
    nodes = np.array(verts) * radius
    return mesh_from_data(nodes, np.array(faces, dtype=np.uint32))


def generate_box_mesh_quad(lx: float, ly: float, lz: float, nx: int, ny: int, nz: int) -> Mesh:
    """Closed box of Quad4 elements, outward (CCW from outside) winding, centred at the origin.

    SURVEY.md 8d config 3 ("cabinet"): 0.32 x 0.44 x 0.64 m, 64 x 88 x 128 -> 50 176 quads.
    The reference's room mesher emits the same kind of per-wall Quad4 grid
    (math-xem-common/src/geometry.rs:434-469).
    """
    key_to_idx = {}
    verts = []

    def vid(ix, iy, iz):
        k = (ix, iy, iz)
        i = key_to_idx.get(k)
        if i is None:
            i = len(verts)
            key_to_idx[k] = i
            verts.append([-lx / 2 + lx * ix / nx, -ly / 2 + ly * iy / ny, -lz / 2 + lz * iz / nz])
        return i

    quads = []
    # z = -lz/2 (normal -z) and z = +lz/2 (normal +z)
    for ix in range(nx):
        for iy in range(ny):
            quads.append([vid(ix, iy, 0), vid(ix, iy + 1, 0), vid(ix + 1, iy + 1, 0), vid(ix + 1, iy, 0)])
            quads.append([vid(ix, iy, nz), vid(ix + 1, iy, nz), vid(ix + 1, iy + 1, nz), vid(ix, iy + 1, nz)])
    # y = -ly/2 (normal -y) and y = +ly/2
    for ix in range(nx):
        for iz in range(nz):
            quads.append([vid(ix, 0, iz), vid(ix + 1, 0, iz), vid(ix + 1, 0, iz + 1), vid(ix, 0, iz + 1)])
            quads.append([vid(ix, ny, iz), vid(ix, ny, iz + 1), vid(ix + 1, ny, iz + 1), vid(ix + 1, ny, iz)])
    # x = -lx/2 and x = +lx/2
    for iy in range(ny):
        for iz in range(nz):
            quads.append([vid(0, iy, iz), vid(0, iy, iz + 1), vid(0, iy + 1, iz + 1), vid(0, iy + 1, iz)])
            quads.append([vid(nx, iy, iz), vid(nx, iy + 1, iz), vid(nx, iy + 1, iz + 1), vid(nx, iy, iz + 1)])
    m = mesh_from_data(np.array(verts), np.array(quads, dtype=np.uint32))
    return m


def fibonacci_directions(n: int) -> np.ndarray:
    """Incident directions of SURVEY.md 8d config 5: z = 1-(2m+1)/n, phi = m*pi*(3-sqrt(5))."""
    out = np.zeros((n, 3))
    for m in range(n):
        z = 1.0 - (2 * m + 1) / n
        r = math.sqrt(max(0.0, 1.0 - z * z))
        ph = m * math.pi * (3.0 - math.sqrt(5.0))
        out[m] = [r * math.cos(ph), r * math.sin(ph), z]
    return out
