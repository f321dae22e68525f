"""CPU restatement of the reference's room-acoustics dense path -- TEST INFRASTRUCTURE ONLY.

Only tests/ (incl. the measurement drivers under tests/drivers/), __graft_entry__.smoke() and the CPU-baseline legs of bench.py may import
this module; the product (math_audio_b200/) never does.

Follows, line by line where arithmetic order matters:
  math-bem/src/room_acoustics/solver.rs
      greens_function_3d :18-24, greens_function_derivative :28-35,
      element_center_and_normal :38-68, element_area :70-122,
      build_bem_matrix_parallel :448-493, solve_bem_system :412-445 (GMRES through oracle.gmres),
      calculate_incident_field_derivative_parallel :638-679,
      calculate_field_pressure_bem_parallel :687-748
  math-xem-common/src/geometry.rs  RectangularRoom::generate_mesh :107-183, add_surface_mesh :434-469,
      LShapedRoom::generate_mesh :500-627
  math-xem-common/src/source.rs    DirectivityPattern::{omnidirectional, cardioid, interpolate} :19-98,
      CrossoverFilter::amplitude_at_frequency :127-155, Source::amplitude_towards :203-219
  math-xem-common/src/types.rs     pressure_to_spl :280-287, log_space :290-302

Parity status: the reference holds no numeric golden vectors for this path (its tests are the
sanity checks restated in tests/test_room_cpu.py) => "parity unpinned by reference vectors,
pinned by faithful restatement", as for the TBEM path.

Scalar loops are literal (small cases); the O(N^2) matrix is also offered vectorised with numpy
(same formulas, element-wise IEEE operations in the same order) for sizes the loops cannot reach.
"""
from __future__ import annotations

import cmath
import math

import numpy as np

PI = math.pi


# ---- solver.rs:18-35 ------------------------------------------------------------------------
def greens_function_3d(r, k):
    if r < 1e-10:
        return 0j
    return cmath.exp(1j * (k * r)) / (4.0 * PI * r)


def greens_function_derivative(r, k, cos_angle):
    if r < 1e-10:
        return 0j
    ikr = complex(0.0, k * r)
    factor = (ikr - 1.0) * cmath.exp(ikr) / (4.0 * PI * r * r)
    return factor * cos_angle


# ---- solver.rs:38-122 -----------------------------------------------------------------------
def element_center_and_normal(nodes):
    n = len(nodes)
    center = [0.0, 0.0, 0.0]
    for d in range(3):
        s = 0.0
        for p in nodes:
            s += p[d]
        center[d] = s / n
    v1 = [nodes[1][d] - nodes[0][d] for d in range(3)]
    v2 = [nodes[2][d] - nodes[0][d] for d in range(3)]
    nx = v1[1] * v2[2] - v1[2] * v2[1]
    ny = v1[2] * v2[0] - v1[0] * v2[2]
    nz = v1[0] * v2[1] - v1[1] * v2[0]
    norm = math.sqrt(nx * nx + ny * ny + nz * nz)
    return center, [nx / norm, ny / norm, nz / norm]


def element_area(nodes):
    def half_cross(a, b):
        cx = a[1] * b[2] - a[2] * b[1]
        cy = a[2] * b[0] - a[0] * b[2]
        cz = a[0] * b[1] - a[1] * b[0]
        return 0.5 * math.sqrt(cx * cx + cy * cy + cz * cz)

    if len(nodes) == 3:
        v1 = [nodes[1][d] - nodes[0][d] for d in range(3)]
        v2 = [nodes[2][d] - nodes[0][d] for d in range(3)]
        return half_cross(v1, v2)
    if len(nodes) == 4:
        v1 = [nodes[1][d] - nodes[0][d] for d in range(3)]
        v2 = [nodes[2][d] - nodes[0][d] for d in range(3)]
        v3 = [nodes[3][d] - nodes[0][d] for d in range(3)]
        return half_cross(v1, v2) + half_cross(v2, v3)
    return 0.0


def element_data(mesh_nodes, mesh_elements):
    """(center, normal, area) per element; mesh_elements rows hold 3 or 4 node indices (0xFFFFFFFF pad allowed)."""
    centers, normals, areas = [], [], []
    for el in mesh_elements:
        ids = [int(i) for i in el if int(i) != 0xFFFFFFFF]
        pts = [[float(v) for v in mesh_nodes[i]] for i in ids]
        c, n = element_center_and_normal(pts)
        centers.append(c)
        normals.append(n)
        areas.append(element_area(pts))
    return np.array(centers), np.array(normals), np.array(areas)


# ---- solver.rs:448-493 ----------------------------------------------------------------------
def build_bem_matrix_loops(centers, normals, areas, k):
    n = len(areas)
    a = np.zeros((n, n), dtype=np.complex128)
    for i in range(n):
        ci, ni = centers[i], normals[i]
        for j in range(n):
            cj = centers[j]
            r = math.sqrt((ci[0] - cj[0]) ** 2 + (ci[1] - cj[1]) ** 2 + (ci[2] - cj[2]) ** 2)
            if i == j:
                a[i, j] = complex(0.0, -k / (2.0 * PI)) * areas[j]
            else:
                dx, dy, dz = ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]
                cos_angle = (dx * ni[0] + dy * ni[1] + dz * ni[2]) / r
                a[i, j] = greens_function_derivative(r, k, cos_angle) * areas[j]
    return a


def build_bem_matrix(centers, normals, areas, k, rows=None):
    """Vectorised form of the same formulas (rows = (r0, r1) for a slab)."""
    n = len(areas)
    r0, r1 = rows if rows is not None else (0, n)
    d = centers[r0:r1, None, :] - centers[None, :, :]
    r = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        cos_angle = (d[..., 0] * normals[r0:r1, None, 0] + d[..., 1] * normals[r0:r1, None, 1] + d[..., 2] * normals[r0:r1, None, 2]) / r
        ikr = 1j * (k * r)
        factor = (ikr - 1.0) * np.exp(ikr) / (4.0 * PI * r * r)
        a = factor * cos_angle * areas[None, :]
    a[r < 1e-10] = 0.0
    ii = np.arange(r0, r1)
    a[ii - r0, ii] = complex(0.0, -k / (2.0 * PI)) * areas[ii]
    return a


# ---- source.rs ------------------------------------------------------------------------------
def directivity_omnidirectional():
    return [[1.0] * 36 for _ in range(19)]


def directivity_cardioid():
    mag = [[0.0] * 36 for _ in range(19)]
    for v_idx in range(19):
        for h_idx in range(36):
            theta_rad = math.radians(v_idx * 10.0)
            phi_rad = math.radians(h_idx * 10.0)
            forward_dot = math.sin(theta_rad) * math.sin(phi_rad)
            mag[v_idx][h_idx] = 0.5 * max(1.0 + forward_dot, 0.0)
    return mag


def directivity_interpolate(mag, theta, phi):
    nv, nh = len(mag), len(mag[0])
    theta_deg = math.degrees(theta)
    phi_deg = math.degrees(phi)
    while phi_deg < 0.0:
        phi_deg += 360.0
    while phi_deg >= 360.0:
        phi_deg -= 360.0
    h_idx = min(int(math.floor(phi_deg / 10.0)), nh - 1)
    v_idx = min(int(math.floor(theta_deg / 10.0)), nv - 1)
    h_next = (h_idx + 1) % nh
    v_next = min(v_idx + 1, nv - 1)
    h_frac = phi_deg / 10.0 - h_idx
    v_frac = theta_deg / 10.0 - v_idx
    m0 = mag[v_idx][h_idx] * (1.0 - h_frac) + mag[v_idx][h_next] * h_frac
    m1 = mag[v_next][h_idx] * (1.0 - h_frac) + mag[v_next][h_next] * h_frac
    return m0 * (1.0 - v_frac) + m1 * v_frac


def crossover_amplitude(kind, frequency, cutoff=0.0, low=0.0, high=0.0, order=2):
    if kind == "fullrange":
        return 1.0
    if kind == "lowpass":
        return 1.0 / math.sqrt(1.0 + (frequency / cutoff) ** (order * 2))
    if kind == "highpass":
        return 1.0 / math.sqrt(1.0 + (cutoff / frequency) ** (order * 2))
    if kind == "bandpass":
        hp = 1.0 / math.sqrt(1.0 + (low / frequency) ** (order * 2))
        lp = 1.0 / math.sqrt(1.0 + (frequency / high) ** (order * 2))
        return hp * lp
    raise ValueError(kind)


def amplitude_towards(src, point, frequency):
    """src = dict(position, amplitude, directivity (19x36 list or None), crossover = dict(kind, ...))."""
    dx, dy, dz = point[0] - src["position"][0], point[1] - src["position"][1], point[2] - src["position"][2]
    r = math.sqrt(dx * dx + dy * dy + dz * dz)
    xo = crossover_amplitude(frequency=frequency, **src.get("crossover", {"kind": "fullrange"}))
    if r < 1e-10:
        return src["amplitude"] * xo
    theta = math.acos(dz / r)
    phi = math.atan2(dy, dx)
    mag = src.get("directivity") or directivity_omnidirectional()
    return src["amplitude"] * directivity_interpolate(mag, theta, phi) * xo


# ---- solver.rs:638-748 ----------------------------------------------------------------------
def incident_field_derivative(centers, normals, sources, k, frequency):
    out = np.zeros(len(centers), dtype=np.complex128)
    for i, (c, n) in enumerate(zip(centers, normals)):
        dpdn = 0j
        for s in sources:
            p = s["position"]
            r = math.sqrt((c[0] - p[0]) ** 2 + (c[1] - p[1]) ** 2 + (c[2] - p[2]) ** 2)
            if r < 1e-10:
                continue
            amplitude = amplitude_towards(s, c, frequency)
            dx, dy, dz = c[0] - p[0], c[1] - p[1], c[2] - p[2]
            cos_angle = (dx * n[0] + dy * n[1] + dz * n[2]) / r
            dpdn += greens_function_derivative(r, k, cos_angle) * amplitude
        out[i] = -dpdn
    return out


def field_pressure(centers, normals, areas, surface_pressure, sources, field_points, k, frequency):
    out = np.zeros(len(field_points), dtype=np.complex128)
    for ip, x in enumerate(field_points):
        p_inc = 0j
        for s in sources:
            p = s["position"]
            r = math.sqrt((x[0] - p[0]) ** 2 + (x[1] - p[1]) ** 2 + (x[2] - p[2]) ** 2)
            if r < 1e-10:
                continue
            p_inc += greens_function_3d(r, k) * amplitude_towards(s, x, frequency)
        d = np.asarray(x)[None, :] - centers
        r = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        ok = r >= 1e-10
        rr = np.where(ok, r, 1.0)
        cos_angle = (d[:, 0] * normals[:, 0] + d[:, 1] * normals[:, 1] + d[:, 2] * normals[:, 2]) / rr
        ikr = 1j * (k * rr)
        dg = (ikr - 1.0) * np.exp(ikr) / (4.0 * PI * rr * rr) * cos_angle
        terms = np.where(ok, dg * surface_pressure * areas, 0.0)
        p_scat = 0j
        for t in terms:  # sequential accumulation as the reference's loop
            p_scat += t
        out[ip] = p_inc + p_scat
    return out


# ---- geometry.rs:107-183, 434-469 -----------------------------------------------------------
def rectangular_room_mesh(width, depth, height, elements_per_meter):
    nx = int(math.ceil(width * elements_per_meter))
    ny = int(math.ceil(depth * elements_per_meter))
    nz = int(math.ceil(height * elements_per_meter))
    nodes, elements = [], []

    def add_surface_mesh(origin, u_dir, v_dir, nu, nv):
        base_idx = len(nodes)
        for j in range(nv + 1):
            for i in range(nu + 1):
                u = i / nu
                v = j / nv
                nodes.append([origin[d] + u * (u_dir[d] - origin[d]) + v * (v_dir[d] - origin[d]) for d in range(3)])
        for j in range(nv):
            for i in range(nu):
                n0 = base_idx + j * (nu + 1) + i
                n1 = base_idx + j * (nu + 1) + i + 1
                n2 = base_idx + (j + 1) * (nu + 1) + i + 1
                n3 = base_idx + (j + 1) * (nu + 1) + i
                elements.append([n0, n1, n2, n3])

    w, dpt, h = float(width), float(depth), float(height)
    add_surface_mesh([0.0, 0.0, 0.0], [w, 0.0, 0.0], [0.0, dpt, 0.0], nx, ny)  # floor
    add_surface_mesh([0.0, 0.0, h], [w, 0.0, h], [0.0, dpt, h], nx, ny)        # ceiling
    add_surface_mesh([0.0, 0.0, 0.0], [w, 0.0, 0.0], [0.0, 0.0, h], nx, nz)    # front wall
    add_surface_mesh([0.0, dpt, 0.0], [w, dpt, 0.0], [0.0, dpt, h], nx, nz)    # back wall
    add_surface_mesh([0.0, 0.0, 0.0], [0.0, dpt, 0.0], [0.0, 0.0, h], ny, nz)  # left wall
    add_surface_mesh([w, 0.0, 0.0], [w, dpt, 0.0], [w, 0.0, h], ny, nz)        # right wall
    return np.array(nodes), np.array(elements, dtype=np.int64)


def lshaped_room_mesh(width1, depth1, width2, depth2, height, elements_per_meter):
    """LShapedRoom::generate_mesh (geometry.rs:500-627) with add_surface_mesh_lshaped (:710-745)."""
    e = elements_per_meter
    nodes, elements = [], []

    def add(origin, u_dir, v_dir, nu, nv):
        base_idx = len(nodes)
        for j in range(nv + 1):
            for i in range(nu + 1):
                u = i / nu
                v = j / nv
                nodes.append([origin[d] + u * (u_dir[d] - origin[d]) + v * (v_dir[d] - origin[d]) for d in range(3)])
        for j in range(nv):
            for i in range(nu):
                n0 = base_idx + j * (nu + 1) + i
                elements.append([n0, n0 + 1, base_idx + (j + 1) * (nu + 1) + i + 1, base_idx + (j + 1) * (nu + 1) + i])

    w1, d1, w2, d2, h = float(width1), float(depth1), float(width2), float(depth2), float(height)
    nx1 = int(math.ceil(w1 * e)); ny1 = int(math.ceil(d1 * e)); nx2 = int(math.ceil(w2 * e)); ny2 = int(math.ceil(d2 * e))
    nz = int(math.ceil(h * e))
    add([0.0, 0.0, 0.0], [w1, 0.0, 0.0], [0.0, d1, 0.0], nx1, ny1)
    add([0.0, d1, 0.0], [w2, d1, 0.0], [0.0, d1 + d2, 0.0], nx2, ny2)
    add([0.0, 0.0, h], [w1, 0.0, h], [0.0, d1, h], nx1, ny1)
    add([0.0, d1, h], [w2, d1, h], [0.0, d1 + d2, h], nx2, ny2)
    add([0.0, 0.0, 0.0], [w1, 0.0, 0.0], [0.0, 0.0, h], nx1, nz)
    add([w1, 0.0, 0.0], [w1, d1, 0.0], [w1, 0.0, h], ny1, nz)
    total_depth = d1 + d2
    ny_total = int(math.ceil(total_depth * e))
    add([0.0, 0.0, 0.0], [0.0, total_depth, 0.0], [0.0, 0.0, h], ny_total, nz)
    add([0.0, total_depth, 0.0], [w2, total_depth, 0.0], [0.0, total_depth, h], nx2, nz)
    add([w2, d1, 0.0], [w2, total_depth, 0.0], [w2, d1, h], ny2, nz)
    internal_width = w1 - w2
    nx_internal = int(math.ceil(internal_width * e))
    add([w2, d1, 0.0], [w1, d1, 0.0], [w2, d1, h], nx_internal, nz)
    return np.array(nodes), np.array(elements, dtype=np.int64)


# ---- types.rs -------------------------------------------------------------------------------
def pressure_to_spl(p):
    m = abs(p)
    return 20.0 * math.log10(m / 20e-6) if m > 1e-20 else -120.0


def log_space(start, end, num):
    if num < 2:
        return [start]
    ls, le = math.log(start), math.log(end)
    return [math.exp(ls + (le - ls) * i / (num - 1)) for i in range(num)]
