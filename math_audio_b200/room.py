"""Room-acoustics dense BEM path -- host-side mirror of

* ``math-xem-common/src/geometry.rs``  RectangularRoom::generate_mesh :107-183, add_surface_mesh :434-469,
  LShapedRoom::generate_mesh :500-627
* ``math-xem-common/src/source.rs``    DirectivityPattern :9-98, CrossoverFilter :103-155, Source :160-219
* ``math-xem-common/src/types.rs``     wavenumber :275, pressure_to_spl :280-287, log_space :290-302, lin_space :305-312
* ``math-bem/src/room_acoustics/solver.rs``  build_bem_matrix_parallel :448-493, solve_bem_system :412-445,
  calculate_incident_field_derivative_parallel :638-679, calculate_field_pressure_bem_parallel :687-748
* ``math-bem/bin/room_simulator_bem.rs``     run_direct_gmres :225-281

Same names and argument meaning as the reference; all O(N^2) / O(N) arithmetic runs in
``csrc/room.cu`` behind the C ABI (no CPU fallback).  The mesh is staged once
(``StagedRoomMesh``) and the matrix buffer is reused across the frequencies of a sweep.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _capi
from . import bem

PAD = 0xFFFFFFFF
SPEED_OF_SOUND_20C = 343.0     # types.rs:265
REFERENCE_PRESSURE = 20e-6     # types.rs:268


# ---- types.rs ------------------------------------------------------------------------------
def wavenumber(frequency: float, speed_of_sound: float) -> float:
    return 2.0 * math.pi * frequency / speed_of_sound


def pressure_to_spl(pressure: complex) -> float:
    magnitude = abs(complex(pressure))
    if magnitude > 1e-20:
        return 20.0 * math.log10(magnitude / REFERENCE_PRESSURE)
    return -120.0


def log_space(start: float, end: float, num: int) -> List[float]:
    if num < 2:
        return [start]
    ls, le = math.log(start), math.log(end)
    return [math.exp(ls + (le - ls) * i / (num - 1)) for i in range(num)]


def lin_space(start: float, end: float, num: int) -> List[float]:
    if num < 2:
        return [start]
    return [start + (end - start) * i / (num - 1) for i in range(num)]


# ---- geometry.rs ---------------------------------------------------------------------------
@dataclass
class RoomMesh:
    """types.rs:185-192 as arrays: nodes (n_nodes, 3); elements (n_elem, 4) node indices, PAD in column 3 for triangles."""
    nodes: np.ndarray
    elements: np.ndarray

    def num_nodes(self) -> int:
        return int(self.nodes.shape[0])

    def num_elements(self) -> int:
        return int(self.elements.shape[0])


def _surface_mesh(origin, u_dir, v_dir, nu: int, nv: int, base: int):
    """add_surface_mesh (geometry.rs:434-469): (nu+1)(nv+1) nodes row by row, nu*nv quads (n0, n1, n2, n3)."""
    origin, u_dir, v_dir = (np.asarray(a, dtype=np.float64) for a in (origin, u_dir, v_dir))
    u = np.arange(nu + 1, dtype=np.float64) / nu
    v = np.arange(nv + 1, dtype=np.float64) / nv
    uu, vv = np.meshgrid(u, v)  # rows = j (v), columns = i (u)
    # origin + u*(u_dir - origin) + v*(v_dir - origin), evaluated in the reference's order
    nodes = (origin[None, None, :] + uu[:, :, None] * (u_dir - origin)[None, None, :]) + vv[:, :, None] * (v_dir - origin)[None, None, :]
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    n0 = base + jj * (nu + 1) + ii
    quads = np.stack([n0, n0 + 1, n0 + (nu + 1) + 1, n0 + (nu + 1)], axis=-1).reshape(-1, 4)
    return nodes.reshape(-1, 3), quads


@dataclass
class RectangularRoom:
    """geometry.rs:86-183: width (x), depth (y), height (z), one corner at the origin."""
    width: float
    depth: float
    height: float

    def generate_mesh(self, elements_per_meter: int) -> RoomMesh:
        nx = int(math.ceil(self.width * elements_per_meter))
        ny = int(math.ceil(self.depth * elements_per_meter))
        nz = int(math.ceil(self.height * elements_per_meter))
        w, d, h = self.width, self.depth, self.height
        walls = [  # (origin, u_dir, v_dir, nu, nv): floor, ceiling, front, back, left, right
            ((0, 0, 0), (w, 0, 0), (0, d, 0), nx, ny),
            ((0, 0, h), (w, 0, h), (0, d, h), nx, ny),
            ((0, 0, 0), (w, 0, 0), (0, 0, h), nx, nz),
            ((0, d, 0), (w, d, 0), (0, d, h), nx, nz),
            ((0, 0, 0), (0, d, 0), (0, 0, h), ny, nz),
            ((w, 0, 0), (w, d, 0), (w, 0, h), ny, nz),
        ]
        nodes, elems, base = [], [], 0
        for o, u, v, nu, nv in walls:
            nd, q = _surface_mesh(o, u, v, nu, nv, base)
            nodes.append(nd)
            elems.append(q)
            base += nd.shape[0]
        return RoomMesh(np.ascontiguousarray(np.concatenate(nodes)), np.ascontiguousarray(np.concatenate(elems).astype(np.uint32)))

    def dimensions(self):
        return self.width, self.depth, self.height

    def volume(self) -> float:
        return self.width * self.depth * self.height


@dataclass
class LShapedRoom:
    """geometry.rs:474-708: main section width1 x depth1, extension width2 x depth2 behind it (x in [0, width2]),
    common height.  Ten wall patches, meshed independently (nodes are not shared between patches)."""
    width1: float
    depth1: float
    width2: float
    depth2: float
    height: float

    def generate_mesh(self, elements_per_meter: int) -> RoomMesh:
        e = elements_per_meter
        nx1 = int(math.ceil(self.width1 * e)); ny1 = int(math.ceil(self.depth1 * e))
        nx2 = int(math.ceil(self.width2 * e)); ny2 = int(math.ceil(self.depth2 * e))
        nz = int(math.ceil(self.height * e))
        w1, d1, w2, d2, h = self.width1, self.depth1, self.width2, self.depth2, self.height
        td = d1 + d2
        ny_total = int(math.ceil(td * e))
        nx_int = int(math.ceil((w1 - w2) * e))
        walls = [
            ((0, 0, 0), (w1, 0, 0), (0, d1, 0), nx1, ny1),        # floor, main
            ((0, d1, 0), (w2, d1, 0), (0, td, 0), nx2, ny2),      # floor, extension
            ((0, 0, h), (w1, 0, h), (0, d1, h), nx1, ny1),        # ceiling, main
            ((0, d1, h), (w2, d1, h), (0, td, h), nx2, ny2),      # ceiling, extension
            ((0, 0, 0), (w1, 0, 0), (0, 0, h), nx1, nz),          # front wall y = 0
            ((w1, 0, 0), (w1, d1, 0), (w1, 0, h), ny1, nz),       # right wall of the main section
            ((0, 0, 0), (0, td, 0), (0, 0, h), ny_total, nz),     # left wall x = 0
            ((0, td, 0), (w2, td, 0), (0, td, h), nx2, nz),       # back wall of the extension
            ((w2, d1, 0), (w2, td, 0), (w2, d1, h), ny2, nz),     # right wall of the extension
            ((w2, d1, 0), (w1, d1, 0), (w2, d1, h), nx_int, nz),  # internal wall at the junction
        ]
        nodes, elems, base = [], [], 0
        for o, u, v, nu, nv in walls:
            nd, q = _surface_mesh(o, u, v, nu, nv, base)
            nodes.append(nd)
            elems.append(q)
            base += nd.shape[0]
        return RoomMesh(np.ascontiguousarray(np.concatenate(nodes)), np.ascontiguousarray(np.concatenate(elems).astype(np.uint32)))

    def dimensions(self):
        return max(self.width1, self.width2), self.depth1 + self.depth2, self.height

    def volume(self) -> float:
        return (self.width1 * self.depth1 + self.width2 * self.depth2) * self.height


# ---- source.rs -----------------------------------------------------------------------------
@dataclass
class DirectivityPattern:
    """source.rs:9-98: magnitude[n_vertical, n_horizontal] sampled every 10 degrees."""
    horizontal_angles: np.ndarray
    vertical_angles: np.ndarray
    magnitude: np.ndarray

    @staticmethod
    def omnidirectional() -> "DirectivityPattern":
        h = np.arange(36) * 10.0
        v = np.arange(19) * 10.0
        return DirectivityPattern(h, v, np.ones((19, 36)))

    @staticmethod
    def cardioid() -> "DirectivityPattern":
        h = np.arange(36) * 10.0
        v = np.arange(19) * 10.0
        mag = np.zeros((19, 36))
        for vi, va in enumerate(v):
            for hi, ha in enumerate(h):
                forward_dot = math.sin(math.radians(va)) * math.sin(math.radians(ha))
                mag[vi, hi] = 0.5 * max(1.0 + forward_dot, 0.0)
        return DirectivityPattern(h, v, mag)

    def interpolate(self, theta: float, phi: float) -> float:
        theta_deg = math.degrees(theta)
        phi_deg = math.degrees(phi)
        while phi_deg < 0.0:
            phi_deg += 360.0
        while phi_deg >= 360.0:
            phi_deg -= 360.0
        nh, nv = len(self.horizontal_angles), len(self.vertical_angles)
        h_idx = min(int(math.floor(phi_deg / 10.0)), nh - 1)
        v_idx = min(int(math.floor(theta_deg / 10.0)), nv - 1)
        h_next = (h_idx + 1) % nh
        v_next = min(v_idx + 1, nv - 1)
        h_frac = phi_deg / 10.0 - h_idx
        v_frac = theta_deg / 10.0 - v_idx
        m = self.magnitude
        m0 = m[v_idx, h_idx] * (1.0 - h_frac) + m[v_idx, h_next] * h_frac
        m1 = m[v_next, h_idx] * (1.0 - h_frac) + m[v_next, h_next] * h_frac
        return float(m0 * (1.0 - v_frac) + m1 * v_frac)

    def is_omnidirectional(self) -> bool:
        return bool(np.all(self.magnitude == 1.0))


@dataclass
class CrossoverFilter:
    """source.rs:103-155.  kind: 'fullrange' | 'lowpass' | 'highpass' | 'bandpass'."""
    kind: str = "fullrange"
    cutoff_freq: float = 0.0
    low_cutoff: float = 0.0
    high_cutoff: float = 0.0
    order: int = 2

    @staticmethod
    def full_range():
        return CrossoverFilter()

    @staticmethod
    def lowpass(cutoff_freq: float, order: int):
        return CrossoverFilter("lowpass", cutoff_freq=cutoff_freq, order=order)

    @staticmethod
    def highpass(cutoff_freq: float, order: int):
        return CrossoverFilter("highpass", cutoff_freq=cutoff_freq, order=order)

    @staticmethod
    def bandpass(low_cutoff: float, high_cutoff: float, order: int):
        return CrossoverFilter("bandpass", low_cutoff=low_cutoff, high_cutoff=high_cutoff, order=order)

    def amplitude_at_frequency(self, frequency: float) -> float:
        if self.kind == "fullrange":
            return 1.0
        p = self.order * 2
        if self.kind == "lowpass":
            return 1.0 / math.sqrt(1.0 + (frequency / self.cutoff_freq) ** p)
        if self.kind == "highpass":
            return 1.0 / math.sqrt(1.0 + (self.cutoff_freq / frequency) ** p)
        if self.kind == "bandpass":
            hp = 1.0 / math.sqrt(1.0 + (self.low_cutoff / frequency) ** p)
            lp = 1.0 / math.sqrt(1.0 + (frequency / self.high_cutoff) ** p)
            return hp * lp
        raise ValueError(f"unknown crossover kind {self.kind!r}")


@dataclass
class Source:
    """source.rs:160-219."""
    position: Sequence[float]
    directivity: DirectivityPattern = field(default_factory=DirectivityPattern.omnidirectional)
    amplitude: float = 1.0
    crossover: CrossoverFilter = field(default_factory=CrossoverFilter.full_range)
    name: str = "Source"

    @staticmethod
    def omnidirectional(position, amplitude: float) -> "Source":
        return Source(position, DirectivityPattern.omnidirectional(), amplitude)

    def with_crossover(self, crossover: CrossoverFilter) -> "Source":
        self.crossover = crossover
        return self

    def amplitude_towards(self, point, frequency: float) -> float:
        dx, dy, dz = (float(point[i]) - float(self.position[i]) for i in range(3))
        r = math.sqrt(dx * dx + dy * dy + dz * dz)
        if r < 1e-10:
            return self.amplitude * self.crossover.amplitude_at_frequency(frequency)
        theta = math.acos(dz / r)
        phi = math.atan2(dy, dx)
        return self.amplitude * self.directivity.interpolate(theta, phi) * self.crossover.amplitude_at_frequency(frequency)


def _csources(sources: Sequence[Source], frequency: float):
    """Source -> bemb200_room_source[]: amplitude x crossover folded on the host (a scalar per source and frequency)."""
    arr = (_capi.CRoomSource * len(sources))()
    keep = []
    for i, s in enumerate(sources):
        arr[i].position = (C.c_double * 3)(*[float(v) for v in s.position])
        arr[i].amplitude = s.amplitude * s.crossover.amplitude_at_frequency(frequency)
        if s.directivity.is_omnidirectional():
            arr[i].directivity = None
            arr[i].n_horizontal = arr[i].n_vertical = 0
        else:
            tab = np.ascontiguousarray(s.directivity.magnitude, dtype=np.float64)
            keep.append(tab)
            arr[i].directivity = tab.ctypes.data
            arr[i].n_horizontal = tab.shape[1]
            arr[i].n_vertical = tab.shape[0]
    return arr, keep


# ---- solver.rs -----------------------------------------------------------------------------
class StagedRoomMesh:
    """RoomMesh resident on the device with element_center_and_normal / element_area (solver.rs:38-122) evaluated there."""

    def __init__(self, mesh: RoomMesh, ctx: Optional[bem.Context] = None):
        self.ctx = ctx or bem.default_context()
        self.mesh = mesh
        nodes = np.ascontiguousarray(mesh.nodes, dtype=np.float64)
        el = np.asarray(mesh.elements)
        if el.ndim != 2 or el.shape[1] not in (3, 4):
            raise ValueError("elements must be (n, 3) or (n, 4) node indices")
        conn = np.full((el.shape[0], 4), PAD, dtype=np.uint32)
        conn[:, :el.shape[1]] = el.astype(np.uint32)
        self._lib = _capi.lib()
        self._h = C.c_void_p()
        _capi.check(self._lib.bemb200_room_mesh_stage(self.ctx._h, _capi.ptr(nodes), nodes.shape[0], _capi.ptr(conn), conn.shape[0],
                                                      C.byref(self._h)), self.ctx._h)
        self.n = conn.shape[0]

    def geometry(self):
        c = np.empty((self.n, 3)); nm = np.empty((self.n, 3)); a = np.empty(self.n)
        _capi.check(self._lib.bemb200_room_mesh_geometry(self._h, _capi.ptr(c), _capi.ptr(nm), _capi.ptr(a)), self.ctx._h)
        return c, nm, a

    def __del__(self):
        try:
            if self._h:
                self._lib.bemb200_room_mesh_free(self._h)
                self._h = None
        except Exception:
            pass


def _staged(mesh, ctx=None) -> StagedRoomMesh:
    return mesh if isinstance(mesh, StagedRoomMesh) else StagedRoomMesh(mesh, ctx)


def build_bem_matrix_parallel(mesh, k: float, ctx: Optional[bem.Context] = None, rows=None,
                              reuse: Optional[bem.DeviceMatrix] = None, stats: Optional[dict] = None) -> bem.DeviceMatrix:
    """solver.rs:448-493.  Returns the device-resident matrix (``.rows()`` downloads it); ``rows=(r0, r1)`` assembles a
    row block (default: this rank's canonical block)."""
    st = _staged(mesh, ctx)
    ctx = st.ctx
    r0, r1 = rows if rows is not None else ctx.partition(st.n)
    h = reuse._h if reuse is not None else C.c_void_p()
    ms = C.c_double(0.0)
    _capi.check(_capi.lib().bemb200_room_assemble(ctx._h, st._h, float(k), r0, r1, C.byref(h), C.byref(ms)), ctx._h)
    if stats is not None:
        stats["kernel_ms"] = ms.value
    return reuse if reuse is not None else bem.DeviceMatrix(ctx, h)


def calculate_incident_field_derivative_parallel(mesh, sources: Sequence[Source], k: float, frequency: float,
                                                 ctx: Optional[bem.Context] = None) -> np.ndarray:
    """solver.rs:638-679."""
    st = _staged(mesh, ctx)
    arr, keep = _csources(sources, frequency)
    out = np.empty(st.n, dtype=np.complex128)
    _capi.check(_capi.lib().bemb200_room_incident_rhs(st._h, float(k), len(sources), arr, _capi.ptr(out), None), st.ctx._h)
    return out


def solve_bem_system(mesh, sources: Sequence[Source], k: float, frequency: float, ctx: Optional[bem.Context] = None,
                     reuse: Optional[bem.DeviceMatrix] = None, info: Optional[dict] = None) -> np.ndarray:
    """solver.rs:412-445: matrix + incident right-hand side + GMRES(max_iterations=100 cycles, restart=50, tol=1e-6)
    through DenseOperator; returns solution.x."""
    st = _staged(mesh, ctx)
    stats: dict = {}
    matrix = build_bem_matrix_parallel(st, k, reuse=reuse, stats=stats)
    rhs = calculate_incident_field_derivative_parallel(st, sources, k, frequency)
    config = bem.GmresConfig(max_iterations=100, restart=50, tolerance=1e-6)
    solution = bem.solve_gmres(bem.DenseOperator(matrix), rhs, config)
    if info is not None:
        info.update(iterations=solution.iterations, restarts=solution.restarts, residual=solution.residual,
                    converged=solution.converged, matrix=matrix, assembly_kernel_ms=stats.get("kernel_ms"))
    return solution.x


def calculate_field_pressure_bem_parallel(mesh, surface_pressure: np.ndarray, sources: Sequence[Source], field_points,
                                          k: float, frequency: float, ctx: Optional[bem.Context] = None) -> np.ndarray:
    """solver.rs:687-748."""
    st = _staged(mesh, ctx)
    ps = np.ascontiguousarray(surface_pressure, dtype=np.complex128)
    if ps.shape != (st.n,):
        raise ValueError("surface_pressure must have one entry per element")
    pts = np.ascontiguousarray(field_points, dtype=np.float64).reshape(-1, 3)
    arr, keep = _csources(sources, frequency)
    out = np.empty(pts.shape[0], dtype=np.complex128)
    _capi.check(_capi.lib().bemb200_room_field_pressure(st._h, float(k), len(sources), arr, pts.shape[0], _capi.ptr(pts),
                                                        _capi.ptr(ps), _capi.ptr(out)), st.ctx._h)
    return out


# ---- bin/room_simulator_bem.rs ----------------------------------------------------------------
@dataclass
class RoomSimulation:
    """What run_direct_gmres reads from RoomSimulation (math-xem-common/src/config.rs): room, sources, listening
    positions, frequency grid, speed of sound."""
    room: object  # RectangularRoom | LShapedRoom (RoomGeometry, geometry.rs:9-14)
    sources: List[Source]
    listening_positions: List[Sequence[float]]
    frequencies: List[float]
    speed_of_sound: float = SPEED_OF_SOUND_20C

    def wavenumber(self, frequency: float) -> float:
        return wavenumber(frequency, self.speed_of_sound)


def simulation_from_config(config) -> tuple:
    """RoomConfig (math-xem-common/src/config.rs:12-35, the JSON files under math-bem/configs/) -> (RoomSimulation,
    mesh_resolution).  ``config``: dict or path of a JSON file.  Rectangular and L-shaped rooms; sources keep their directivity (omnidirectional / custom) and crossover (config.rs:209-340)."""
    import json
    from pathlib import Path

    if not isinstance(config, dict):
        config = json.loads(Path(config).read_text())
    rc = config["room"]
    if rc.get("type") == "rectangular":
        room = RectangularRoom(float(rc["width"]), float(rc["depth"]), float(rc["height"]))
    elif rc.get("type") == "lshaped":
        room = LShapedRoom(float(rc["width1"]), float(rc["depth1"]), float(rc["width2"]), float(rc["depth2"]), float(rc["height"]))
    else:
        raise ValueError(f"unknown room type {rc.get('type')!r}")
    sources = []
    for sc in config["sources"]:
        dc = sc.get("directivity", {"type": "omnidirectional"})
        if dc.get("type", "omnidirectional") == "omnidirectional":
            pattern = DirectivityPattern.omnidirectional()
        elif dc["type"] == "custom":
            mag = np.asarray(dc["magnitude"], dtype=np.float64)
            if mag.size == 0:
                raise ValueError("Empty magnitude array")
            if mag.shape[0] != len(dc["vertical_angles"]) or mag.shape[1] != len(dc["horizontal_angles"]):
                raise ValueError("directivity angles / magnitude shape mismatch")
            pattern = DirectivityPattern(np.asarray(dc["horizontal_angles"], float), np.asarray(dc["vertical_angles"], float), mag)
        else:
            raise ValueError(f"unknown directivity type {dc['type']!r}")
        xc = sc.get("crossover", {"type": "fullrange"})
        kind = xc.get("type", "fullrange")
        if kind == "fullrange":
            xo = CrossoverFilter.full_range()
        elif kind == "lowpass":
            xo = CrossoverFilter.lowpass(float(xc["cutoff_freq"]), int(xc["order"]))
        elif kind == "highpass":
            xo = CrossoverFilter.highpass(float(xc["cutoff_freq"]), int(xc["order"]))
        elif kind == "bandpass":
            xo = CrossoverFilter.bandpass(float(xc["low_cutoff"]), float(xc["high_cutoff"]), int(xc["order"]))
        else:
            raise ValueError(f"unknown crossover type {kind!r}")
        pos = sc["position"]
        sources.append(Source([float(pos["x"]), float(pos["y"]), float(pos["z"])], pattern, float(sc.get("amplitude", 1.0)), xo,
                              sc.get("name", "Source")))
    lps = [[float(p["x"]), float(p["y"]), float(p["z"])] for p in config["listening_positions"]]
    fc = config["frequencies"]
    gen = lin_space if str(fc.get("spacing", "logarithmic")).lower() == "linear" else log_space
    freqs = gen(float(fc["min_freq"]), float(fc["max_freq"]), int(fc["num_points"]))
    sim = RoomSimulation(room, sources, lps, freqs, float(config.get("speed_of_sound", SPEED_OF_SOUND_20C)))
    return sim, int(config.get("solver", {}).get("mesh_resolution", 2))


def run_direct_gmres(simulation: RoomSimulation, mesh_resolution: int, ctx: Optional[bem.Context] = None,
                     details: Optional[list] = None) -> List[float]:
    """room_simulator_bem.rs:225-281 ("direct" mode): per frequency solve_bem_system + field pressure at the first
    listening position -> SPL.  The mesh is staged once and the matrix buffer reused over the sweep."""
    mesh = simulation.room.generate_mesh(mesh_resolution)
    st = StagedRoomMesh(mesh, ctx)
    lp = np.asarray(simulation.listening_positions[0], dtype=np.float64).reshape(1, 3)
    spl = []
    matrix = None
    for freq in simulation.frequencies:
        k = simulation.wavenumber(freq)
        info: dict = {}
        x = solve_bem_system(st, simulation.sources, k, freq, reuse=matrix, info=info)
        matrix = info.pop("matrix")
        p = calculate_field_pressure_bem_parallel(st, x, simulation.sources, lp, k, freq)
        spl.append(pressure_to_spl(p[0]))
        if details is not None:
            details.append(dict(frequency=freq, spl=spl[-1], **info))
    return spl
