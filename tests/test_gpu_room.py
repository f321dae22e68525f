"""GPU parity of the room-acoustics dense path through the C ABI
(math_audio_b200.room -> ctypes -> libbemb200.so) against oracle/room_oracle.py.

Tolerances: element geometry bit-exact; matrix / right-hand side / field entries relative 1e-10
(row-normwise where the double-layer kernel vanishes: coplanar elements have cos = 0); solution
relative 1e-8 with identical GMRES iteration counts (the reference's own settings: restart 50,
tol 1e-6, 100 cycles); SPL within 1e-6 dB.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def room():
    from math_audio_b200 import bem, room as r

    bem.default_context()
    return r


def _osrc(s, xo):
    return dict(position=[float(v) for v in s.position], amplitude=s.amplitude,
                directivity=None if s.directivity.is_omnidirectional() else s.directivity.magnitude.tolist(), crossover=xo)


def _sources(room):
    srcs = [room.Source.omnidirectional([1.2, 0.9, 1.1], 1.0),
            room.Source([3.9, 0.7, 0.4], room.DirectivityPattern.cardioid(), 0.8, room.CrossoverFilter.lowpass(120.0, 4), "Sub")]
    osrcs = [_osrc(srcs[0], dict(kind="fullrange")), _osrc(srcs[1], dict(kind="lowpass", cutoff=120.0, order=4))]
    return srcs, osrcs


def test_room_geometry_bit_exact_incl_triangles(room):
    from oracle import room_oracle as ro

    mesh = room.RectangularRoom(3.3, 2.1, 2.4).generate_mesh(3)
    st = room.StagedRoomMesh(mesh)
    c, n, a = st.geometry()
    co, no, ao = ro.element_data(mesh.nodes, mesh.elements)
    assert np.array_equal(c, co) and np.array_equal(n, no) and np.array_equal(a, ao)
    # triangles + a skewed quad (element_area's two-triangle split)
    nodes = np.array([[0, 0, 0], [1, 0, 0], [0.5, 1, 0], [1.7, 1.2, 0.3], [0.1, 0.2, 1.0], [2.0, -0.3, 0.9]], dtype=float)
    elems = np.array([[0, 1, 2, room.PAD], [1, 3, 2, room.PAD], [0, 1, 3, 4], [2, 3, 5, 4]], dtype=np.uint32)
    st2 = room.StagedRoomMesh(room.RoomMesh(nodes, elems))
    c, n, a = st2.geometry()
    co, no, ao = ro.element_data(nodes, elems)
    assert np.array_equal(c, co) and np.array_equal(n, no) and np.array_equal(a, ao)


@pytest.mark.parametrize("freq", [63.0, 250.0, 1000.0])
def test_room_matrix_rhs_field_parity(room, freq):
    from oracle import room_oracle as ro

    mesh = room.RectangularRoom(5.0, 4.0, 2.5).generate_mesh(4)   # 1 360 quads
    st = room.StagedRoomMesh(mesh)
    k = room.wavenumber(freq, 343.0)
    c, n, a = ro.element_data(mesh.nodes, mesh.elements)
    Aref = ro.build_bem_matrix(c, n, a, k)
    A = room.build_bem_matrix_parallel(st, k).rows()
    scale = np.max(np.abs(Aref), axis=1, keepdims=True)
    assert np.max(np.abs(A - Aref) / scale) < 1e-13
    big = np.abs(Aref) > 1e-9 * scale
    assert np.max(np.abs(A - Aref)[big] / np.abs(Aref)[big]) < 1e-10
    assert np.array_equal(np.diag(A), np.diag(Aref))
    # a row block equals the corresponding rows
    Ablk = room.build_bem_matrix_parallel(st, k, rows=(100, 333)).rows()
    assert np.array_equal(Ablk, A[100:333])
    srcs, osrcs = _sources(room)
    rhs = room.calculate_incident_field_derivative_parallel(st, srcs, k, freq)
    rref = ro.incident_field_derivative(c, n, osrcs, k, freq)
    assert np.max(np.abs(rhs - rref)) / np.max(np.abs(rref)) < 1e-12
    rng = np.random.default_rng(3)
    p = rng.standard_normal(len(a)) + 1j * rng.standard_normal(len(a))
    pts = np.column_stack([rng.uniform(0.3, 4.7, 40), rng.uniform(0.3, 3.7, 40), rng.uniform(0.3, 2.2, 40)])
    got = room.calculate_field_pressure_bem_parallel(st, p, srcs, pts, k, freq)
    ref = ro.field_pressure(c, n, a, p, osrcs, pts, k, freq)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-10


def test_room_solve_and_spl_equal_oracle(room, orc):
    from oracle import room_oracle as ro

    sim = room.RoomSimulation(room.RectangularRoom(4.2, 3.1, 2.4), _sources(room)[0], [[2.9, 2.0, 1.2]],
                              room.log_space(40.0, 400.0, 5))
    res = 4
    details = []
    spl = room.run_direct_gmres(sim, res, details=details)
    mesh = sim.room.generate_mesh(res)
    c, n, a = ro.element_data(mesh.nodes, mesh.elements)
    osrcs = _sources(room)[1]
    for f, s, d in zip(sim.frequencies, spl, details):
        k = room.wavenumber(f, 343.0)
        A = ro.build_bem_matrix(c, n, a, k)
        b = ro.incident_field_derivative(c, n, osrcs, k, f)
        xo, io = orc.gmres(A, b, max_iterations=100, restart=50, tolerance=1e-6)
        assert d["iterations"] == io["iterations"] and d["converged"] == bool(io["converged"])
        p = ro.field_pressure(c, n, a, xo, osrcs, np.array(sim.listening_positions), k, f)
        assert abs(s - ro.pressure_to_spl(p[0])) < 1e-6
    # and the solution itself at one frequency
    f = sim.frequencies[2]
    k = room.wavenumber(f, 343.0)
    x = room.solve_bem_system(mesh, sim.sources, k, f)
    xo, _ = orc.gmres(ro.build_bem_matrix(c, n, a, k), ro.incident_field_derivative(c, n, osrcs, k, f), max_iterations=100,
                      restart=50, tolerance=1e-6)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-8


def test_room_invalid_inputs(room):
    from math_audio_b200 import _capi

    mesh = room.RectangularRoom(1.0, 1.0, 1.0).generate_mesh(2)
    st = room.StagedRoomMesh(mesh)
    with pytest.raises(_capi.Bemb200Error):
        room.build_bem_matrix_parallel(st, -1.0)
    with pytest.raises(_capi.Bemb200Error):
        room.build_bem_matrix_parallel(st, 1.0, rows=(0, st.n + 1))
    bad = room.RoomMesh(mesh.nodes, np.array([[0, 1, 2, 10 ** 6]], dtype=np.uint32))
    with pytest.raises(_capi.Bemb200Error):
        room.StagedRoomMesh(bad)
    with pytest.raises(ValueError):
        room.calculate_field_pressure_bem_parallel(st, np.zeros(3, dtype=complex), [], [[0.5, 0.5, 0.5]], 1.0, 50.0)
