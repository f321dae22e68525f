"""Room-acoustics dense path, CPU side: the reference's own tests for this path restated on the
oracle (oracle/room_oracle.py) and on the host-side mirror (math_audio_b200/room.py), plus the
mirror's mesh / source model against the oracle's literal loops.  No GPU."""
import math

import numpy as np
import pytest

from math_audio_b200 import room
from oracle import room_oracle as ro


# ---- math-bem/src/room_acoustics/solver.rs:1155-1170 ------------------------------------------
def test_greens_function():
    k = 2.0 * math.pi * 1000.0 / 343.0
    g = ro.greens_function_3d(1.0, k)
    assert abs(abs(g) - 1.0 / (4.0 * math.pi)) < 0.1
    assert ro.greens_function_3d(1e-11, k) == 0 and ro.greens_function_derivative(1e-11, k, 1.0) == 0


def test_pressure_to_spl():
    for fn in (ro.pressure_to_spl, room.pressure_to_spl):
        assert abs(fn(1.0 + 0j) - 94.0) < 1.0       # 1 Pa = 94 dB SPL
        assert fn(0j) == -120.0


# ---- math-xem-common/src/geometry.rs:765-779 --------------------------------------------------
def test_rectangular_room_mesh():
    r = room.RectangularRoom(2.0, 2.0, 2.0)
    m = r.generate_mesh(1)
    assert m.num_nodes() > 0 and m.num_elements() > 0
    assert (r.width, r.depth, r.height) == (2.0, 2.0, 2.0)
    assert m.num_elements() == 6 * 4 and m.num_nodes() == 6 * 9


@pytest.mark.parametrize("dims,epm", [((5.0, 4.0, 2.5), 2), ((3.3, 2.1, 2.4), 3), ((1.0, 1.0, 1.0), 1)])
def test_room_mesh_equals_oracle_loops(dims, epm):
    m = room.RectangularRoom(*dims).generate_mesh(epm)
    nodes, elems = ro.rectangular_room_mesh(*dims, epm)
    assert m.nodes.shape == nodes.shape and np.array_equal(m.nodes, nodes)      # bit-identical coordinates
    assert np.array_equal(m.elements.astype(np.int64), elems)
    # closed box: areas sum to the surface, normals are axis aligned
    c, n, a = ro.element_data(nodes, elems)
    w, d, h = dims
    assert abs(a.sum() - 2 * (w * d + w * h + d * h)) < 1e-12
    assert np.allclose(np.abs(n).max(axis=1), 1.0)


def test_lshaped_room_mesh_equals_oracle_loops():
    r = room.LShapedRoom(5.0, 4.0, 3.0, 3.0, 2.5)                       # geometry.rs:781-788
    assert (r.width1, r.depth1, r.width2, r.depth2, r.height) == (5.0, 4.0, 3.0, 3.0, 2.5)
    for epm in (1, 3):
        m = r.generate_mesh(epm)
        nodes, elems = ro.lshaped_room_mesh(5.0, 4.0, 3.0, 3.0, 2.5, epm)
        assert np.array_equal(m.nodes, nodes) and np.array_equal(m.elements.astype(np.int64), elems)
    c, n, a = ro.element_data(nodes, elems)
    # ten patches: two floors, two ceilings, six vertical walls (the junction wall closes the L)
    floor = 5.0 * 4.0 + 3.0 * 3.0
    walls = (5.0 + 4.0 + 7.0 + 3.0 + 3.0 + 2.0) * 2.5
    assert abs(a.sum() - (2 * floor + walls)) < 1e-12
    sim, res = room.simulation_from_config(dict(room=dict(type="lshaped", width1=5.0, depth1=4.0, width2=3.0, depth2=3.0, height=2.6),
                                                sources=[dict(position=dict(x=1, y=1, z=1))], listening_positions=[dict(x=2, y=2, z=1)],
                                                frequencies=dict(min_freq=40.0, max_freq=500.0, num_points=4), solver=dict(mesh_resolution=3)))
    assert isinstance(sim.room, room.LShapedRoom) and res == 3 and len(sim.frequencies) == 4 and sim.sources[0].amplitude == 1.0


def test_simulation_from_config_rectangular():
    cfg = dict(room=dict(type="rectangular", width=5.5, depth=7.0, height=2.6),
               sources=[dict(name="Sub", position=dict(x=0.5, y=0.5, z=0.3), amplitude=1.0, directivity=dict(type="omnidirectional"),
                             crossover=dict(type="lowpass", cutoff_freq=80.0, order=4)),
                        dict(name="Main", position=dict(x=1.2, y=0.3, z=1.1), amplitude=0.8,
                             directivity=dict(type="custom", horizontal_angles=[0.0, 10.0], vertical_angles=[0.0, 10.0, 20.0],
                                              magnitude=[[1.0, 0.5], [0.9, 0.4], [0.8, 0.3]]),
                             crossover=dict(type="bandpass", low_cutoff=60.0, high_cutoff=3000.0, order=2))],
               listening_positions=[dict(x=2.75, y=4.5, z=1.2)],
               frequencies=dict(min_freq=20.0, max_freq=300.0, num_points=100, spacing="linear"), solver=dict(method="gmres"))
    sim, res = room.simulation_from_config(cfg)
    assert res == 2                                                      # default_mesh_resolution (config.rs:413-415)
    assert sim.frequencies[1] - sim.frequencies[0] == pytest.approx(280.0 / 99)
    assert sim.sources[0].crossover.kind == "lowpass" and sim.sources[0].crossover.order == 4
    assert sim.sources[1].directivity.magnitude.shape == (3, 2) and not sim.sources[1].directivity.is_omnidirectional()
    assert sim.speed_of_sound == 343.0
    with pytest.raises(ValueError):
        room.simulation_from_config(dict(cfg, room=dict(type="dome")))
    bad = dict(cfg, sources=[dict(position=dict(x=0, y=0, z=0), directivity=dict(type="custom", horizontal_angles=[0.0], vertical_angles=[0.0],
                                                                          magnitude=[[1.0, 2.0]]))])
    with pytest.raises(ValueError):
        room.simulation_from_config(bad)                                 # config.rs:246-259 angle / magnitude mismatch


# ---- math-xem-common/src/source.rs:228-257 ----------------------------------------------------
def test_omnidirectional_pattern():
    p = room.DirectivityPattern.omnidirectional()
    for th, ph in [(0.0, 0.0), (math.pi / 2, math.pi), (math.pi, 0.0)]:
        assert abs(p.interpolate(th, ph) - 1.0) < 1e-6
        assert abs(ro.directivity_interpolate(ro.directivity_omnidirectional(), th, ph) - 1.0) < 1e-6


def test_crossover_lowpass():
    x = room.CrossoverFilter.lowpass(100.0, 2)
    assert abs(x.amplitude_at_frequency(10.0) - 1.0) < 0.1
    assert 0.6 < x.amplitude_at_frequency(100.0) < 0.8
    assert x.amplitude_at_frequency(1000.0) < 0.1
    assert x.amplitude_at_frequency(100.0) == ro.crossover_amplitude("lowpass", 100.0, cutoff=100.0, order=2)


def test_source_amplitude():
    s = room.Source.omnidirectional([0.0, 0.0, 0.0], 1.0)
    assert abs(s.amplitude_towards([1.0, 0.0, 0.0], 1000.0) - 1.0) < 1e-6


def test_source_model_equals_oracle():
    rng = np.random.default_rng(5)
    card = room.DirectivityPattern.cardioid()
    assert np.array_equal(card.magnitude, np.array(ro.directivity_cardioid()))
    for xo, oxo in [(room.CrossoverFilter.full_range(), dict(kind="fullrange")),
                    (room.CrossoverFilter.highpass(80.0, 4), dict(kind="highpass", cutoff=80.0, order=4)),
                    (room.CrossoverFilter.bandpass(60.0, 3000.0, 2), dict(kind="bandpass", low=60.0, high=3000.0, order=2))]:
        s = room.Source([1.0, 2.0, 0.5], card, 0.7, xo)
        o = dict(position=[1.0, 2.0, 0.5], amplitude=0.7, directivity=ro.directivity_cardioid(), crossover=oxo)
        for _ in range(50):
            p = rng.uniform(-3, 5, 3)
            f = float(rng.uniform(20, 5000))
            assert s.amplitude_towards(p, f) == pytest.approx(ro.amplitude_towards(o, p, f), rel=1e-15, abs=0)
        assert s.amplitude_towards([1.0, 2.0, 0.5], 100.0) == pytest.approx(0.7 * xo.amplitude_at_frequency(100.0))


# ---- math-xem-common/src/types.rs:336-354 -----------------------------------------------------
def test_log_space_and_wavenumber():
    f = room.log_space(20.0, 20000.0, 200)
    assert len(f) == 200 and abs(f[0] - 20.0) < 1e-6 and abs(f[199] - 20000.0) < 1e-6 and f[1] / f[0] > 1.0
    assert f == ro.log_space(20.0, 20000.0, 200)
    assert room.log_space(5.0, 9.0, 1) == [5.0]
    assert abs(room.wavenumber(1000.0, 343.0) - 2.0 * math.pi * 1000.0 / 343.0) < 1e-10
    assert room.lin_space(1.0, 2.0, 3) == [1.0, 1.5, 2.0]


# ---- oracle self-consistency ------------------------------------------------------------------
def test_oracle_matrix_vectorised_equals_loops():
    nodes, elems = ro.rectangular_room_mesh(1.2, 1.0, 0.8, 3)
    c, n, a = ro.element_data(nodes, elems)
    k = 7.3
    A1 = ro.build_bem_matrix_loops(c, n, a, k)
    A2 = ro.build_bem_matrix(c, n, a, k)
    assert np.max(np.abs(A1 - A2)) <= 1e-15 * np.max(np.abs(A1))
    assert np.array_equal(np.diag(A2), 1j * (-k / (2 * math.pi)) * a)
    # coplanar elements see each other with cos = 0: the double-layer kernel vanishes inside a wall
    same_wall = (np.abs(n @ n.T) > 0.5) & (np.abs((c[:, None, :] - c[None, :, :]) * n[:, None, :]).sum(-1) < 1e-12)
    off = same_wall & ~np.eye(len(a), dtype=bool)
    assert np.max(np.abs(A2[off])) < 1e-15
    assert np.array_equal(ro.build_bem_matrix(c, n, a, k, rows=(5, 17)), A2[5:17])


def test_room_calls_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from math_audio_b200 import _capi

    with pytest.raises((_capi.Bemb200Error, FileNotFoundError)):
        room.build_bem_matrix_parallel(room.RectangularRoom(1, 1, 1).generate_mesh(1), 1.0)
