#!/usr/bin/env python3
"""Room-acoustics dense path measured on one B200: the reference's "direct
GMRES" mode (math-bem/bin/room_simulator_bem.rs:225-281) on a home-theatre configuration
restated from math-bem/configs/home_theater_2_1.json (5.5 x 7.0 x 2.6 m, two mains high-passed at
80 Hz + a subwoofer low-passed at 80 Hz, listening position (2.75, 4.5, 1.2), 20-300 Hz log grid,
mesh_resolution 10 -> 14 200 Quad4 elements).

    python tests/drivers/run_room.py [--frequencies 12] [--mesh-resolution 10] [--rows-checked 48]

Per frequency: build_bem_matrix_parallel + incident right-hand side + GMRES(restart 50, tol 1e-6,
100 cycles) + field pressure at the listening position -> SPL.  Reports seconds per frequency,
the matrix kernel against the HBM write roofline (16 B per entry), the ZGEMV bandwidth, sampled-row
/ right-hand-side / SPL parity against oracle/room_oracle.py, and a bounded CPU baseline (the
oracle's vectorised numpy matrix on a row slab + the C++ oracle's threaded zgemv, extrapolated).
Writes gpurun_out/room_home_theater.json.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

CONFIG = {
    "room": {"type": "rectangular", "width": 5.5, "depth": 7.0, "height": 2.6},
    "sources": [
        {"name": "Front Left", "position": {"x": 1.2, "y": 0.3, "z": 1.1}, "amplitude": 0.8,
         "directivity": {"type": "omnidirectional"}, "crossover": {"type": "highpass", "cutoff_freq": 80.0, "order": 4}},
        {"name": "Front Right", "position": {"x": 4.3, "y": 0.3, "z": 1.1}, "amplitude": 0.8,
         "directivity": {"type": "omnidirectional"}, "crossover": {"type": "highpass", "cutoff_freq": 80.0, "order": 4}},
        {"name": "Subwoofer", "position": {"x": 0.5, "y": 0.5, "z": 0.3}, "amplitude": 1.0,
         "directivity": {"type": "omnidirectional"}, "crossover": {"type": "lowpass", "cutoff_freq": 80.0, "order": 4}},
    ],
    "listening_positions": [{"x": 2.75, "y": 4.5, "z": 1.2}],
    "frequencies": {"min_freq": 20.0, "max_freq": 300.0, "num_points": 100, "spacing": "logarithmic"},
    "solver": {"method": "gmres", "mesh_resolution": 10},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frequencies", type=int, default=12, help="how many of the 100 grid frequencies to run (evenly spread)")
    ap.add_argument("--mesh-resolution", type=int, default=None)
    ap.add_argument("--rows-checked", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    import torch

    from math_audio_b200 import bem, room
    from oracle import oracle as orc
    from oracle import room_oracle as ro

    sim, res = room.simulation_from_config(CONFIG)
    if args.mesh_resolution:
        res = args.mesh_resolution
    idx = np.unique(np.linspace(0, len(sim.frequencies) - 1, args.frequencies).round().astype(int))
    freqs = [sim.frequencies[i] for i in idx]
    mesh = sim.room.generate_mesh(res)
    n = mesh.num_elements()
    ctx = bem.default_context()
    st = room.StagedRoomMesh(mesh, ctx)
    lp = np.asarray(sim.listening_positions[0], dtype=np.float64).reshape(1, 3)
    hbm = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6451.8) if (ROOT / "MEASURED_PEAKS.json").exists() else 6451.8

    # warm-up (first-call costs), then the timed sweep
    room.solve_bem_system(st, sim.sources, sim.wavenumber(freqs[0]), freqs[0])
    torch.cuda.synchronize()
    per = []
    matrix = None
    t_all = time.perf_counter()
    for f in freqs:
        k = sim.wavenumber(f)
        t0 = time.perf_counter()
        info = {}
        x = room.solve_bem_system(st, sim.sources, k, f, reuse=matrix, info=info)
        matrix = info.pop("matrix")
        p = room.calculate_field_pressure_bem_parallel(st, x, sim.sources, lp, k, f)
        dt = time.perf_counter() - t0
        stt = matrix.solver_stats()
        per.append(dict(frequency=f, seconds=dt, spl=room.pressure_to_spl(p[0]), iterations=info["iterations"], converged=info["converged"],
                        residual=info["residual"], assembly_kernel_ms=info["assembly_kernel_ms"],
                        matvec_ms=stt["matvec_ms"], matvecs=stt["matvecs"]))
    wall = time.perf_counter() - t_all
    asm_ms = float(np.mean([p["assembly_kernel_ms"] for p in per]))
    mv_ms = sum(p["matvec_ms"] for p in per) / max(1, sum(p["matvecs"] for p in per))
    out = dict(config="home_theater_2_1-like (restated)", n_elements=n, mesh_resolution=res, n_frequencies=len(freqs),
               seconds_per_frequency=wall / len(freqs), per_frequency=per,
               matrix_kernel=dict(avg_ms=asm_ms, bytes=16 * n * n, gbs=16 * n * n / (asm_ms * 1e-3) / 1e9, bound="hbm (16 B written per entry)",
                                  frac_of_measured_hbm=16 * n * n / (asm_ms * 1e-3) / 1e9 / hbm),
               zgemv=dict(avg_ms=mv_ms, gbs=(16 * n * n + 32 * n) / (mv_ms * 1e-3) / 1e9))

    # ---- parity against the oracle: sampled rows, right-hand side, SPL at the last frequency ----
    f = freqs[-1]
    k = sim.wavenumber(f)
    c, nm, a = ro.element_data(mesh.nodes, mesh.elements)
    gc, gn, ga = st.geometry()
    osrcs = [dict(position=[float(v) for v in s.position], amplitude=s.amplitude, directivity=None,
                  crossover=dict(kind=s.crossover.kind, cutoff=s.crossover.cutoff_freq, order=s.crossover.order)) for s in sim.sources]
    rng = np.random.default_rng(11)
    rows = np.sort(rng.choice(n, size=min(args.rows_checked, n), replace=False))
    A = room.build_bem_matrix_parallel(st, k, reuse=matrix).rows()
    worst = 0.0
    for r in rows:
        ref = ro.build_bem_matrix(c, nm, a, k, rows=(int(r), int(r) + 1))[0]
        worst = max(worst, float(np.max(np.abs(A[r] - ref)) / np.max(np.abs(ref))))
    rhs = room.calculate_incident_field_derivative_parallel(st, sim.sources, k, f)
    rref = ro.incident_field_derivative(c[rows], nm[rows], osrcs, k, f)
    x = room.solve_bem_system(st, sim.sources, k, f, reuse=matrix)
    p_gpu = room.calculate_field_pressure_bem_parallel(st, x, sim.sources, lp, k, f)
    p_ref = ro.field_pressure(c, nm, a, x, osrcs, lp, k, f)
    resid = float(np.linalg.norm(A @ x - rhs) / np.linalg.norm(rhs))
    out["parity"] = dict(geometry_bit_exact=bool(np.array_equal(gc, c) and np.array_equal(gn, nm) and np.array_equal(ga, a)),
                         rows_checked=int(len(rows)), max_rownorm_entry_err=worst,
                         max_rhs_rel_err=float(np.max(np.abs(rhs[rows] - rref)) / np.max(np.abs(rref))),
                         field_rel_err=float(abs(p_gpu[0] - p_ref[0]) / abs(p_ref[0])),
                         independent_residual=resid)

    # ---- bounded CPU baseline -------------------------------------------------------------------
    if not args.no_cpu_baseline:
        slab = min(n, 512)
        t0 = time.perf_counter()
        As = ro.build_bem_matrix(c, nm, a, k, rows=(0, slab))
        t_mat = (time.perf_counter() - t0) * n / slab
        xs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        orc.build()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            orc.zgemv(As, xs)
        t_mv = (time.perf_counter() - t0) / reps * n / slab
        mean_mv = float(np.mean([p["matvecs"] for p in per]))
        out["cpu_baseline"] = dict(kind="port", cores=orc.num_threads(), seconds_per_frequency=t_mat + mean_mv * t_mv,
                                   matrix_s=t_mat, matvec_s=t_mv, matvecs=mean_mv,
                                   sample=f"numpy oracle matrix on {slab} of {n} rows + C++ oracle zgemv on that slab, extrapolated")
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "room_home_theater.json").write_text(json.dumps(out, indent=1))
    brief = {k2: v for k2, v in out.items() if k2 != "per_frequency"}
    brief["spl_first_last"] = [per[0]["spl"], per[-1]["spl"]]
    brief["iterations"] = [p["iterations"] for p in per]
    print(json.dumps(brief))


if __name__ == "__main__":
    main()
