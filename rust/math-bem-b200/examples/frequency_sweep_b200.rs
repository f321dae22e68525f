//! Caller-side switch to the B200 backend: the loop of math-bem/examples/audio_frequency_sweep.rs with
//! `build_tbem_system_with_beta` + `solve_gmres` replaced by their GPU drop-ins.  The switch lives HERE, in the caller --
//! math-bem itself is unchanged and does not depend on the GPU crates.
//!
//!     cargo run --release -p math-bem-b200 --example frequency_sweep_b200 [--cpu]
use math_audio_bem::core::assembly::tbem::build_tbem_system_with_beta;
use math_audio_bem::core::incident::IncidentField;
use math_audio_bem::core::mesh::generators::generate_icosphere_mesh;
use math_audio_bem::core::solver::fmm_interface::{solve_gmres, DenseOperator};
use math_audio_bem::core::types::PhysicsParams;
use math_audio_solvers::iterative::GmresConfig;
use math_bem_b200::{build_tbem_system_gpu, GpuContext, GpuSweep};
use ndarray::{Array1, Array2};
use num_complex::Complex64;

fn main() -> Result<(), String> {
    let use_cpu = std::env::args().any(|a| a == "--cpu");
    let radius = 0.1;
    let mesh = generate_icosphere_mesh(radius, 5);                       // 20 480 Tri3
    let config = GmresConfig { max_iterations: 1000, restart: 50, tolerance: 1e-10, print_interval: 0 };
    let frequencies: Vec<f64> = (0..64).map(|i| 136.5 * (32.0f64).powf(i as f64 / 63.0)).collect();
    // centers / normals as (n, 3) arrays, as in audio_frequency_sweep.rs:102-110
    let n = mesh.elements.len();
    let mut centers = Array2::<f64>::zeros((n, 3));
    let mut normals = Array2::<f64>::zeros((n, 3));
    for (i, e) in mesh.elements.iter().enumerate() {
        for j in 0..3 { centers[[i, j]] = e.center[j]; normals[[i, j]] = e.normal[j]; }
    }
    let case = |f: f64| {
        let physics = PhysicsParams::new(f, 343.0, 1.21, false);
        let (beta, _) = physics.burton_miller_beta_adaptive(radius);
        let rhs = IncidentField::plane_wave_z().compute_rhs_with_beta(&centers, &normals, &physics, beta);
        (physics, beta, rhs)
    };
    let mut solutions: Vec<Array1<Complex64>> = Vec::new();
    if use_cpu {
        // the reference path, unchanged
        for &f in &frequencies {
            let (physics, beta, rhs) = case(f);
            let system = build_tbem_system_with_beta(&mesh.elements, &mesh.nodes, &physics, beta);
            let b = &system.rhs + &rhs;
            solutions.push(solve_gmres(&DenseOperator::new(system.matrix), &b, &config).x);
        }
    } else if std::env::args().any(|a| a == "--sequential") {
        // drop-in, call by call: build_tbem_system_gpu + GpuDenseOperator::gmres
        let ctx = GpuContext::new(0)?;
        for &f in &frequencies {
            let (physics, beta, rhs) = case(f);
            let system = build_tbem_system_gpu(&ctx, &mesh.elements, &mesh.nodes, &physics, beta)?;
            let b = &system.rhs + &rhs;
            solutions.push(system.operator.gmres(&b, &config).x);
        }
    } else {
        // pipelined: assembly of frequency f + 1 underneath the solve of f, two frequencies in flight
        let mut sweep = GpuSweep::new(0, &mesh.elements, &mesh.nodes)?;
        for &f in frequencies.iter().take(2) { let (p, b, r) = case(f); sweep.submit(&p, b, &r, &config)?; }
        for i in 0..frequencies.len() {
            solutions.push(sweep.next()?.x);
            if i + 2 < frequencies.len() { let (p, b, r) = case(frequencies[i + 2]); sweep.submit(&p, b, &r, &config)?; }
        }
    }
    println!("{} frequencies solved, |p| at element 0 of the last one: {:.6e}", solutions.len(), solutions.last().unwrap()[0].norm());
    Ok(())
}
