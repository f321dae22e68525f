//! FFI declarations of `include/bemb200.h` and safe wrappers with the reference's signatures:
//! `build_tbem_system_gpu` (<- `build_tbem_system_with_beta`, math-bem/src/core/assembly/tbem.rs:96)
//! and `GpuDenseOperator: LinearOperator<Complex64>` (<- `DenseOperator`,
//! math-bem/src/core/solver/fmm_interface.rs:25-52).  Written without a Rust toolchain at hand
//!; the executable twin of this file is math_audio_b200/bem.py.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

use math_audio_bem::core::types::{BoundaryCondition, Element, ElementType, PhysicsParams};
use math_audio_solvers::iterative::{BiCgstabConfig, BiCgstabSolution, CgsConfig, CgsSolution, GmresConfig, GmresSolution};
use math_audio_solvers::traits::LinearOperator;
use ndarray::{Array1, Array2};
use num_complex::Complex64;

#[repr(C)] pub struct bemb200_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bemb200_staged_mesh { _p: [u8; 0] }
#[repr(C)] pub struct bemb200_matrix { _p: [u8; 0] }

#[repr(C)]
pub struct bemb200_mesh {
    pub n_nodes: u64, pub n_elem: u64,
    pub nodes: *const f64, pub conn: *const u32, pub etype: *const u8,
    pub center: *const f64, pub normal: *const f64, pub area: *const f64,
    pub bc_type: *const i32, pub bc_len: *const u8, pub bc_val: *const f64,
    pub dof: *const u32, pub is_eval: *const u8,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct bemb200_physics { pub wave_number: f64, pub harmonic_factor: f64, pub tau: f64, pub gamma: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct bemb200_gmres_info { pub iterations: u64, pub restarts: u64, pub residual: f64, pub converged: i32 }

extern "C" {
    pub fn bemb200_ctx_create(device: c_int, out: *mut *mut bemb200_ctx) -> c_int;
    pub fn bemb200_ctx_create_ex(device: c_int, rank: c_int, nranks: c_int, nccl_id: *const u8,
                                 cuda_stream: *mut c_void, out: *mut *mut bemb200_ctx) -> c_int;
    pub fn bemb200_nccl_unique_id(out: *mut u8) -> c_int;
    pub fn bemb200_ctx_destroy(ctx: *mut bemb200_ctx);
    pub fn bemb200_ctx_set_background(ctx: *mut bemb200_ctx, blocks_per_sm: c_int) -> c_int;
    pub fn bemb200_ctx_set_shared_gpu(ctx: *mut bemb200_ctx, shared: c_int) -> c_int;
    pub fn bemb200_ctx_peer_exchange_active(ctx: *const bemb200_ctx, active: *mut c_int) -> c_int;
    pub fn bemb200_last_error(ctx: *const bemb200_ctx) -> *const c_char;
    pub fn bemb200_partition(n: u64, nranks: c_int, rank: c_int, row_begin: *mut u64, row_end: *mut u64);
    pub fn bemb200_mesh_stage(ctx: *mut bemb200_ctx, mesh: *const bemb200_mesh, out: *mut *mut bemb200_staged_mesh) -> c_int;
    pub fn bemb200_staged_mesh_free(sm: *mut bemb200_staged_mesh);
    pub fn bemb200_assemble_staged(ctx: *mut bemb200_ctx, sm: *const bemb200_staged_mesh, phys: *const bemb200_physics,
                                   beta_re: c_double, beta_im: c_double, row_begin: u64, row_end: u64,
                                   inout: *mut *mut bemb200_matrix) -> c_int;
    pub fn bemb200_assemble(ctx: *mut bemb200_ctx, mesh: *const bemb200_mesh, phys: *const bemb200_physics,
                            beta_re: c_double, beta_im: c_double, row_begin: u64, row_end: u64,
                            out: *mut *mut bemb200_matrix) -> c_int;
    pub fn bemb200_matrix_from_host(ctx: *mut bemb200_ctx, a_rows: *const f64, n_rows_global: u64, n_cols: u64,
                                    row_begin: u64, row_end: u64, out: *mut *mut bemb200_matrix) -> c_int;
    pub fn bemb200_matrix_free(m: *mut bemb200_matrix);
    pub fn bemb200_num_rows(m: *const bemb200_matrix) -> u64;
    pub fn bemb200_num_cols(m: *const bemb200_matrix) -> u64;
    pub fn bemb200_matrix_download(m: *const bemb200_matrix, row_begin: u64, row_end: u64, out: *mut f64) -> c_int;
    pub fn bemb200_rhs_download_full(m: *const bemb200_matrix, out: *mut f64) -> c_int;
    pub fn bemb200_row_sum_correction(m: *mut bemb200_matrix, avg: *mut f64) -> c_int;
    pub fn bemb200_apply(m: *const bemb200_matrix, x: *const f64, y: *mut f64) -> c_int;
    pub fn bemb200_apply_transpose(m: *const bemb200_matrix, x: *const f64, y: *mut f64) -> c_int;
    pub fn bemb200_bicgstab(m: *const bemb200_matrix, b: *const f64, max_iterations: u32, tolerance: f64, x_out: *mut f64,
                            info: *mut bemb200_gmres_info) -> c_int;
    pub fn bemb200_cgs(m: *const bemb200_matrix, b: *const f64, max_iterations: u32, tolerance: f64, x_out: *mut f64,
                       info: *mut bemb200_gmres_info) -> c_int;
    pub fn bemb200_lu_solve(m: *const bemb200_matrix, b: *const f64, x_out: *mut f64, overwrite_matrix: c_int,
                            factor_ms: *mut f64) -> c_int;
    pub fn bemb200_compute_rcs(sm: *const bemb200_staged_mesh, phys: *const bemb200_physics, n_dirs: u32, dirs: *const f64,
                               surface_pressure: *const f64, rcs_out: *mut f64) -> c_int;
    pub fn bemb200_gmres_preconditioned(m: *const bemb200_matrix, inv_diag: *const f64, b: *const f64, x0: *const f64,
                                        max_iterations: u32, restart: u32, tolerance: f64, x_out: *mut f64,
                                        info: *mut bemb200_gmres_info) -> c_int;
    pub fn bemb200_matrix_diagonal(m: *const bemb200_matrix, out: *mut f64) -> c_int;
    pub fn bemb200_gmres_batched(m: *const bemb200_matrix, b_all: *const f64, nrhs: u32, max_iterations: u32, restart: u32,
                                 tolerance: f64, x_all: *mut f64, infos: *mut bemb200_gmres_info, block_matvec_ms: *mut f64,
                                 block_matvecs: *mut u64) -> c_int;
    pub fn bemb200_incident_rhs(sm: *const bemb200_staged_mesh, phys: *const bemb200_physics, beta_re: f64, beta_im: f64,
                                n_sources: u32, kinds: *const i32, vecs: *const f64, amps: *const f64, rhs_host: *mut f64,
                                rhs_dev: *mut f64) -> c_int;
    pub fn bemb200_scattered_field(sm: *const bemb200_staged_mesh, phys: *const bemb200_physics, n_eval: u64, eval_pts: *const f64,
                                   surface_pressure: *const f64, surface_velocity: *const f64, out: *mut f64) -> c_int;
    pub fn bemb200_gmres(m: *const bemb200_matrix, b: *const f64, x0: *const f64, max_iterations: u32, restart: u32,
                         tolerance: f64, x_out: *mut f64, info: *mut bemb200_gmres_info) -> c_int;
}

/// Owns the device context (one per GPU / per rank).
pub struct GpuContext(*mut bemb200_ctx);
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bemb200_ctx_create(device, &mut h) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(Self(h))
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { bemb200_ctx_destroy(self.0) } } }

fn last_error(ctx: *const bemb200_ctx) -> String {
    unsafe { std::ffi::CStr::from_ptr(bemb200_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Device-resident replacement of `TbemSystem` (tbem.rs:13-20): the matrix stays on the GPU.
pub struct GpuTbemSystem { pub operator: GpuDenseOperator, pub rhs: Array1<Complex64>, pub num_dofs: usize }

/// Drop-in for `build_tbem_system_with_beta(elements, nodes, physics, beta)` (tbem.rs:96-101).
pub fn build_tbem_system_gpu(ctx: &GpuContext, elements: &[Element], nodes: &Array2<f64>, physics: &PhysicsParams,
                             beta: Complex64) -> Result<GpuTbemSystem, String> {
    // AoS -> SoA (the only host work; O(N))
    let n = elements.len();
    let nodes_c = nodes.as_standard_layout();
    let (mut conn, mut etype) = (vec![u32::MAX; 4 * n], vec![0u8; n]);
    let (mut center, mut normal, mut area) = (vec![0f64; 3 * n], vec![0f64; 3 * n], vec![0f64; n]);
    let (mut bc_type, mut bc_len, mut bc_val) = (vec![0i32; n], vec![1u8; n], vec![0f64; 8 * n]);
    let (mut dof, mut is_eval) = (vec![0u32; n], vec![0u8; n]);
    for (i, e) in elements.iter().enumerate() {
        etype[i] = match e.element_type { ElementType::Tri3 => 3, ElementType::Quad4 => 4 };
        for (v, &c) in e.connectivity.iter().enumerate() { conn[4 * i + v] = c as u32; }
        for d in 0..3 { center[3 * i + d] = e.center[d]; normal[3 * i + d] = e.normal[d]; }
        area[i] = e.area;
        // get_bc_type_and_value(): tbem.rs:234-244
        let (t, vals): (i32, Vec<Complex64>) = match &e.boundary_condition {
            BoundaryCondition::Velocity(v) => (0, v.clone()),
            BoundaryCondition::Pressure(p) => (1, p.clone()),
            BoundaryCondition::VelocityWithAdmittance { velocity, .. } => (0, velocity.clone()),
            _ => (2, vec![Complex64::new(0.0, 0.0)]),
        };
        bc_type[i] = t;
        bc_len[i] = vals.len().min(4) as u8;
        for (k, z) in vals.iter().take(4).enumerate() { bc_val[8 * i + 2 * k] = z.re; bc_val[8 * i + 2 * k + 1] = z.im; }
        dof[i] = e.dof_addresses[0] as u32;
        is_eval[i] = e.property.is_evaluation() as u8;
    }
    let mesh = bemb200_mesh {
        n_nodes: nodes.nrows() as u64, n_elem: n as u64, nodes: nodes_c.as_ptr(), conn: conn.as_ptr(), etype: etype.as_ptr(),
        center: center.as_ptr(), normal: normal.as_ptr(), area: area.as_ptr(), bc_type: bc_type.as_ptr(),
        bc_len: bc_len.as_ptr(), bc_val: bc_val.as_ptr(), dof: dof.as_ptr(), is_eval: is_eval.as_ptr(),
    };
    let phys = bemb200_physics { wave_number: physics.wave_number, harmonic_factor: physics.harmonic_factor,
                                 tau: physics.tau, gamma: physics.gamma() };
    let ndof = elements.iter().filter(|e| !e.property.is_evaluation()).count();
    let mut m = std::ptr::null_mut();
    let rc = unsafe { bemb200_assemble(ctx.0, &mesh, &phys, beta.re, beta.im, 0, ndof as u64, &mut m) };
    if rc != 0 { return Err(last_error(ctx.0)); }
    let mut rhs = Array1::<Complex64>::zeros(ndof);
    let rc = unsafe { bemb200_rhs_download_full(m, rhs.as_mut_ptr() as *mut f64) };
    if rc != 0 { unsafe { bemb200_matrix_free(m) }; return Err(last_error(ctx.0)); }
    Ok(GpuTbemSystem { operator: GpuDenseOperator(m), rhs, num_dofs: ndof })
}

/// `DenseOperator` on the device.  `Complex64` is `#[repr(C)] {re, im}` = two doubles.
pub struct GpuDenseOperator(*mut bemb200_matrix);
unsafe impl Send for GpuDenseOperator {}   // the library serialises submissions per context
unsafe impl Sync for GpuDenseOperator {}
impl Drop for GpuDenseOperator { fn drop(&mut self) { unsafe { bemb200_matrix_free(self.0) } } }

impl GpuDenseOperator {
    /// `DenseOperator::new(matrix)`: upload an existing host matrix.
    pub fn from_array(ctx: &GpuContext, a: &Array2<Complex64>) -> Result<Self, String> {
        let a = a.as_standard_layout();
        let mut m = std::ptr::null_mut();
        let rc = unsafe { bemb200_matrix_from_host(ctx.0, a.as_ptr() as *const f64, a.nrows() as u64, a.ncols() as u64,
                                                   0, a.nrows() as u64, &mut m) };
        if rc != 0 { return Err(last_error(ctx.0)); }
        Ok(Self(m))
    }
    /// `gmres(operator, b, config)` (gmres.rs:96) executed on the device.
    pub fn gmres(&self, b: &Array1<Complex64>, config: &GmresConfig<f64>) -> GmresSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");   // the reference panics too
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_gmres(self.0, b.as_ptr() as *const f64, std::ptr::null(), config.max_iterations as u32,
                                        config.restart as u32, config.tolerance, x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize,
                        residual: info.residual, converged: info.converged != 0 }
    }
    /// `gmres_preconditioned(operator, &DiagonalPreconditioner::from_diagonal(&diag), b, config)` (gmres.rs:282,
    /// preconditioners/diagonal.rs:40-50) with the Jacobi preconditioner of the matrix itself; `jacobi = false`
    /// is `IdentityPreconditioner` (traits.rs:377-385).
    pub fn gmres_preconditioned(&self, jacobi: bool, b: &Array1<Complex64>, config: &GmresConfig<f64>) -> GmresSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let n = b.len();
        let mut inv = Array1::<Complex64>::from_elem(n, Complex64::new(1.0, 0.0));
        if jacobi {
            let mut d = Array1::<Complex64>::zeros(n);
            let rc = unsafe { bemb200_matrix_diagonal(self.0, d.as_mut_ptr() as *mut f64) };
            assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
            for i in 0..n { if d[i].norm() > 1e-30 { inv[i] = d[i].inv(); } }   // diagonal.rs:29-35
        }
        let mut x = Array1::<Complex64>::zeros(n);
        let mut info = bemb200_gmres_info::default();
        let pinv = if jacobi { inv.as_ptr() as *const f64 } else { std::ptr::null() };
        let rc = unsafe { bemb200_gmres_preconditioned(self.0, pinv, b.as_ptr() as *const f64, std::ptr::null(),
                                                       config.max_iterations as u32, config.restart as u32, config.tolerance,
                                                       x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize,
                        residual: info.residual, converged: info.converged != 0 }
    }
    /// Several right-hand sides at once (what the reference does with a loop of `gmres` calls): the solves advance in
    /// lockstep on one tensor-core block matvec; each result has the semantics of its own `gmres` call.  At most 32.
    pub fn gmres_batched(&self, bs: &[Array1<Complex64>], config: &GmresConfig<f64>) -> Vec<GmresSolution<Complex64>> {
        let n = self.num_rows();
        let nrhs = bs.len();
        let mut b_all = Vec::<Complex64>::with_capacity(nrhs * n);
        for b in bs { assert_eq!(b.len(), n, "Vector lengths must match"); b_all.extend(b.iter().cloned()); }
        let mut x_all = vec![Complex64::new(0.0, 0.0); nrhs * n];
        let mut infos = vec![bemb200_gmres_info::default(); nrhs];
        let rc = unsafe { bemb200_gmres_batched(self.0, b_all.as_ptr() as *const f64, nrhs as u32, config.max_iterations as u32,
                                                config.restart as u32, config.tolerance, x_all.as_mut_ptr() as *mut f64,
                                                infos.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        (0..nrhs).map(|s| GmresSolution {
            x: Array1::from(x_all[s * n..(s + 1) * n].to_vec()), iterations: infos[s].iterations as usize,
            restarts: infos[s].restarts as usize, residual: infos[s].residual, converged: infos[s].converged != 0,
        }).collect()
    }
    /// `bicgstab(operator, b, config)` (bicgstab.rs:53), the solver of `BemSolver::solve_dense_system`.
    pub fn bicgstab(&self, b: &Array1<Complex64>, config: &BiCgstabConfig<f64>) -> BiCgstabSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_bicgstab(self.0, b.as_ptr() as *const f64, config.max_iterations as u32, config.tolerance,
                                           x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        BiCgstabSolution { x, iterations: info.iterations as usize, residual: info.residual, converged: info.converged != 0 }
    }
    /// `cgs(operator, b, config)` (cgs.rs:46); what `solve_cgs` / `solve_with_ilu` / `solve_tbem_with_ilu`
    /// (fmm_interface.rs:360-366,389-447) run on a dense matrix.
    pub fn cgs(&self, b: &Array1<Complex64>, config: &CgsConfig<f64>) -> CgsSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_cgs(self.0, b.as_ptr() as *const f64, config.max_iterations as u32, config.tolerance,
                                      x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        CgsSolution { x, iterations: info.iterations as usize, residual: info.residual, converged: info.converged != 0 }
    }
    /// `lu_solve(&a, &b)` (direct/lu.rs:136): cuSOLVER zgetrf + zgetrs on a copy; `Err` = `LuError::SingularMatrix`.
    pub fn lu_solve(&self, b: &Array1<Complex64>) -> Result<Array1<Complex64>, String> {
        if b.len() != self.num_rows() { return Err("Matrix dimensions mismatch".into()); }
        let mut x = Array1::<Complex64>::zeros(b.len());
        let rc = unsafe { bemb200_lu_solve(self.0, b.as_ptr() as *const f64, x.as_mut_ptr() as *mut f64, 0, std::ptr::null_mut()) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(x)
    }
}

impl LinearOperator<Complex64> for GpuDenseOperator {
    fn num_rows(&self) -> usize { unsafe { bemb200_num_rows(self.0) as usize } }
    fn num_cols(&self) -> usize { unsafe { bemb200_num_cols(self.0) as usize } }
    fn apply(&self, x: &Array1<Complex64>) -> Array1<Complex64> {
        assert_eq!(x.len(), self.num_cols());
        let mut y = Array1::<Complex64>::zeros(self.num_rows());
        let rc = unsafe { bemb200_apply(self.0, x.as_ptr() as *const f64, y.as_mut_ptr() as *mut f64) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        y
    }
    fn apply_transpose(&self, x: &Array1<Complex64>) -> Array1<Complex64> {
        assert_eq!(x.len(), self.num_rows());
        let mut y = Array1::<Complex64>::zeros(self.num_cols());
        let rc = unsafe { bemb200_apply_transpose(self.0, x.as_ptr() as *const f64, y.as_mut_ptr() as *mut f64) };
        assert_eq!(rc, 0, "libbemb200: {}", last_error(std::ptr::null()));
        y
    }
}
