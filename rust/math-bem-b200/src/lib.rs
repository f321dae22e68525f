//! Safe wrappers over `bem-b200-sys` with the reference's signatures: `build_tbem_system_gpu`
//! (<- `build_tbem_system_with_beta`, math-bem/src/core/assembly/tbem.rs:96), `GpuDenseOperator: LinearOperator<Complex64>`
//! (<- `DenseOperator`, math-bem/src/core/solver/fmm_interface.rs:25-52), `GpuStagedMesh` (staged assembly, field evaluation, RCS), `GpuSweep` (the per-frequency loop of
//! math-bem/examples/audio_frequency_sweep.rs with the assembly of f + 1 underneath the solve of f) and `GpuGroup`
//! (one process, several devices).  This crate depends on math-bem and math-solvers; neither depends on it: the switch
//! to the GPU is made by the CALLER (see examples/frequency_sweep_b200.rs), so there is no dependency cycle.
//! Written without a Rust toolchain at hand; the executable twin is math_audio_b200/bem.py.
#![allow(non_camel_case_types)]
use std::sync::Arc;

use bem_b200_sys::*;
use math_audio_bem::core::types::{BoundaryCondition, Element, ElementType, PhysicsParams};
use math_audio_solvers::iterative::{BiCgstabConfig, BiCgstabSolution, CgsConfig, CgsSolution, GmresConfig, GmresSolution};
use math_audio_solvers::traits::{LinearOperator, Preconditioner};
use ndarray::{Array1, Array2};
use num_complex::Complex64;

/// Owns the device context (one per GPU / per rank).  Cheap to clone: every matrix / operator created through a
/// context keeps its own `Arc` to it, so the context outlives them whatever the drop order in the caller.
struct CtxInner(*mut bemb200_ctx);
unsafe impl Send for CtxInner {}
unsafe impl Sync for CtxInner {}
impl Drop for CtxInner { fn drop(&mut self) { unsafe { bemb200_ctx_destroy(self.0) } } }
#[derive(Clone)]
pub struct GpuContext(Arc<CtxInner>);
impl GpuContext {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bemb200_ctx_create(device, &mut h) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(Self(Arc::new(CtxInner(h))))
    }
    fn raw(&self) -> *mut bemb200_ctx { (self.0).0 }
    fn error(&self) -> String { last_error(self.raw()) }
}

fn last_error(ctx: *const bemb200_ctx) -> String {
    unsafe { std::ffi::CStr::from_ptr(bemb200_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Device-resident replacement of `TbemSystem` (tbem.rs:13-20): the matrix stays on the GPU.
pub struct GpuTbemSystem { pub operator: GpuDenseOperator, pub rhs: Array1<Complex64>, pub num_dofs: usize }

/// Drop-in for `build_tbem_system_with_beta(elements, nodes, physics, beta)` (tbem.rs:96-101).
pub fn build_tbem_system_gpu(ctx: &GpuContext, elements: &[Element], nodes: &Array2<f64>, physics: &PhysicsParams,
                             beta: Complex64) -> Result<GpuTbemSystem, String> {
    // AoS -> SoA (the only host work; O(N))
    let flat = flatten(elements, nodes)?;
    let (mesh, phys) = (flat.view(), physics_of(physics));
    let ndof = elements.iter().filter(|e| !e.property.is_evaluation()).count();
    let mut m = std::ptr::null_mut();
    let rc = unsafe { bemb200_assemble(ctx.raw(), &mesh, &phys, beta.re, beta.im, 0, ndof as u64, &mut m) };
    if rc != 0 { return Err(last_error(ctx.raw())); }
    let mut rhs = Array1::<Complex64>::zeros(ndof);
    let rc = unsafe { bemb200_rhs_download_full(m, rhs.as_mut_ptr() as *mut f64) };
    if rc != 0 { unsafe { bemb200_matrix_free(m) }; return Err(last_error(ctx.raw())); }
    Ok(GpuTbemSystem { operator: GpuDenseOperator { m, ctx: ctx.clone() }, rhs, num_dofs: ndof })
}

/// `DenseOperator` on the device.  `Complex64` is `#[repr(C)] {re, im}` = two doubles.
pub struct GpuDenseOperator { m: *mut bemb200_matrix, ctx: GpuContext }   // `ctx` keeps the context alive: bemb200_matrix_free reads it
unsafe impl Send for GpuDenseOperator {}   // the library serialises submissions per context
unsafe impl Sync for GpuDenseOperator {}
impl Drop for GpuDenseOperator { fn drop(&mut self) { unsafe { bemb200_matrix_free(self.m) } } }

impl GpuDenseOperator {
    /// `DenseOperator::new(matrix)`: upload an existing host matrix.
    pub fn from_array(ctx: &GpuContext, a: &Array2<Complex64>) -> Result<Self, String> {
        let a = a.as_standard_layout();
        let mut m = std::ptr::null_mut();
        let rc = unsafe { bemb200_matrix_from_host(ctx.raw(), a.as_ptr() as *const f64, a.nrows() as u64, a.ncols() as u64,
                                                   0, a.nrows() as u64, &mut m) };
        if rc != 0 { return Err(last_error(ctx.raw())); }
        Ok(Self { m, ctx: ctx.clone() })
    }
    /// `gmres(operator, b, config)` (gmres.rs:96) executed on the device.
    pub fn gmres(&self, b: &Array1<Complex64>, config: &GmresConfig<f64>) -> GmresSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");   // the reference panics too
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_gmres(self.m, contiguous(b).as_ptr() as *const f64, std::ptr::null(), config.max_iterations as u32,
                                        config.restart as u32, config.tolerance, x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
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
            let rc = unsafe { bemb200_matrix_diagonal(self.m, d.as_mut_ptr() as *mut f64) };
            assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
            for i in 0..n { if d[i].norm() > 1e-30 { inv[i] = d[i].inv(); } }   // diagonal.rs:29-35
        }
        let mut x = Array1::<Complex64>::zeros(n);
        let mut info = bemb200_gmres_info::default();
        let pinv = if jacobi { inv.as_ptr() as *const f64 } else { std::ptr::null() };
        let rc = unsafe { bemb200_gmres_preconditioned(self.m, pinv, contiguous(b).as_ptr() as *const f64, std::ptr::null(),
                                                       config.max_iterations as u32, config.restart as u32, config.tolerance,
                                                       x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
        GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize,
                        residual: info.residual, converged: info.converged != 0 }
    }
    /// `gmres_preconditioned(operator, &AdditiveSchwarzPreconditioner::from_csr(&csr, num_subdomains, 0), b, config)`
    /// (preconditioners/schwarz.rs:66, gmres.rs:282): block-Jacobi on the contiguous diagonal blocks, built on the device from
    /// the assembled operator.  `subdomains` = explicit index sets instead (e.g. spatial clusters; overlapping sets get the
    /// reference's weights); on a row-sharded operator every set must lie inside one rank's row block.
    pub fn gmres_block_jacobi(&self, num_subdomains: usize, subdomains: Option<&[Vec<u64>]>, b: &Array1<Complex64>,
                              config: &GmresConfig<f64>) -> Result<GmresSolution<Complex64>, String> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let mut p: *mut bemb200_precond = std::ptr::null_mut();
        let rc = match subdomains {
            None => unsafe { bemb200_schwarz_create(self.m, num_subdomains as u32, std::ptr::null(), std::ptr::null(), &mut p) },
            Some(sets) => {
                let mut ptr = vec![0u64];
                let mut idx = Vec::<u64>::new();
                for s in sets { idx.extend_from_slice(s); ptr.push(idx.len() as u64); }
                unsafe { bemb200_schwarz_create(self.m, sets.len() as u32, ptr.as_ptr(), idx.as_ptr(), &mut p) }
            }
        };
        if rc != 0 { return Err(self.ctx.error()); }
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_gmres_schwarz(self.m, p, contiguous(b).as_ptr() as *const f64, std::ptr::null(),
                                                config.max_iterations as u32, config.restart as u32, config.tolerance,
                                                x.as_mut_ptr() as *mut f64, &mut info) };
        unsafe { bemb200_precond_free(p) };
        if rc != 0 { return Err(self.ctx.error()); }
        Ok(GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize,
                           residual: info.residual, converged: info.converged != 0 })
    }
    /// `gmres_preconditioned(operator, precond, b, config)` (gmres.rs:282) with ANY implementation of the reference's
    /// `Preconditioner` trait (traits.rs:366-371) -- ILU, AMG, hierarchical, the caller's own: `precond.apply` runs on the host
    /// exactly where the reference calls it while the Arnoldi process stays on the device (`bemb200_gmres_callback`).  A panic
    /// inside `apply` (or a result of the wrong length) ends the solve with `Err`.  Single-GPU operators only.
    pub fn gmres_preconditioned_with<P: Preconditioner<Complex64>>(&self, precond: &P, b: &Array1<Complex64>,
                                                                   config: &GmresConfig<f64>) -> Result<GmresSolution<Complex64>, String> {
        unsafe extern "C" fn trampoline<P: Preconditioner<Complex64>>(user: *mut std::os::raw::c_void, r: *const f64, z: *mut f64,
                                                                      n: u64) -> std::os::raw::c_int {
            let n = n as usize;
            let precond = &*(user as *const P);
            let r = Array1::from(std::slice::from_raw_parts(r as *const Complex64, n).to_vec());
            // unwinding must not cross the C frames
            match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| precond.apply(&r))) {
                Ok(out) if out.len() == n => {
                    let dst = std::slice::from_raw_parts_mut(z as *mut Complex64, n);
                    for (d, s) in dst.iter_mut().zip(out.iter()) { *d = *s; }
                    0
                }
                _ => 1,
            }
        }
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_gmres_callback(self.m, Some(trampoline::<P>), precond as *const P as *mut std::os::raw::c_void,
                                                 contiguous(b).as_ptr() as *const f64, std::ptr::null(), config.max_iterations as u32,
                                                 config.restart as u32, config.tolerance, x.as_mut_ptr() as *mut f64, &mut info,
                                                 std::ptr::null_mut()) };
        if rc != 0 { return Err(self.ctx.error()); }
        Ok(GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize,
                           residual: info.residual, converged: info.converged != 0 })
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
        let rc = unsafe { bemb200_gmres_batched(self.m, b_all.as_ptr() as *const f64, nrhs as u32, config.max_iterations as u32,
                                                config.restart as u32, config.tolerance, x_all.as_mut_ptr() as *mut f64,
                                                infos.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
        (0..nrhs).map(|s| GmresSolution {
            x: Array1::from(x_all[s * n..(s + 1) * n].to_vec()), iterations: infos[s].iterations as usize,
            restarts: infos[s].restarts as usize, residual: infos[s].residual, converged: infos[s].converged != 0,
        }).collect()
    }
    /// `bicgstab(operator, b, config)` (bicgstab.rs:46), the solver of `BemSolver::solve_dense_system`.
    pub fn bicgstab(&self, b: &Array1<Complex64>, config: &BiCgstabConfig<f64>) -> BiCgstabSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_bicgstab(self.m, contiguous(b).as_ptr() as *const f64, config.max_iterations as u32, config.tolerance,
                                           x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
        BiCgstabSolution { x, iterations: info.iterations as usize, residual: info.residual, converged: info.converged != 0 }
    }
    /// `cgs(operator, b, config)` (cgs.rs:46); what `solve_cgs` / `solve_with_ilu` / `solve_tbem_with_ilu`
    /// (fmm_interface.rs:360-366,389-447) run on a dense matrix.
    pub fn cgs(&self, b: &Array1<Complex64>, config: &CgsConfig<f64>) -> CgsSolution<Complex64> {
        assert_eq!(b.len(), self.num_rows(), "Vector lengths must match");
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_cgs(self.m, contiguous(b).as_ptr() as *const f64, config.max_iterations as u32, config.tolerance,
                                      x.as_mut_ptr() as *mut f64, &mut info) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
        CgsSolution { x, iterations: info.iterations as usize, residual: info.residual, converged: info.converged != 0 }
    }
    /// `lu_solve(&a, &b)` (direct/lu.rs:142): cuSOLVER zgetrf + zgetrs on a copy; `Err` = `LuError::SingularMatrix`.
    pub fn lu_solve(&self, b: &Array1<Complex64>) -> Result<Array1<Complex64>, String> {
        if b.len() != self.num_rows() { return Err("Matrix dimensions mismatch".into()); }
        let mut x = Array1::<Complex64>::zeros(b.len());
        let rc = unsafe { bemb200_lu_solve(self.m, contiguous(b).as_ptr() as *const f64, x.as_mut_ptr() as *mut f64, 0, std::ptr::null_mut()) };
        if rc != 0 { return Err(self.ctx.error()); }
        Ok(x)
    }
}

impl LinearOperator<Complex64> for GpuDenseOperator {
    fn num_rows(&self) -> usize { unsafe { bemb200_num_rows(self.m) as usize } }
    fn num_cols(&self) -> usize { unsafe { bemb200_num_cols(self.m) as usize } }
    fn apply(&self, x: &Array1<Complex64>) -> Array1<Complex64> {
        assert_eq!(x.len(), self.num_cols());
        let mut y = Array1::<Complex64>::zeros(self.num_rows());
        let rc = unsafe { bemb200_apply(self.m, contiguous(x).as_ptr() as *const f64, y.as_mut_ptr() as *mut f64) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
        y
    }
    fn apply_transpose(&self, x: &Array1<Complex64>) -> Array1<Complex64> {
        assert_eq!(x.len(), self.num_rows());
        let mut y = Array1::<Complex64>::zeros(self.num_cols());
        let rc = unsafe { bemb200_apply_transpose(self.m, contiguous(x).as_ptr() as *const f64, y.as_mut_ptr() as *mut f64) };
        assert_eq!(rc, 0, "libbemb200: {}", self.ctx.error());
        y
    }
}

/// `Array1` handed in by a caller may be a strided view's owner: the ABI wants unit stride.
fn contiguous(v: &Array1<Complex64>) -> std::borrow::Cow<'_, [Complex64]> {
    match v.as_slice() { Some(s) => std::borrow::Cow::Borrowed(s), None => std::borrow::Cow::Owned(v.iter().cloned().collect()) }
}

/// SoA image of `&[Element]` + `nodes` (kept alive while the C side copies it).
struct FlatMesh { nodes: Vec<f64>, conn: Vec<u32>, etype: Vec<u8>, center: Vec<f64>, normal: Vec<f64>, area: Vec<f64>,
                  bc_type: Vec<i32>, bc_len: Vec<u8>, bc_val: Vec<f64>, dof: Vec<u32>, is_eval: Vec<u8>, n_nodes: u64 }
impl FlatMesh {
    fn view(&self) -> bemb200_mesh {
        bemb200_mesh { n_nodes: self.n_nodes, n_elem: self.etype.len() as u64, nodes: self.nodes.as_ptr(), conn: self.conn.as_ptr(),
                       etype: self.etype.as_ptr(), center: self.center.as_ptr(), normal: self.normal.as_ptr(), area: self.area.as_ptr(),
                       bc_type: self.bc_type.as_ptr(), bc_len: self.bc_len.as_ptr(), bc_val: self.bc_val.as_ptr(), dof: self.dof.as_ptr(),
                       is_eval: self.is_eval.as_ptr() }
    }
}
fn flatten(elements: &[Element], nodes: &Array2<f64>) -> Result<FlatMesh, String> {
    let n = elements.len();
    let mut f = FlatMesh { nodes: nodes.as_standard_layout().iter().cloned().collect(), conn: vec![u32::MAX; 4 * n], etype: vec![0; n],
                           center: vec![0.0; 3 * n], normal: vec![0.0; 3 * n], area: vec![0.0; n], bc_type: vec![0; n], bc_len: vec![1; n],
                           bc_val: vec![0.0; 8 * n], dof: vec![0; n], is_eval: vec![0; n], n_nodes: nodes.nrows() as u64 };
    for (i, e) in elements.iter().enumerate() {
        f.etype[i] = match e.element_type { ElementType::Tri3 => 3, ElementType::Quad4 => 4 };
        for (v, &c) in e.connectivity.iter().enumerate() { f.conn[4 * i + v] = c as u32; }
        for d in 0..3 { f.center[3 * i + d] = e.center[d]; f.normal[3 * i + d] = e.normal[d]; }
        f.area[i] = e.area;
        let (t, vals): (i32, Vec<Complex64>) = match &e.boundary_condition {      // get_bc_type_and_value(): tbem.rs:234-244
            BoundaryCondition::Velocity(v) => (0, v.clone()),
            BoundaryCondition::Pressure(p) => (1, p.clone()),
            BoundaryCondition::VelocityWithAdmittance { velocity, .. } => (0, velocity.clone()),
            _ => (2, vec![Complex64::new(0.0, 0.0)]),
        };
        if vals.is_empty() || vals.len() > 4 { return Err(format!("element {i}: {} boundary-condition values (1..=4 supported)", vals.len())); }
        f.bc_type[i] = t;
        f.bc_len[i] = vals.len() as u8;
        for (k, z) in vals.iter().enumerate() { f.bc_val[8 * i + 2 * k] = z.re; f.bc_val[8 * i + 2 * k + 1] = z.im; }
        f.dof[i] = e.dof_addresses[0] as u32;
        f.is_eval[i] = e.property.is_evaluation() as u8;
    }
    Ok(f)
}
fn physics_of(p: &PhysicsParams) -> bemb200_physics {
    bemb200_physics { wave_number: p.wave_number, harmonic_factor: p.harmonic_factor, tau: p.tau, gamma: p.gamma() }
}

/// Frequency-independent device copy of a mesh (`bemb200_mesh_stage`): staged once, reused by the assembly of every frequency
/// and by the field evaluation after the solve (`compute_scattered_field` postprocess/pressure.rs:81-137, `compute_rcs` :438-478).
pub struct GpuStagedMesh { h: *mut bemb200_staged_mesh, ctx: GpuContext, num_dofs: usize, enum_to_dof: Option<Vec<usize>> }
unsafe impl Send for GpuStagedMesh {}
unsafe impl Sync for GpuStagedMesh {}
impl Drop for GpuStagedMesh { fn drop(&mut self) { unsafe { bemb200_staged_mesh_free(self.h) } } }
impl GpuStagedMesh {
    pub fn new(ctx: &GpuContext, elements: &[Element], nodes: &Array2<f64>) -> Result<Self, String> {
        let flat = flatten(elements, nodes)?;
        let mesh = flat.view();
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bemb200_mesh_stage(ctx.raw(), &mesh, &mut h) };
        if rc != 0 { return Err(ctx.error()); }
        // DOF address of the j-th non-evaluation element; `None` for the sequential map every generator produces
        let map: Vec<usize> = elements.iter().filter(|e| !e.property.is_evaluation()).map(|e| e.dof_addresses[0]).collect();
        let sequential = map.iter().enumerate().all(|(j, &d)| d == j);
        let num_dofs = map.len();
        Ok(Self { h, ctx: ctx.clone(), num_dofs, enum_to_dof: if sequential { None } else { Some(map) } })
    }
    pub fn num_dofs(&self) -> usize { self.num_dofs }
    /// `build_tbem_system_with_beta` on the staged mesh (tbem.rs:96): only the frequency-dependent work is repeated.
    pub fn build_tbem_system(&self, physics: &PhysicsParams, beta: Complex64) -> Result<GpuTbemSystem, String> {
        let phys = physics_of(physics);
        let mut m = std::ptr::null_mut();
        let rc = unsafe { bemb200_assemble_staged(self.ctx.raw(), self.h, &phys, beta.re, beta.im, 0, self.num_dofs as u64, &mut m) };
        if rc != 0 { return Err(self.ctx.error()); }
        let mut rhs = Array1::<Complex64>::zeros(self.num_dofs);
        let rc = unsafe { bemb200_rhs_download_full(m, rhs.as_mut_ptr() as *mut f64) };
        if rc != 0 { unsafe { bemb200_matrix_free(m) }; return Err(self.ctx.error()); }
        Ok(GpuTbemSystem { operator: GpuDenseOperator { m, ctx: self.ctx.clone() }, rhs, num_dofs: self.num_dofs })
    }
    /// The reference pairs entry j of a surface vector with the j-th non-evaluation element (pressure.rs:96-113); the ABI wants
    /// the entry at the element's DOF address.  Identity for sequential DOF maps.
    fn in_dof_order(&self, values: &Array1<Complex64>) -> Vec<Complex64> {
        match &self.enum_to_dof {
            None => values.iter().cloned().collect(),
            Some(map) => {
                let mut out = vec![Complex64::new(0.0, 0.0); values.len()];
                for (j, v) in values.iter().enumerate() { out[map[j]] = *v; }
                out
            }
        }
    }
    /// `compute_scattered_field(eval_points, elements, nodes, surface_pressure, surface_velocity, physics)` (pressure.rs:81).
    pub fn compute_scattered_field(&self, eval_points: &Array2<f64>, surface_pressure: &Array1<Complex64>,
                                   surface_velocity: Option<&Array1<Complex64>>, physics: &PhysicsParams) -> Result<Array1<Complex64>, String> {
        assert_eq!(eval_points.ncols(), 3, "eval_points must be n x 3");
        assert_eq!(surface_pressure.len(), self.num_dofs, "Vector lengths must match");
        let pts = eval_points.as_standard_layout();
        let ps = self.in_dof_order(surface_pressure);
        let vs = surface_velocity.map(|v| { assert_eq!(v.len(), self.num_dofs, "Vector lengths must match"); self.in_dof_order(v) });
        let phys = physics_of(physics);
        let mut out = Array1::<Complex64>::zeros(eval_points.nrows());
        let rc = unsafe { bemb200_scattered_field(self.h, &phys, eval_points.nrows() as u64, pts.as_ptr(), ps.as_ptr() as *const f64,
                                                  vs.as_ref().map_or(std::ptr::null(), |v| v.as_ptr() as *const f64),
                                                  out.as_mut_ptr() as *mut f64) };
        if rc != 0 { return Err(self.ctx.error()); }
        Ok(out)
    }
    /// `compute_rcs(surface_pressure, elements, direction, physics)` (pressure.rs:438).
    pub fn compute_rcs(&self, surface_pressure: &Array1<Complex64>, direction: [f64; 3], physics: &PhysicsParams) -> Result<f64, String> {
        assert_eq!(surface_pressure.len(), self.num_dofs, "Vector lengths must match");
        let ps = self.in_dof_order(surface_pressure);
        let phys = physics_of(physics);
        let mut out = 0.0f64;
        let rc = unsafe { bemb200_compute_rcs(self.h, &phys, 1, direction.as_ptr(), ps.as_ptr() as *const f64, &mut out) };
        if rc != 0 { return Err(self.ctx.error()); }
        Ok(out)
    }
}

/// Frequency sweep with the assembly of frequency f + 1 underneath the solve of f (`bemb200_sweep_*`): `submit` a
/// frequency, `next` returns the oldest one solved.  Keep two in flight (see examples/frequency_sweep_b200.rs).
pub struct GpuSweep { h: *mut bemb200_sweep, n: usize }
unsafe impl Send for GpuSweep {}
impl Drop for GpuSweep { fn drop(&mut self) { unsafe { bemb200_sweep_destroy(self.h) } } }
impl GpuSweep {
    pub fn new(device: i32, elements: &[Element], nodes: &Array2<f64>) -> Result<Self, String> {
        let flat = flatten(elements, nodes)?;
        let mesh = flat.view();
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bemb200_sweep_create(device, 0, 1, std::ptr::null(), &mesh, 1, 2, &mut h) };  // pipelined, 2 background blocks per SM
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        let n = unsafe { bemb200_sweep_num_dofs(h) } as usize;
        Ok(Self { h, n })
    }
    /// `rhs_extra` = `IncidentField::compute_rhs_with_beta(...)`; it is added to `TbemSystem.rhs`.
    pub fn submit(&mut self, physics: &PhysicsParams, beta: Complex64, rhs_extra: &Array1<Complex64>, config: &GmresConfig<f64>) -> Result<(), String> {
        assert_eq!(rhs_extra.len(), self.n, "Vector lengths must match");
        let phys = physics_of(physics);
        let rc = unsafe { bemb200_sweep_submit(self.h, &phys, beta.re, beta.im, contiguous(rhs_extra).as_ptr() as *const f64,
                                               config.max_iterations as u32, config.restart as u32, config.tolerance) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(())
    }
    /// Solve the frequencies returned from now on with `gmres_preconditioned` + the block-Jacobi preconditioner
    /// (`AdditiveSchwarzPreconditioner::from_csr(.., num_subdomains, 0)`, schwarz.rs:66) rebuilt from every frequency's matrix;
    /// `0` switches back to plain `gmres`.
    pub fn set_block_jacobi(&mut self, num_subdomains: usize) -> Result<(), String> {
        let rc = unsafe { bemb200_sweep_set_block_jacobi(self.h, num_subdomains as u32, std::ptr::null(), std::ptr::null()) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(())
    }
    pub fn next(&mut self) -> Result<GmresSolution<Complex64>, String> {
        let mut x = Array1::<Complex64>::zeros(self.n);
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_sweep_next(self.h, x.as_mut_ptr() as *mut f64, &mut info, std::ptr::null_mut(), std::ptr::null_mut()) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize, residual: info.residual,
                           converged: info.converged != 0 })
    }
}

/// One process, several GPUs (`bemb200_multi_*`): rows block-partitioned over `devices`, persistent fused GMRES kernel with
/// peer-memory exchange.  The shape `BemSolver::solve` (bem_solver.rs:273-322) needs: one call site, P devices.
pub struct GpuGroup(*mut bemb200_multi);
unsafe impl Send for GpuGroup {}
impl Drop for GpuGroup { fn drop(&mut self) { unsafe { bemb200_multi_destroy(self.0) } } }
pub struct GpuGroupSystem<'g> { m: *mut bemb200_multi_matrix, group: &'g GpuGroup, pub rhs: Array1<Complex64>, pub num_dofs: usize }
impl Drop for GpuGroupSystem<'_> { fn drop(&mut self) { unsafe { bemb200_multi_matrix_free(self.m) } } }
impl GpuGroup {
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bemb200_multi_create(devices.as_ptr(), devices.len() as i32, &mut h) };
        if rc != 0 { return Err(last_error(std::ptr::null())); }
        Ok(Self(h))
    }
    fn error(&self) -> String { unsafe { std::ffi::CStr::from_ptr(bemb200_multi_last_error(self.0)).to_string_lossy().into_owned() } }
    /// Drop-in for `build_tbem_system_with_beta` on the group (the borrow ties the system to the group's lifetime).
    pub fn build_tbem_system(&self, elements: &[Element], nodes: &Array2<f64>, physics: &PhysicsParams, beta: Complex64)
                             -> Result<GpuGroupSystem<'_>, String> {
        let flat = flatten(elements, nodes)?;
        let (mesh, phys) = (flat.view(), physics_of(physics));
        let mut m = std::ptr::null_mut();
        let rc = unsafe { bemb200_multi_assemble(self.0, &mesh, &phys, beta.re, beta.im, &mut m) };
        if rc != 0 { return Err(self.error()); }
        let n = unsafe { bemb200_multi_num_rows(m) } as usize;
        let mut rhs = Array1::<Complex64>::zeros(n);
        let rc = unsafe { bemb200_multi_rhs_download(m, rhs.as_mut_ptr() as *mut f64) };
        if rc != 0 { unsafe { bemb200_multi_matrix_free(m) }; return Err(self.error()); }
        Ok(GpuGroupSystem { m, group: self, rhs, num_dofs: n })
    }
}
impl GpuGroupSystem<'_> {
    /// `gmres(operator, b, config)` (gmres.rs:96) on the sharded operator.
    pub fn gmres(&self, b: &Array1<Complex64>, config: &GmresConfig<f64>) -> Result<GmresSolution<Complex64>, String> {
        assert_eq!(b.len(), self.num_dofs, "Vector lengths must match");
        let mut x = Array1::<Complex64>::zeros(b.len());
        let mut info = bemb200_gmres_info::default();
        let rc = unsafe { bemb200_multi_gmres(self.m, contiguous(b).as_ptr() as *const f64, std::ptr::null(), config.max_iterations as u32,
                                              config.restart as u32, config.tolerance, x.as_mut_ptr() as *mut f64, &mut info) };
        if rc != 0 { return Err(self.group.error()); }
        Ok(GmresSolution { x, iterations: info.iterations as usize, restarts: info.restarts as usize, residual: info.residual,
                           converged: info.converged != 0 })
    }
}
