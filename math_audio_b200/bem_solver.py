"""High-level caller of the hot path -- mirror of ``math-bem/src/core/bem_solver.rs``:

* ``SolverMethod`` :51-60, ``AssemblyMethod`` :63-72, ``BoundaryConditionType`` :75-83
* ``BemProblem`` :86-199 (``rigid_sphere_scattering`` :107-139, ``_custom`` :141-161, builders, ``ka``)
* ``BemSolver`` :202-497 (``solve`` :273-317: prepare_elements -> assemble_system (TBEM, beta =
  ``burton_miller_beta_scaled(beta_scale)``) -> add_incident_field_rhs -> solve_dense_system
  (``Direct`` = lu_solve, ``Cgs`` / ``BiCgStab`` = bicgstab on ``DenseMatrixOperator``))
* ``BemSolution`` :500-563, ``BemError`` :567-589

Every stage runs on the device through the C ABI (assembly kernels, cuSOLVER LU or the device
BiCGSTAB, field evaluation).  The FMM assembly methods are out of this repository's scope and
raise ``BemError`` ("NotImplemented"), as ``Mlfmm`` does in the reference.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import bem
from .incident import IncidentField
from .postprocess import FieldPoint, compute_total_field
from .mesh import BC_PRESSURE, BC_VELOCITY, Mesh, generate_icosphere_mesh, generate_sphere_mesh
from .types import PhysicsParams


class SolverMethod(enum.Enum):
    Direct = "direct"
    Cgs = "cgs"
    BiCgStab = "bicgstab"


class AssemblyMethod(enum.Enum):
    Tbem = "tbem"
    Slfmm = "slfmm"
    Mlfmm = "mlfmm"


class BoundaryConditionType(enum.Enum):
    Rigid = "rigid"
    Soft = "soft"
    Impedance = "impedance"


class BemError(RuntimeError):
    """bem_solver.rs:567-589 (InvalidMesh / AssemblyFailed / SolverFailed / NotImplemented)."""


@dataclass
class BemProblem:
    mesh: Mesh
    physics: PhysicsParams
    incident_field: IncidentField
    bc_type: BoundaryConditionType = BoundaryConditionType.Rigid
    use_burton_miller: bool = True

    @staticmethod
    def rigid_sphere_scattering(radius: float, frequency: float, speed_of_sound: float, density: float) -> "BemProblem":
        k = 2.0 * math.pi * frequency / speed_of_sound
        ka = k * radius
        subdivisions = 2 if ka < 1.0 else (3 if ka < 5.0 else 4)  # bem_solver.rs:117-125
        return BemProblem(generate_icosphere_mesh(radius, subdivisions), PhysicsParams.new(frequency, speed_of_sound, density, False),
                          IncidentField.plane_wave_z())

    @staticmethod
    def rigid_sphere_scattering_custom(radius: float, frequency: float, speed_of_sound: float, density: float, n_theta: int,
                                       n_phi: int) -> "BemProblem":
        return BemProblem(generate_sphere_mesh(radius, n_theta, n_phi), PhysicsParams.new(frequency, speed_of_sound, density, False),
                          IncidentField.plane_wave_z())

    def with_incident_field(self, incident: IncidentField) -> "BemProblem":
        self.incident_field = incident
        return self

    def with_boundary_condition(self, bc_type: BoundaryConditionType) -> "BemProblem":
        self.bc_type = bc_type
        return self

    def with_burton_miller(self, use_bm: bool) -> "BemProblem":
        self.use_burton_miller = use_bm
        return self

    def mesh_radius(self) -> float:
        return float(np.sqrt((self.mesh.nodes ** 2).sum(axis=1)).max())

    def ka(self) -> float:
        return self.physics.wave_number * self.mesh_radius()


@dataclass
class BemSolution:
    surface_pressure: np.ndarray
    mesh: Mesh                       # elements + nodes with the boundary conditions the solve used
    incident_field: IncidentField
    physics: PhysicsParams
    staged: Optional[bem.StagedMesh] = field(default=None, repr=False)

    def evaluate_pressure_field(self, points) -> List[FieldPoint]:
        """compute_total_field (pressure.rs:273-309): incident + scattered (device field kernel)."""
        if self.staged is None:
            self.staged = bem.StagedMesh(self.mesh)
        return compute_total_field(points, self.staged, self.surface_pressure, None, self.incident_field, self.physics)

    def evaluate_pressure(self, point) -> complex:
        return self.evaluate_pressure_field(np.asarray(point, dtype=np.float64).reshape(1, 3))[0].p_total

    def max_surface_pressure(self) -> float:
        return float(np.abs(self.surface_pressure).max())

    def mean_surface_pressure(self) -> float:
        return float(np.abs(self.surface_pressure).sum() / len(self.surface_pressure))

    def num_dofs(self) -> int:
        return int(len(self.surface_pressure))


@dataclass
class BemSolver:
    """bem_solver.rs:202-230 defaults: Direct, Tbem, 1000 iterations, 1e-8, beta_scale 4."""
    solver_method: SolverMethod = SolverMethod.Direct
    assembly_method: AssemblyMethod = AssemblyMethod.Tbem
    max_iterations: int = 1000
    tolerance: float = 1e-8
    verbose: bool = False
    beta_scale: float = 4.0

    @staticmethod
    def new() -> "BemSolver":
        return BemSolver()

    def with_solver_method(self, method: SolverMethod) -> "BemSolver":
        self.solver_method = method
        return self

    def with_assembly_method(self, method: AssemblyMethod) -> "BemSolver":
        self.assembly_method = method
        return self

    def with_max_iterations(self, max_iter: int) -> "BemSolver":
        self.max_iterations = max_iter
        return self

    def with_tolerance(self, tol: float) -> "BemSolver":
        self.tolerance = tol
        return self

    def with_verbose(self, verbose: bool) -> "BemSolver":
        self.verbose = verbose
        return self

    # -- bem_solver.rs:320-350 ------------------------------------------------------------------
    def prepare_elements(self, problem: BemProblem) -> Mesh:
        import copy

        mesh = copy.deepcopy(problem.mesh)
        n = mesh.n_elem
        mesh.bc_val[:] = 0.0
        mesh.bc_len[:] = 1
        if problem.bc_type == BoundaryConditionType.Soft:
            mesh.bc_type[:] = BC_PRESSURE      # Pressure(vec![0])
        else:
            mesh.bc_type[:] = BC_VELOCITY      # Velocity(vec![0]) / VelocityWithAdmittance{velocity: vec![0], ..} (tbem.rs:236-238)
        mesh.dof[:] = np.arange(n, dtype=np.uint32)  # elem.dof_addresses = vec![i]
        return mesh

    # -- bem_solver.rs:273-317 ------------------------------------------------------------------
    def solve(self, problem: BemProblem, ctx: Optional[bem.Context] = None, stats: Optional[dict] = None) -> BemSolution:
        if self.assembly_method != AssemblyMethod.Tbem:
            raise BemError(f"Not implemented: {self.assembly_method.name} assembly is outside the dense path of this library")
        mesh = self.prepare_elements(problem)
        staged = bem.StagedMesh(mesh, ctx)
        beta = problem.physics.burton_miller_beta_scaled(self.beta_scale)
        system = bem.build_tbem_system_with_beta(staged, problem.physics, beta)
        n = mesh.num_dofs
        rows = mesh.is_eval == 0
        centers, normals = mesh.center[rows], mesh.normal[rows]
        if problem.use_burton_miller:
            inc = problem.incident_field.compute_rhs_with_beta(centers, normals, problem.physics, beta)
        else:
            inc = problem.incident_field.compute_rhs(centers, normals, problem.physics, False)
        rhs = system.rhs_full() + inc
        x = self.solve_dense_system(system, rhs, stats)
        if self.verbose:
            print(f"Solution complete. Max surface pressure: {np.abs(x).max():.6f}")
        return BemSolution(x, mesh, problem.incident_field, problem.physics, staged)

    # -- bem_solver.rs:435-463 ------------------------------------------------------------------
    def solve_dense_system(self, system, rhs: np.ndarray, stats: Optional[dict] = None) -> np.ndarray:
        if self.solver_method == SolverMethod.Direct:
            try:
                return bem.lu_solve(system, rhs, stats=stats)
            except bem.LuError as e:
                raise BemError(f"Solver failed: {e}") from e
        sol = bem.bicgstab(bem.DenseOperator(system), rhs, bem.BiCgstabConfig(self.max_iterations, self.tolerance, 0))
        if stats is not None:
            stats.update(iterations=sol.iterations, residual=sol.residual)
        if sol.converged:
            return sol.x
        raise BemError(f"Solver failed: BiCGSTAB did not converge: residual = {sol.residual}")
