"""Oracle (test infrastructure): the reference's conduction operators and driver loop.

Restates, on top of ``oracle.fem`` / ``oracle.solvers``:

* ``Config``                          /root/reference/src/config_file.cpp:7-172 (keys, defaults, quirks)
* ``MaterialProperty``                /root/reference/src/material_property.cpp:6-108
* BC value objects                    /root/reference/src/boundary_condition.cpp:49-75
* ``JOTSIterator``                    /root/reference/src/jots_iterator.cpp:5-43
* ``LinearConductionOperator``        /root/reference/src/linear_conduction_operator.cpp:15-194
* ``NonlinearConductionOperator``     /root/reference/src/nl_conduction_operator.cpp:27-323
* ``SteadyConductionOperator``        /root/reference/src/steady_conduction_operator.cpp:3-38
* ``JOTSDriver`` (hot-path callers)   /root/reference/src/jots_driver.cpp:568-725
* ODE solvers                         MFEM BackwardEuler / ForwardEuler / RK4
"""
from __future__ import annotations

import configparser
import math
import os

import numpy as np
import scipy.sparse as sp

from . import fem, mesh as M, solvers as S


# ----------------------------------------------------------------------------------------------
# Config
# ----------------------------------------------------------------------------------------------

def _split(s):
    return [t.strip() for t in s.split(",")]


def _get_int(sec, key, default):
    """boost ptree ``get(key, int default)``: silently falls back to the default when the
    text is not an integer (e.g. Max_Timesteps=5e6)."""
    if sec is None or key not in sec:
        return default
    try:
        return int(sec[key].strip())
    except ValueError:
        return default


def _get_float(sec, key, default):
    if sec is None or key not in sec:
        return default
    try:
        return float(sec[key].strip())
    except ValueError:
        return default


class Config:
    def __init__(self, path):
        cp = configparser.ConfigParser(interpolation=None, delimiters=("=",), comment_prefixes=(";", "#"),
                                       inline_comment_prefixes=None)
        cp.optionxform = str
        with open(path) as f:
            cp.read_file(f)
        self.path = path
        g = lambda name: cp[name] if cp.has_section(name) else None
        fe = g("FiniteElementSetup")
        self.sim_type = fe.get("Simulation_Type", "Unsteady") if fe else "Unsteady"
        self.use_restart = (fe.get("Use_Restart", "No").strip() == "Yes") if fe else False
        self.mesh_file = fe.get("Mesh_File", "mesh_name.mesh").strip()
        self.fe_order = _get_int(fe, "FE_Order", 1)
        self.serial_refine = _get_int(fe, "Serial_Refine", 0)
        self.parallel_refine = _get_int(fe, "Parallel_Refine", 0)
        self.initial_temp = _get_float(fe, "Initial_Temperature", 100.0)
        self.mat_props = {k: _split(v) for k, v in cp["MaterialProperties"].items()}
        self.with_precice = cp.has_section("preCICE")
        self.bcs = {}
        for k, v in cp["BoundaryConditions"].items():
            self.bcs[int(k.split("_")[-1])] = _split(v)
        ti = g("TimeIntegration")
        self.using_time_integration = ti is not None
        self.time_scheme = ti.get("Time_Scheme", "Euler_Implicit").strip() if ti else "Euler_Implicit"
        self.dt = _get_float(ti, "Delta_Time", 0.1)
        self.max_timesteps = _get_int(ti, "Max_Timesteps", 100)
        ls = g("LinearSolverSettings")
        self.solver = ls.get("Solver", "FGMRES").strip() if ls else "FGMRES"
        self.prec = ls.get("Preconditioner", "Chebyshev").strip() if ls else "Chebyshev"
        self.abs_tol = _get_float(ls, "Absolute_Tolerance", 1e-16)
        self.rel_tol = _get_float(ls, "Relative_Tolerance", 1e-10)
        self.max_iter = _get_int(ls, "Max_Iterations", 100)
        nw = g("NewtonSolverSettings")
        self.using_newton = nw is not None
        self.newton_max_iter = _get_int(nw, "Max_Iterations", 100) if nw else 0
        self.newton_abs_tol = _get_float(nw, "Absolute_Tolerance", 1e-16)
        self.newton_rel_tol = _get_float(nw, "Relative_Tolerance", 1e-10)


# ----------------------------------------------------------------------------------------------
# material properties and boundary conditions
# ----------------------------------------------------------------------------------------------

class MaterialProperty:
    """Uniform / Polynomial property.  ``coeff`` etc. are scalars (Uniform) or nodal fields
    (Polynomial: the polynomial is evaluated at the NODES, highest power first, by repeated
    multiplication: material_property.cpp:36-81)."""

    def __init__(self, info):
        self.model = info[0]
        if self.model == "Uniform":
            self.value = float(info[1])
            self.poly = None
        elif self.model == "Polynomial":
            self.poly = [float(t) for t in info[1:]]
            if len(self.poly) < 2:
                raise ValueError("PolynomialProperty needs >= 2 coefficients (material_property.cpp:70)")
        else:
            raise ValueError("Unknown/Invalid material model specified")
        self.coeff = self.value if self.poly is None else None
        self.dcoeff = 0.0
        self.d2coeff = 0.0

    def is_constant(self):
        return self.poly is None

    def local_value(self, t):
        """GetLocalValue (material_property.cpp:101-108): the property at one temperature."""
        if self.poly is None:
            return self.value
        n = len(self.poly)
        return sum(c * t ** (n - 1 - i) for i, c in enumerate(self.poly))

    def update(self, u):
        if self.poly is None:
            return
        c = self.poly
        n = len(c)
        z1 = np.zeros_like(u)
        for i in range(n):
            z2 = np.ones_like(u)
            for _ in range(n - i - 1):
                z2 = z2 * u
            z1 = z1 + z2 * c[i]
        self.coeff = z1
        z1 = np.zeros_like(u)
        for i in range(n - 1):
            z2 = np.full_like(u, float(n - 1 - i))
            for _ in range(n - i - 2):
                z2 = z2 * u
            z1 = z1 + z2 * c[i]
        self.dcoeff = z1
        z1 = np.zeros_like(u)
        for i in range(n - 2):
            z2 = np.full_like(u, float((n - 1 - i) * (n - 2 - i)))
            for _ in range(n - i - 3):
                z2 = z2 * u
            z1 = z1 + z2 * c[i]
        self.d2coeff = z1


class BoundaryCondition:
    def __init__(self, attr, info):
        self.attr = attr
        kind = info[0]
        self.kind = kind
        if kind in ("Isothermal", "HeatFlux"):
            self.value = float(info[1])
            self.constant = True
        elif kind in ("Sinusoidal_Isothermal", "Sinusoidal_HeatFlux"):
            self.amp, self.afreq, self.phase, self.shift = (float(t) for t in info[1:5])
            self.constant = False
            self.value = None
        elif kind in ("preCICE_Isothermal", "preCICE_HeatFlux"):
            # PreciceBC (boundary_condition.cpp:77-147): info = kind, coupling-mesh name, default value; the coefficient is
            # a nodal field (GridFunctionCoefficient) whose boundary values the adapter overwrites (set_read_data)
            self.mesh_name, self.default = info[1], float(info[2])
            self.constant = False
            self.value = self.default      # scalar until data is read, then a nodal field (space attached by the driver)
        else:
            raise ValueError("Invalid/Unknown boundary condition specified: '%s'" % kind)
        self.essential = kind.endswith("Isothermal")
        self.precice = kind.startswith("preCICE")

    def update(self, t):
        if not self.constant and not self.precice:
            self.value = self.amp * math.sin(self.afreq * t + self.phase) + self.shift

    def set_read_data(self, space, values):
        """PreciceBC::UpdateCoeff: coeff_gf.SetSubVector(bdr_dof_indices, read_data_arr) (later entries win)."""
        dofs, _ = space.bdr_nodes(self.attr)
        field = np.full(space.ndof, self.default) if np.isscalar(self.value) else self.value
        field[dofs] = np.asarray(values, dtype=float)
        self.value = field

    def write_data(self, space, u, k_prop):
        """RetrieveWriteData: wall heat flux for the isothermal kind (:152-191), boundary temperatures otherwise (:194-198)."""
        if self.essential:
            return space.wall_heat_flux(self.attr, u, lambda t: k_prop.local_value(t))
        return u[space.bdr_nodes(self.attr)[0]]


# ----------------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------------

class JOTSIterator:
    def __init__(self, cfg, space, bcs):
        self.cfg, self.space, self.bcs = cfg, space, bcs
        attrs = space.mesh.bdr_attributes
        # dbc_bdr[i] = 1 for essential BC i; marker index i <-> attribute i+1 (jots_iterator.cpp:12-25)
        self.ess_attrs = [i + 1 for i, bc in enumerate(bcs) if bc.essential]
        self.ess = space.essential_dofs(self.ess_attrs)
        self.neumann_divisor = None
        self.update_neumann()

    def neumann_values(self):
        return {bc.attr: bc.value for bc in self.bcs if not bc.essential}

    def update_neumann(self):
        self.b_vec = self.space.neumann_vector(self.neumann_values(), self.neumann_divisor)

    def make_lin_solver(self):
        s = S.make_solver(self.cfg.solver, self.cfg.rel_tol, self.cfg.abs_tol, self.cfg.max_iter)
        s.prec = S.make_prec(self.cfg.prec)
        return s


class _ODE:
    """BackwardEuler / ForwardEuler / RK4 ``Step`` (App. A.6)."""

    @staticmethod
    def step(op, scheme, u, t, dt):
        if scheme == "Euler_Implicit":
            k = op.implicit_solve(dt, u)
            return u + dt * k, t + dt
        if scheme == "Euler_Explicit":
            k = op.explicit_mult(u)
            return u + dt * k, t + dt
        if scheme == "RK4":
            k1 = op.explicit_mult(u)
            k2 = op.explicit_mult(u + 0.5 * dt * k1)
            k3 = op.explicit_mult(u + 0.5 * dt * k2)
            k4 = op.explicit_mult(u + dt * k3)
            return u + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4), t + dt
        raise KeyError(scheme)


class LinearConductionOperator(JOTSIterator):
    def __init__(self, cfg, space, bcs, rho, C, k, dt):
        super().__init__(cfg, space, bcs)
        self.rho, self.C, self.k, self.dt = rho, C, k, dt
        self.expl_solver = self.make_lin_solver()
        self.impl_solver = self.make_lin_solver()
        self.stats = dict(krylov_iters=0, applies=0, steps=0)
        self.reassemble_M()
        self.reassemble_K()
        self.reassemble_A()

    def _rhoC(self):
        if self.rho.is_constant() and self.C.is_constant():
            return self.rho.coeff * self.C.coeff
        return (self.rho.coeff, self.C.coeff)

    def reassemble_M(self):
        rc = self._rhoC()
        sp_ = self.space
        if isinstance(rc, tuple):  # ProductCoefficient of two interpolated fields
            cq = sp_._coef(rc[0]) * sp_._coef(rc[1])
            wq = sp_.W[None, :] * sp_.detJ * cq
            el = np.einsum("eq,qi,qj->eij", wq, sp_.Phi, sp_.Phi, optimize=True)
            Mm = sp_._scatter_matrix(el)
        else:
            Mm = sp_.mass_matrix(rc)
        self.M_mat = fem.eliminate_rows_cols(Mm, self.ess)
        self.expl_solver.set_operator(self.M_mat)

    def reassemble_K(self):
        self.K_mat = self.space.diffusion_matrix(self.k.coeff)  # NOT eliminated (:106)

    def reassemble_A(self):
        self.A_mat = fem.eliminate_rows_cols(self.M_mat + self.dt * self.K_mat, self.ess)
        self.impl_solver.set_operator(self.A_mat)

    def process_mat_prop_update(self, which):
        if which in ("rho", "C"):
            self.reassemble_M()
            self.reassemble_A()
        else:
            self.reassemble_K()
            self.reassemble_A()

    def rhs(self, u):
        y = -(self.K_mat @ u) + self.b_vec
        y[self.ess] = 0.0
        return y

    def explicit_mult(self, u):
        k = self.expl_solver.mult(self.rhs(u))
        self._acc(self.expl_solver)
        return k

    def implicit_solve(self, dt, u):
        k = self.impl_solver.mult(self.rhs(u))
        self._acc(self.impl_solver)
        return k

    def _acc(self, s):
        self.stats["krylov_iters"] += s.info.iterations
        self.stats["applies"] += s.info.applies

    def iterate(self, u, t):
        u, t = _ODE.step(self, self.cfg.time_scheme, u, t, self.dt)
        self.stats["steps"] += 1
        return u, t


class _CoeffSet:
    """The six derived coefficients of NonlinearConductionOperator
    (/root/reference/src/nl_conduction_operator.cpp:92-163) evaluated at a set of points through
    an interpolation callback ``at(c)`` (c scalar or nodal field): every factor is an interpolated
    nodal field, ratios are ratios of interpolants."""

    def __init__(self, rho, C, k):
        self.rho, self.C, self.k = rho, C, k

    def alpha(self, at):
        return at(self.k.coeff) / (at(self.rho.coeff) * at(self.C.coeff))

    def dalpha(self, at):
        k, dk = at(self.k.coeff), at(self.k.dcoeff)
        r, dr = at(self.rho.coeff), at(self.rho.dcoeff)
        c, dc = at(self.C.coeff), at(self.C.dcoeff)
        return dk / (r * c) - k * dr / (r * r * c) - k * dc / (c * c * r)

    def neg_beta(self, at):
        k = at(self.k.coeff)
        r, dr = at(self.rho.coeff), at(self.rho.dcoeff)
        c, dc = at(self.C.coeff), at(self.C.dcoeff)
        return -(k * (-dr / (r * r * c) - dc / (c * c * r)))

    def neg_dbeta(self, at):
        k, dk = at(self.k.coeff), at(self.k.dcoeff)
        r, dr, d2r = at(self.rho.coeff), at(self.rho.dcoeff), at(self.rho.d2coeff)
        c, dc, d2c = at(self.C.coeff), at(self.C.dcoeff), at(self.C.d2coeff)
        return -(dk * (-dr / (r * r * c) - dc / (c * c * r))
                 + k * (-d2c / (c * c * r) + 2 * dc * dr / (c * c * r * r) + 2 * dc * dc / (c * c * c * r)
                        - d2r / (c * r * r) + 2 * dr * dr / (c * r * r * r)))

    def g_over_rhoC(self, g, at):
        return at(g) / (at(self.rho.coeff) * at(self.C.coeff))

    def dg_over_rhoC(self, g, at):
        r, dr = at(self.rho.coeff), at(self.rho.dcoeff)
        c, dc = at(self.C.coeff), at(self.C.dcoeff)
        g = at(g)
        return -g * dr / (r * r * c) - g * dc / (c * c * r)


class NonlinearConductionOperator(JOTSIterator):
    def __init__(self, cfg, space, bcs, rho, C, k, dt):
        super().__init__(cfg, space, bcs)
        self.rho, self.C, self.k, self.dt = rho, C, k, dt
        self.cs = _CoeffSet(rho, C, k)
        sp_ = space
        self.M_mat = fem.eliminate_rows_cols(sp_.mass_matrix(1.0), self.ess)  # FormSystemMatrix (:189-192)
        self.all_const = rho.is_constant() and C.is_constant() and k.is_constant()
        self.has_BN = not (rho.is_constant() and C.is_constant())
        if self.all_const:
            self.K_mat = sp_.diffusion_matrix(k.coeff / (rho.coeff * C.coeff))
        if not self.has_BN:
            # Neumann term re-created with g / (rho C) (:230-234)
            self.neumann_divisor = (rho.coeff, C.coeff)
            self.update_neumann()
        self.lin_solver = self.make_lin_solver()
        self.newton = S.NewtonSolver(cfg.newton_rel_tol, cfg.newton_abs_tol, cfg.newton_max_iter)
        self.stats = dict(krylov_iters=0, applies=0, newton_iters=0, residual_evals=0, steps=0)
        self.u_n = None

    # material state (JOTSNewtonSolver::ProcessNewState) ------------------------------------
    def _update_props(self, z):
        for mp in (self.k, self.C, self.rho):
            mp.update(z)

    def _bdr_g(self):
        return self.space.face_nodal_values(self.neumann_values())     # [nb, nfl]: uniform values or nodal (preCICE) fields

    # A(u) and its gradient ----------------------------------------------------------------------
    def A_mult(self, u):
        sp_ = self.space
        if self.all_const:
            y = -(self.K_mat @ u)
        else:
            y = self._kappa(u)
            y[self.ess] = 0.0  # ParNonlinearForm::Mult zeroes essential entries
            y = -y
        if self.has_BN:
            bn = self._BN(u)
            bn[self.ess] = 0.0
            y = y + bn
        else:
            y = y + self.b_vec
        return y

    def _kappa(self, u):
        sp_ = self.space
        wq = sp_.W[None, :] * sp_.detJ * self.cs.alpha(sp_._coef)
        gu = sp_.grad_at_quad(u)
        dP = sp_.phys_dPhi()
        el = np.einsum("eq,eqi,eqli->el", wq, gu, dP, optimize=True)
        return sp_._scatter_vector(el)

    def _BN(self, u):
        sp_ = self.space
        gu = sp_.grad_at_quad(u)
        wq = sp_.W[None, :] * sp_.detJ * self.cs.neg_beta(sp_._coef) * np.einsum("eqi,eqi->eq", gu, gu)
        y = sp_._scatter_vector(wq @ sp_.Phi)
        g = self._bdr_g()
        y += sp_.bdr_coef_vector(lambda at: self.cs.g_over_rhoC(g, at), fem.bdr_lf_quad_points(sp_.p))
        return y

    def A_gradient(self, u):
        sp_ = self.space
        if self.all_const:
            return -self.K_mat
        # K gradient (jots_nlfis.cpp:24-40), essential rows/cols eliminated by ParNonlinearForm
        alpha_q = self.cs.alpha(sp_._coef)
        dalpha_q = self.cs.dalpha(sp_._coef)
        Kg = self._diff_matrix_q(alpha_q) + self._weakdiv_q(dalpha_q, u)
        Kg = fem.eliminate_rows_cols(Kg, self.ess)
        J = -Kg
        if self.has_BN:
            nb_q = self.cs.neg_beta(sp_._coef)
            ndb_q = self.cs.neg_dbeta(sp_._coef)
            Bg = 2.0 * self._conv_q(nb_q, u) + self._gradsq_mass_q(ndb_q, u)
            g = self._bdr_g()
            Bg = Bg + sp_.bdr_mass_matrix(lambda at: self.cs.dg_over_rhoC(g, at))
            Bg = fem.eliminate_rows_cols(Bg, self.ess)
            J = J + Bg
        return J

    def _diff_matrix_q(self, cq):
        sp_ = self.space
        dP = sp_.phys_dPhi()
        el = np.einsum("eq,eqli,eqmi->elm", sp_.W[None, :] * sp_.detJ * cq, dP, dP, optimize=True)
        return sp_._scatter_matrix(el)

    def _weakdiv_q(self, cq, z):
        sp_ = self.space
        t = np.einsum("eqli,eqi->eql", sp_.phys_dPhi(), sp_.grad_at_quad(z))
        el = np.einsum("eq,eql,qm->elm", sp_.W[None, :] * sp_.detJ * cq, t, sp_.Phi, optimize=True)
        return sp_._scatter_matrix(el)

    def _conv_q(self, cq, z):
        sp_ = self.space
        t = np.einsum("eqmi,eqi->eqm", sp_.phys_dPhi(), sp_.grad_at_quad(z))
        el = np.einsum("eq,ql,eqm->elm", sp_.W[None, :] * sp_.detJ * cq, sp_.Phi, t, optimize=True)
        return sp_._scatter_matrix(el)

    def _gradsq_mass_q(self, cq, z):
        sp_ = self.space
        gz = sp_.grad_at_quad(z)
        wq = sp_.W[None, :] * sp_.detJ * cq * np.einsum("eqi,eqi->eq", gz, gz)
        el = np.einsum("eq,qi,qj->eij", wq, sp_.Phi, sp_.Phi, optimize=True)
        return sp_._scatter_matrix(el)

    # R(k) (ReducedSystemOperatorR, :70-90) as the operator handed to Newton -------------------
    def new_state(self, kvec):
        self._update_props(self.u_n + self.dt * kvec)

    def mult(self, kvec):
        z = self.u_n + self.dt * kvec
        y = -self.A_mult(z) + self.M_mat @ kvec
        y[self.ess] = 0.0
        return y

    def gradient(self, kvec):
        z = self.u_n + self.dt * kvec
        J = self.M_mat + (-self.dt) * self.A_gradient(z)
        return fem.eliminate_rows_cols(J, self.ess)

    def implicit_solve(self, dt, u):
        self.u_n = u
        k0 = np.zeros_like(u)
        k = self.newton.mult(self, self.lin_solver, np.zeros_like(u), k0)
        if not self.newton.converged:
            raise RuntimeError("Newton solver did not converge.")
        st = self.stats
        st["krylov_iters"] += self.newton.linear_iterations
        st["applies"] += self.newton.linear_applies
        st["newton_iters"] += self.newton.iterations
        st["residual_evals"] += self.newton.residual_evals
        return k

    def explicit_mult(self, u):
        rhs = self.A_mult(u)
        rhs[self.ess] = 0.0
        self.lin_solver.set_operator(self.M_mat)
        return self.lin_solver.mult(rhs)

    def process_mat_prop_update(self, which):
        pass

    def iterate(self, u, t):
        u, t = _ODE.step(self, self.cfg.time_scheme, u, t, self.dt)
        self.stats["steps"] += 1
        return u, t


class SteadyConductionOperator(JOTSIterator):
    """Newton on k(u) = b_vec with JOTSNonlinearDiffusionIntegrator(k, k')."""

    def __init__(self, cfg, space, bcs, k):
        super().__init__(cfg, space, bcs)
        self.k = k
        self.lin_solver = self.make_lin_solver()
        self.newton = S.NewtonSolver(cfg.newton_rel_tol, cfg.newton_abs_tol, cfg.newton_max_iter)
        self.stats = dict(krylov_iters=0, applies=0, newton_iters=0, residual_evals=0, steps=0)

    def new_state(self, u):
        self.k.update(u)

    def mult(self, u):
        y = self.space.diffusion_action(self.k.coeff, u)
        y[self.ess] = 0.0
        return y

    def gradient(self, u):
        sp_ = self.space
        J = sp_.diffusion_matrix(self.k.coeff) + sp_.weakdiv_term_matrix(self.k.dcoeff, u)
        return fem.eliminate_rows_cols(J, self.ess)

    def iterate(self, u, t):
        # reference quirk: b_vec is NOT zeroed on essential rows
        u = self.newton.mult(self, self.lin_solver, self.b_vec, u.copy())
        st = self.stats
        st["krylov_iters"] += self.newton.linear_iterations
        st["applies"] += self.newton.linear_applies
        st["newton_iters"] += self.newton.iterations
        st["residual_evals"] += self.newton.residual_evals
        st["steps"] += 1
        return u, t

    def process_mat_prop_update(self, which):
        pass


# ----------------------------------------------------------------------------------------------
# driver
# ----------------------------------------------------------------------------------------------

_PROP_KEYS = [("Density_rho", "rho"), ("Specific_Heat_C", "C"), ("Thermal_Conductivity_k", "k")]


class JOTSDriver:
    def __init__(self, cfg: Config, mesh: M.Mesh = None, base_dir=None):
        self.cfg = cfg
        if cfg.with_precice or cfg.use_restart:
            raise NotImplementedError("preCICE / restart paths are out of scope (SURVEY.md 2.1 #12,#13)")
        if mesh is None:
            path = cfg.mesh_file
            if not os.path.isabs(path) and base_dir:
                path = os.path.join(base_dir, path)
            mesh = M.read_mfem_mesh(path)
            for _ in range(cfg.serial_refine + cfg.parallel_refine):
                mesh = M.uniform_refine(mesh)
        self.mesh = mesh
        self.space = fem.H1Space(mesh, cfg.fe_order)
        self.u = np.full(self.space.ndof, cfg.initial_temp)
        self.props = {}
        for key, short in _PROP_KEYS:
            self.props[short] = MaterialProperty(cfg.mat_props[key]) if key in cfg.mat_props else None
        attrs = list(mesh.bdr_attributes)
        if len(attrs) != len(cfg.bcs):
            raise ValueError("Input file BC count and mesh BC counts do not match.")
        for a in attrs:
            if int(a) not in cfg.bcs:
                raise ValueError("No matching boundary attribute in config file for attribute %d" % a)
        self.bcs = [BoundaryCondition(int(a), cfg.bcs[int(a)]) for a in attrs]
        self.initialized_bcs = False
        if cfg.using_time_integration:
            self.it_num, self.max_timesteps, self.dt = 0, cfg.max_timesteps, cfg.dt
        else:
            self.it_num, self.max_timesteps, self.dt = -2, -1, 0.0
        self.time = 0.0
        self.update_mat_props(False)
        # BC coefficients must have a value before the iterator assembles its Neumann form
        for bc in self.bcs:
            bc.update(self.time)
        st = cfg.sim_type
        p = self.props
        if st == "Linearized_Unsteady":
            self.it = LinearConductionOperator(cfg, self.space, self.bcs, p["rho"], p["C"], p["k"], self.dt)
        elif st == "Nonlinear_Unsteady":
            self.it = NonlinearConductionOperator(cfg, self.space, self.bcs, p["rho"], p["C"], p["k"], self.dt)
        elif st == "Steady":
            self.it = SteadyConductionOperator(cfg, self.space, self.bcs, p["k"])
        else:
            raise KeyError(st)

    def update_mat_props(self, apply_changes):
        for short in ("rho", "C", "k"):  # MATERIAL_PROPERTY enum order (option_structure.hpp:18-23)
            mp = self.props[short]
            if mp is not None and not mp.is_constant():
                mp.update(self.u)
                if apply_changes:
                    self.it.process_mat_prop_update(short)

    def update_and_apply_bcs(self):
        n_changed = False
        for i, bc in enumerate(self.bcs):
            if (not bc.constant) or (not self.initialized_bcs):
                bc.update(self.time)
                if bc.essential:
                    # marker i <-> attribute i+1 (jots_driver.cpp:488-495)
                    dofs = self.space.essential_dofs([i + 1])
                    self.u[dofs] = bc.value if np.isscalar(bc.value) else bc.value[dofs]
                else:
                    n_changed = True
        self.initialized_bcs = True
        if n_changed:
            self.it.update_neumann()

    def step(self):
        self.update_and_apply_bcs()
        self.update_mat_props(True)
        self.u, self.time = self.it.iterate(self.u, self.time)
        self.it_num += 1
        if self.u.max() > 1e10:
            raise RuntimeError("JOTS has blown up")

    def run(self, max_steps=None):
        n = 0
        while self.it_num < self.max_timesteps:
            self.step()
            n += 1
            if max_steps is not None and n >= max_steps:
                break
        return self.u
