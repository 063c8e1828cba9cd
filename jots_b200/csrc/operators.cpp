#include "operators.hpp"

namespace jots {

static bool is_const(const Poly& p) { return p.n == 1; }

JOTSIterator::JOTSIterator(Ctx* ctx, const SolverSpec& ls_) : c(ctx), ls(ls_) {
    b_vec.alloc(c->n);
    UpdateNeumann();
}

void JOTSIterator::UpdateNeumann() {
    // b.Assemble(); b.ParallelAssemble(b_vec)  (jots_iterator.cpp:39-43)
    c->upload_flux_values();
    c->neumann_vector(neumann_div, nullptr, b_vec.p);
}

std::unique_ptr<Krylov> JOTSIterator::make_solver() {
    std::unique_ptr<Krylov> s(new Krylov(c, ls.solver, ls.prec));
    s->rel_tol = ls.rel_tol; s->abs_tol = ls.abs_tol; s->max_iter = ls.max_iter; s->print = ls.print;
    return s;
}

void JOTSIterator::ode_step(double* u) {
    const int64_t n = c->n;
    cudaStream_t st = c->stream;
    kvec.alloc(n);
    if (time_scheme == TS_EULER_IMPLICIT) {
        ImplicitSolve(dt, u, kvec.p);
        v_axpy(st, n, dt, kvec.p, u);
    } else if (time_scheme == TS_EULER_EXPLICIT) {
        Mult(u, kvec.p);
        v_axpy(st, n, dt, kvec.p, u);
    } else if (time_scheme == TS_RK4) {
        for (auto& r : rk) r.alloc(n);
        utmp.alloc(n);
        Mult(u, rk[0].p);
        v_waxpby(st, n, 1.0, u, 0.5 * dt, rk[0].p, utmp.p);
        Mult(utmp.p, rk[1].p);
        v_waxpby(st, n, 1.0, u, 0.5 * dt, rk[1].p, utmp.p);
        Mult(utmp.p, rk[2].p);
        v_waxpby(st, n, 1.0, u, dt, rk[2].p, utmp.p);
        Mult(utmp.p, rk[3].p);
        v_axpy(st, n, dt / 6.0, rk[0].p, u);
        v_axpy(st, n, dt / 3.0, rk[1].p, u);
        v_axpy(st, n, dt / 3.0, rk[2].p, u);
        v_axpy(st, n, dt / 6.0, rk[3].p, u);
        c->stats.kernel_launches += 6;
    } else throw std::runtime_error("unknown time scheme");
    ++c->stats.kernel_launches;
    time += dt;
    ++c->stats.steps;
}

// ------------------------------------------------------------------------------------------------
// LinearConductionOperator
// ------------------------------------------------------------------------------------------------
LinearConductionOperator::LinearConductionOperator(Ctx* ctx, const SolverSpec& ls_, double dt_, int scheme)
    : JOTSIterator(ctx, ls_) {
    dt = dt_; time_scheme = scheme;
    expl_solver = make_solver();
    impl_solver = make_solver();
    rhs.alloc(c->n);
    // the constructor assembles everything once with the properties as they are now
    ReassembleM();
    ReassembleK();
    ReassembleA();
}

int LinearConductionOperator::coef_mode() const {
    const MatSet& m = c->mat;
    if (is_const(m.rho) && is_const(m.C) && is_const(m.k)) return CF_CONST;
    if (is_const(m.rho) && is_const(m.C)) return CF_K;
    return CF_LINFIELDS;
}

FormSpec LinearConductionOperator::M_spec() {
    // M = int rho_h C_h phi_i phi_j, essential rows/cols eliminated, diag 1 (:59, :76-92)
    FormSpec f;
    f.op = OP_LIN; f.cf = coef_mode();
    f.cm = 1.0; f.ck = 0.0;
    f.m0 = c->mat.rho.c[0] * c->mat.C.c[0]; f.a0 = c->mat.k.c[0]; f.inv_rc = 1.0;
    f.mask_in = f.mask_out = f.diag_one = true;
    f.state = c->mat_state.p;
    return f;
}
FormSpec LinearConductionOperator::K_spec() {
    // K = int k_h grad phi_i . grad phi_j, NOT eliminated (:63, :94-107)
    FormSpec f = M_spec();
    f.cm = 0.0; f.ck = 1.0;
    f.mask_in = f.mask_out = f.diag_one = false;
    return f;
}
FormSpec LinearConductionOperator::A_spec() {
    // A = M + dt K, eliminated (:109-120)
    FormSpec f = M_spec();
    f.ck = dt;
    return f;
}

void LinearConductionOperator::ReassembleM() {
    M_op.reset(new FormOp(c, M_spec()));
    expl_solver->set_operator(M_op.get());
}
void LinearConductionOperator::ReassembleK() {}
void LinearConductionOperator::ReassembleA() {
    A_op.reset(new FormOp(c, A_spec()));
    impl_solver->set_operator(A_op.get());
}

void LinearConductionOperator::ProcessMatPropUpdate(int mp) {
    // :180-194
    if (mp == MP_DENSITY || mp == MP_SPECIFIC_HEAT) { ReassembleM(); ReassembleA(); }
    else if (mp == MP_THERMAL_CONDUCTIVITY) { ReassembleK(); ReassembleA(); }
}

void LinearConductionOperator::CalculateRHS(const double* u, double* out) {
    // rhs = -K u + b_vec (:122-132)
    FormOp K(c, K_spec());
    K.mult(u, out);
    v_waxpby(c->stream, c->n, -1.0, out, 1.0, b_vec.p, out);
    ++c->stats.kernel_launches;
}

void LinearConductionOperator::Mult(const double* u, double* k) {
    CalculateRHS(u, rhs.p);
    c->zero_ess(rhs.p);
    expl_solver->mult(rhs.p, k);
}

void LinearConductionOperator::ImplicitSolve(double, const double* u, double* k) {
    CalculateRHS(u, rhs.p);
    c->zero_ess(rhs.p);
    impl_solver->mult(rhs.p, k);
}

void LinearConductionOperator::Iterate(double* u) { ode_step(u); }

// ------------------------------------------------------------------------------------------------
// NonlinearConductionOperator
// ------------------------------------------------------------------------------------------------
NonlinearConductionOperator::NonlinearConductionOperator(Ctx* ctx, const SolverSpec& ls_, const NewtonSpec& nw_, double dt_, int scheme)
    : JOTSIterator(ctx, ls_), nw(nw_), newton(ctx) {
    dt = dt_; time_scheme = scheme;
    const MatSet& m = c->mat;
    all_const = is_const(m.rho) && is_const(m.C) && is_const(m.k);
    has_BN = !(is_const(m.rho) && is_const(m.C));
    if (!has_BN) {
        // Neumann term re-created with g / (rho C) (:230-234)
        inv_rc = 1.0 / (m.rho.c[0] * m.C.c[0]);
        neumann_div = m.rho.c[0] * m.C.c[0];
        UpdateNeumann();
    }
    lin_solver = make_solver();
    newton.rel_tol = nw.rel_tol; newton.abs_tol = nw.abs_tol; newton.max_iter = nw.max_iter; newton.print = nw.print;
    z.alloc(c->n); tmp.alloc(c->n); tmp2.alloc(c->n);
    // unit mass, FormSystemMatrix -> eliminated (:189-192)
    FormSpec ms;
    ms.op = OP_LIN; ms.cf = CF_CONST; ms.cm = 1.0; ms.ck = 0.0; ms.m0 = 1.0; ms.a0 = 0.0;
    ms.mask_in = ms.mask_out = ms.diag_one = true;
    M_op.reset(new FormOp(c, ms));
}

void NonlinearConductionOperator::new_state(const double* k) {
    // JOTSNewtonSolver::ProcessNewState: z = u_n + dt k; the nodal polynomials are re-evaluated
    // from z inside the kernels (jots_common.inl:124-147)
    v_waxpby(c->stream, c->n, 1.0, u_n, dt, k, z.p);
    ++c->stats.kernel_launches;
}

// y = M k - A(z), y[ess] = 0 with A(z) = -kappa(z) + BN(z) | -K z + N   (:27-39, :70-79)
void NonlinearConductionOperator::mult(const double* k, double* y) {
    const MatSet& m = c->mat;
    FormSpec f;
    f.op = OP_RES;
    f.cm = 1.0; f.ck = 1.0; f.cb = -1.0; f.m0 = 1.0;
    f.mask_in = true; f.mask_out = true;
    f.state = z.p;
    if (all_const) { f.cf = CF_CONST; f.a0 = m.k.c[0] * inv_rc; }
    else if (!has_BN) { f.cf = CF_K; f.inv_rc = inv_rc; }
    else f.cf = CF_FULL;
    c->apply_form(f, k, y);
    if (has_BN) {
        if (c->any_flux()) {
            // boundary part of BN(z): int (g / (rho_h C_h)) phi_i (jots_nlfis.cpp:51-54), assembled and
            // interface-summed on its own, then subtracted
            c->neumann_vector(1.0, z.p, tmp2.p, true, 1.0, true);
            v_axpy(c->stream, c->n, -1.0, tmp2.p, y);
            ++c->stats.kernel_launches;
        }
    } else {
        v_axpy(c->stream, c->n, -1.0, b_vec.p, y);
        ++c->stats.kernel_launches;
    }
    c->zero_ess(y);
}

void NonlinearConductionOperator::A_mult(const double* zz, double* y) {
    // A(z) alone: -(R(0) at state z) without the mass term
    c->copy(zz, z.p);
    c->zero(tmp.p);
    const MatSet& m = c->mat;
    FormSpec f;
    f.op = OP_RES; f.cm = 0.0; f.ck = -1.0; f.cb = 1.0; f.m0 = 1.0;
    f.mask_in = true; f.mask_out = !all_const;
    f.state = z.p;
    // explicit Mult never refreshes the property GridFunctions (nl_conduction_operator.cpp:281-297): the polynomials stay
    // evaluated on the vector of the last UpdateMatProps / ProcessNewState, only the gradient follows the stage vector
    const double* cstate = c->mat_state.p ? c->mat_state.p : z.p;
    f.coef_state = cstate;
    if (all_const) { f.cf = CF_CONST; f.a0 = m.k.c[0] * inv_rc; }
    else if (!has_BN) { f.cf = CF_K; f.inv_rc = inv_rc; }
    else f.cf = CF_FULL;
    c->apply_form(f, tmp.p, y);
    if (has_BN) {
        if (c->any_flux()) {
            c->neumann_vector(1.0, cstate, tmp2.p, true, 1.0, true);
            v_axpy(c->stream, c->n, 1.0, tmp2.p, y);
            ++c->stats.kernel_launches;
        }
    } else { v_axpy(c->stream, c->n, 1.0, b_vec.p, y); ++c->stats.kernel_launches; }
}

FormSpec NonlinearConductionOperator::jacobian_spec(const double* zz) {
    // J = M - dt dA/du at z, essential rows/cols eliminated (:41-67, :81-90)
    const MatSet& m = c->mat;
    FormSpec f;
    f.cm = 1.0; f.ck = dt; f.cb = -dt; f.m0 = 1.0;
    f.mask_in = f.mask_out = f.diag_one = true;
    f.state = zz;
    if (all_const) { f.op = OP_LIN; f.cf = CF_CONST; f.a0 = m.k.c[0] * inv_rc; }
    else if (!has_BN) { f.op = OP_JAC; f.cf = CF_K; f.inv_rc = inv_rc; }
    else { f.op = OP_JAC; f.cf = CF_FULL; f.bdr_mass_scale = -dt; }
    return f;
}

LinOp* NonlinearConductionOperator::gradient(const double*) {
    J_op.reset(new FormOp(c, jacobian_spec(z.p)));
    return J_op.get();
}

void NonlinearConductionOperator::ImplicitSolve(double, const double* u, double* k) {
    u_n = u;
    c->zero(k);  // iterative_mode = false (:260)
    newton.mult(*this, *lin_solver, nullptr, k);
    if (!newton.converged) throw std::runtime_error("Newton solver did not converge.");  // MFEM_VERIFY (:312)
}

void NonlinearConductionOperator::Mult(const double* u, double* k) {
    // rhs = A(u); rhs[ess] = 0; solve M k = rhs (:281-297)
    A_mult(u, tmp.p);
    c->zero_ess(tmp.p);
    lin_solver->set_operator(M_op.get());
    lin_solver->mult(tmp.p, k);
}

void NonlinearConductionOperator::Iterate(double* u) { ode_step(u); }

// ------------------------------------------------------------------------------------------------
// SteadyConductionOperator
// ------------------------------------------------------------------------------------------------
SteadyConductionOperator::SteadyConductionOperator(Ctx* ctx, const SolverSpec& ls_, const NewtonSpec& nw_)
    : JOTSIterator(ctx, ls_), nw(nw_), newton(ctx) {
    lin_solver = make_solver();
    newton.rel_tol = nw.rel_tol; newton.abs_tol = nw.abs_tol; newton.max_iter = nw.max_iter; newton.print = nw.print;
    rhs.alloc(c->n);
}

void SteadyConductionOperator::mult(const double* u, double* y) {
    // k(u)_i = int k_h(u) grad u . grad phi_i, essential rows zeroed (:19-20)
    FormSpec f;
    f.op = OP_RES; f.cf = is_const(c->mat.k) ? CF_CONST : CF_K;
    f.cm = 0.0; f.ck = 1.0; f.cb = 0.0; f.m0 = 1.0; f.a0 = c->mat.k.c[0]; f.inv_rc = 1.0;
    f.mask_in = false; f.mask_out = true;
    f.state = u;
    c->apply_form(f, u, y);
}

LinOp* SteadyConductionOperator::gradient(const double* u) {
    FormSpec f;
    f.cm = 0.0; f.ck = 1.0; f.cb = 0.0; f.m0 = 1.0; f.a0 = c->mat.k.c[0]; f.inv_rc = 1.0;
    if (is_const(c->mat.k)) { f.op = OP_LIN; f.cf = CF_CONST; }
    else { f.op = OP_JAC; f.cf = CF_K; }
    f.mask_in = f.mask_out = f.diag_one = true;
    f.state = u;
    J_op.reset(new FormOp(c, f));
    return J_op.get();
}

void SteadyConductionOperator::Iterate(double* u) {
    // newton.Mult(b_vec, u), iterative_mode = true (:34-38)
    c->copy(b_vec.p, rhs.p);
    if (zero_rhs_on_essential) c->zero_ess(rhs.p);
    newton.mult(*this, *lin_solver, rhs.p, u);
    ++c->stats.steps;
}

}  // namespace jots
