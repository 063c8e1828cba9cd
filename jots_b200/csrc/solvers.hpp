// Device-resident Krylov / Newton solvers with MFEM 4.7 control flow and stopping rules
// (SURVEY.md App. A.4-A.5): the objects Factory::GetSolver / GetPrec build in the reference
// (/root/reference/src/jots_common.inl:7-38) and JOTSNewtonSolver (jots_common.inl:113-147).
// The host sequences iterations and reads back 1-3 scalars per iteration; all vectors stay in HBM.
#pragma once
#include <memory>
#include <vector>

#include "ctx.hpp"

namespace jots {

enum { SOLVER_CG = 0, SOLVER_GMRES = 1, SOLVER_FGMRES = 2 };
enum { PREC_JACOBI = 0, PREC_CHEBYSHEV = 1, PREC_NONE = -1 };
// mfem::IterativeSolver::PrintLevel flags (Factory::CreatePrintLevel, jots_common.inl:57-91)
enum { PRINT_ERRORS = 1, PRINT_WARNINGS = 2, PRINT_ITERATIONS = 4, PRINT_SUMMARY = 8, PRINT_FIRST_AND_LAST = 16 };
bool quiet();   // JOTS_QUIET=1 (driver.cpp)

struct LinOp {
    virtual ~LinOp() {}
    virtual void mult(const double* x, double* y) = 0;
    virtual void assemble_diag(double* d) = 0;
    // y = A x and, when the operator can do it in the same pass, *d_dot += (x, A x) over the owned rows (*d_dot zeroed by the
    // caller, not yet all-reduced): returns whether it did
    virtual bool mult_dot(const double* x, double* y, double* d_dot) { (void)d_dot; mult(x, y); return false; }
    // > 0: mult_dot_side is available -- y = A x into a y that is ALREADY zero, *d_dot += (x, A x), and in the same launch the
    // conjugate-gradient side stream side.x += *side.alpha * side.d, side.zero = 0 (FormArgs::side_*)
    virtual int side_chunk() { return 0; }
    virtual void mult_dot_side(const double* x, double* y, double* d_dot, const CgSide& side) { (void)x; (void)y; (void)d_dot; (void)side; throw std::runtime_error("operator has no side stream"); }
};

// a FormSpec bound to a context
struct FormOp : LinOp {
    Ctx* c;
    FormSpec f;
    FormOp(Ctx* ctx, const FormSpec& spec) : c(ctx), f(spec) {}
    void mult(const double* x, double* y) override { c->apply_form(f, x, y); }
    bool mult_dot(const double* x, double* y, double* d_dot) override { return c->apply_form(f, x, y, d_dot); }
    void assemble_diag(double* d) override { c->diag_form(f, d); }
    int side_chunk() override { return c->form_side_chunk(f); }
    void mult_dot_side(const double* x, double* y, double* d_dot, const CgSide& side) override { c->apply_form(f, x, y, d_dot, &side); }
};

struct Precond {
    virtual ~Precond() {}
    virtual void set_operator(LinOp* A) = 0;
    virtual void mult(const double* r, double* z) = 0;
    virtual const double* jacobi_dinv() const { return nullptr; }  // non-null: z = dinv * r (fusable)
};

struct JacobiPrec : Precond {
    Ctx* c;
    DArr<double> dinv;
    explicit JacobiPrec(Ctx* ctx) : c(ctx) {}
    void set_operator(LinOp* A) override;
    void mult(const double* r, double* z) override;
    const double* jacobi_dinv() const override { return dinv.p; }
};

// HypreSmoother::Chebyshev (type 16): poly_order 2 (= degree-1 polynomial in hypre: cheby_order = order - 1),
// fraction 0.3, 10 CG iterations for the eigenvalue estimate of D^-1/2 A D^-1/2 from a deterministic pseudo-random vector
struct ChebyshevPrec : Precond {
    Ctx* c;
    LinOp* A = nullptr;
    DArr<double> ds, t0, t1, t2, t3;
    double coef[3] = {0, 0, 0};
    double lmax = 0, lmin = 0;
    int poly_order = 2, order = 1, eig_iters = 10;   // order = degree of p = poly_order - 1
    double fraction = 0.3;
    explicit ChebyshevPrec(Ctx* ctx) : c(ctx) {}
    void set_operator(LinOp* A) override;
    void mult(const double* r, double* z) override;
private:
    void ahat(const double* v, double* out, double* tmp);
};

struct SolveInfo {
    bool converged = false;
    int iterations = 0;
    double final_norm = 0.0;
    long long applies = 0;
};

struct Krylov {
    Ctx* c;
    int kind;
    double rel_tol = 1e-10, abs_tol = 1e-16;
    int max_iter = 100, print = 0;
    std::unique_ptr<Precond> prec;
    LinOp* A = nullptr;
    SolveInfo info;
    std::vector<std::unique_ptr<DArr<double>>> work;
    static constexpr int M = 50;  // restart length of GMRES / FGMRES

    Krylov(Ctx* ctx, int kind_, int prec_kind);
    void set_operator(LinOp* op) { A = op; if (prec) prec->set_operator(op); }
    void mult(const double* b, double* x);   // iterative_mode = false: x0 = 0
    // CG with Jacobi / no preconditioner keeps its scalars on the device (no host round trip per iteration);
    // JOTS_CG_HOST_SCALARS=1 selects the host-sequenced loop (same results; kept for A/B timing)
    bool device_scalars = true;
    bool fused_dot = true;           // (d, A d) formed by the element kernel when it can (JOTS_CG_FUSED_DOT=0: separate pass)
    bool side_stream = true;         // x += alpha d and the zeroing of the next output carried by the element kernel when it can
                                     // (JOTS_CG_SIDE=0: in the vector kernels)
    ~Krylov();
private:
    struct CgDevice {
        DArr<double> S;              // vec.cu: k_cgd_* scalar block
        double* h = nullptr;         // pinned: NSLOT polling slots + the initial values
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    } cgd;
    DArr<const double*> d_vptr;      // basis-vector pointers of the one-block Gram-Schmidt sweep
    void cg_device_loop(double* x, double* r, double* z, double* d, const double* dinv, double nom, double den, double r0);
    double* w(int i);
    void cg(const double* b, double* x);
    void gmres(const double* b, double* x, bool flexible);
    void apply(const double* x, double* y) { A->mult(x, y); ++info.applies; }
};

// operator handed to Newton: residual, gradient, ProcessNewState hook
struct NewtonOp {
    virtual ~NewtonOp() {}
    virtual void new_state(const double* x) = 0;
    virtual void mult(const double* x, double* y) = 0;
    virtual LinOp* gradient(const double* x) = 0;
};

struct Newton {
    Ctx* c;
    double rel_tol = 1e-10, abs_tol = 1e-16;
    int max_iter = 100, print = 0;
    bool converged = false;
    int iterations = 0;
    double final_norm = 0.0;
    DArr<double> r, cvec;
    explicit Newton(Ctx* ctx) : c(ctx) {}
    // b may be null (zero right-hand side); x is the initial guess and the result
    void mult(NewtonOp& op, Krylov& lin, const double* b, double* x);
};

void lanczos_extreme_eigs(const std::vector<double>& alphas, const std::vector<double>& betas, double& lmax, double& lmin);

}  // namespace jots
