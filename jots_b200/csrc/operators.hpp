// The three conduction operators of the reference behind their own interface
// (mfem::TimeDependentOperator::Mult / ImplicitSolve + JOTSIterator::Iterate /
// ProcessMatPropUpdate / UpdateNeumann), running matrix-free on the device:
//   LinearConductionOperator      /root/reference/src/linear_conduction_operator.cpp:15-194
//   NonlinearConductionOperator   /root/reference/src/nl_conduction_operator.cpp:27-323
//   SteadyConductionOperator      /root/reference/src/steady_conduction_operator.cpp:3-38
// All vector arguments are device pointers of length ctx->n.
#pragma once
#include <memory>
#include <string>

#include "solvers.hpp"

namespace jots {

enum { MP_DENSITY = 0, MP_SPECIFIC_HEAT = 1, MP_THERMAL_CONDUCTIVITY = 2 };  // option_structure.hpp:18-23
enum { TS_EULER_IMPLICIT = 0, TS_EULER_EXPLICIT = 1, TS_RK4 = 2 };

struct SolverSpec {
    int solver = SOLVER_FGMRES, prec = PREC_CHEBYSHEV;
    double rel_tol = 1e-10, abs_tol = 1e-16;
    int max_iter = 100, print = 0;
};
struct NewtonSpec {
    double rel_tol = 1e-10, abs_tol = 1e-16;
    int max_iter = 0, print = 0;
};

// JOTSIterator (/root/reference/src/jots_iterator.{hpp,cpp})
class JOTSIterator {
public:
    JOTSIterator(Ctx* c, const SolverSpec& ls);
    virtual ~JOTSIterator() {}
    virtual void Iterate(double* u) = 0;
    virtual void ProcessMatPropUpdate(int mp) {}
    virtual void UpdateNeumann();
    // material state: the vector the nodal polynomials were last evaluated on
    // (PolynomialProperty::UpdateAllCoeffs called from JOTSDriver::UpdateMatProps)
    virtual void SetMatState(const double* u) { c->set_mat_state(u); }
    virtual void set_print_level(int linear_mask, int newton_mask) { ls.print = linear_mask; }
    Ctx* c;
    SolverSpec ls;
    DArr<double> b_vec;     // Neumann linear form
    double neumann_div = 1.0;
    double time = 0.0, dt = 0.0;
    int time_scheme = TS_EULER_IMPLICIT;
protected:
    std::unique_ptr<Krylov> make_solver();
    void ode_step(double* u);      // BackwardEuler / ForwardEuler / RK4 ::Step
    virtual void Mult(const double* u, double* k) { throw std::runtime_error("explicit Mult not available"); }
    virtual void ImplicitSolve(double dt_, const double* u, double* k) { throw std::runtime_error("ImplicitSolve not available"); }
    DArr<double> kvec, rk[4], utmp;
};

class LinearConductionOperator : public JOTSIterator {
public:
    LinearConductionOperator(Ctx* c, const SolverSpec& ls, double dt, int scheme);
    void Iterate(double* u) override;
    void ProcessMatPropUpdate(int mp) override;
    void Mult(const double* u, double* k) override;
    void ImplicitSolve(double dt_, const double* u, double* k) override;
    void CalculateRHS(const double* u, double* rhs);
    FormSpec M_spec(), K_spec(), A_spec();
    void set_print_level(int linear_mask, int) override { ls.print = linear_mask; expl_solver->print = linear_mask; impl_solver->print = linear_mask; }
private:
    void ReassembleM();
    void ReassembleK();
    void ReassembleA();
    int coef_mode() const;
    std::unique_ptr<Krylov> expl_solver, impl_solver;
    std::unique_ptr<FormOp> M_op, A_op;
    DArr<double> rhs;
};

class NonlinearConductionOperator : public JOTSIterator, public NewtonOp {
public:
    NonlinearConductionOperator(Ctx* c, const SolverSpec& ls, const NewtonSpec& nw, double dt, int scheme);
    void Iterate(double* u) override;
    void Mult(const double* u, double* k) override;
    void ImplicitSolve(double dt_, const double* u, double* k) override;
    // ReducedSystemOperatorR as a NewtonOp (iterating on k)
    void new_state(const double* k) override;
    void mult(const double* k, double* y) override;
    LinOp* gradient(const double* k) override;
    // building blocks exposed for parity tests
    void A_mult(const double* z, double* y);          // ReducedSystemOperatorA::Mult
    FormSpec jacobian_spec(const double* z);          // R.GetGradient at state z
    void set_un(const double* un) { u_n = un; }       // JOTS_k_Operator::SetParameters
    void set_print_level(int linear_mask, int newton_mask) override { ls.print = linear_mask; lin_solver->print = linear_mask; newton.print = newton_mask; }
    bool all_const = false, has_BN = false;
private:
    NewtonSpec nw;
    std::unique_ptr<Krylov> lin_solver;
    Newton newton;
    std::unique_ptr<FormOp> J_op, M_op;
    const double* u_n = nullptr;
    DArr<double> z, tmp, tmp2;
    double inv_rc = 1.0;
};

class SteadyConductionOperator : public JOTSIterator, public NewtonOp {
public:
    SteadyConductionOperator(Ctx* c, const SolverSpec& ls, const NewtonSpec& nw);
    void Iterate(double* u) override;
    void new_state(const double* u) override {}
    void mult(const double* u, double* y) override;
    LinOp* gradient(const double* u) override;
    void set_print_level(int linear_mask, int newton_mask) override { ls.print = linear_mask; lin_solver->print = linear_mask; newton.print = newton_mask; }
    bool zero_rhs_on_essential = false;  // reference quirk B.3: b_vec is NOT zeroed by default
private:
    NewtonSpec nw;
    std::unique_ptr<Krylov> lin_solver;
    Newton newton;
    std::unique_ptr<FormOp> J_op;
    DArr<double> rhs;
};

}  // namespace jots
