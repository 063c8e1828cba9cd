// Element "forms" of the conduction solve evaluated matrix-free on the device.
//
// Every domain integral the reference assembles into a sparse matrix or element vector
// (a3, a6-a10, a13, a15) has the quadrature-point shape
//
//      y_i = sum_q w_q |J| [ F_q . grad phi_i  +  S_q phi_i ]
//
// with, for an input field x (value x_q, physical gradient gx) and a frozen state z (gradient gz):
//
//   OP_LIN : F = a1 gx                     S = m x_q
//            MassIntegrator / DiffusionIntegrator of linear_conduction_operator.cpp:59,63 and
//            nl_conduction_operator.cpp:189,200
//   OP_JAC : F = a1 gx + a2 x_q gz         S = m x_q + 2 nb (gz.gx) + ndb |gz|^2 x_q
//            JOTSNonlinearDiffusionIntegrator::AssembleElementGrad (jots_nlfis.cpp:24-40) and
//            JOTSNonlinearConvectionIntegrator::AssembleElementGrad (jots_nlfis.cpp:83-95)
//   OP_RES : F = a1 gz                     S = m x_q + nb |gz|^2
//            ::AssembleElementVector of the same integrators (jots_nlfis.cpp:16-22, 75-81) plus the
//            mass term of ReducedSystemOperatorR::Mult (nl_conduction_operator.cpp:70-79)
//
// The coefficients are *nodal fields interpolated to the quadrature points* (PolynomialProperty ->
// GridFunctionCoefficient, material_property.cpp:31-49), ratios are ratios of interpolants
// (nl_conduction_operator.cpp:92-163):
//
//   CF_CONST     : m = cm*m0, a1 = ck*a0
//   CF_LINFIELDS : m = cm*rho_h*C_h, a1 = ck*k_h                     (linearized operator)
//   CF_K         : m = cm*m0, a1 = ck*k_h*inv_rc, a2 = ck*k'_h*inv_rc   (rho, C uniform)
//   CF_FULL      : m = cm, a1 = ck*alpha, a2 = ck*alpha', nb = cb*(-beta), ndb = cb*(-beta')
#pragma once
#include <cstdint>

namespace jots {

enum { OP_LIN = 0, OP_JAC = 1, OP_RES = 2 };
enum { CF_CONST = 0, CF_LINFIELDS = 1, CF_K = 2, CF_FULL = 3,
       CF_FULLC = 4,     // CF_FULL with a uniform density (slab kernel only, chosen at launch): k, k', C, C', C''
       CF_KL = 5 };      // CF_K with k linear in T (two polynomial coefficients): k'_h is uniform, one interpolated field
                         // instead of two in the Jacobian apply (slab kernel only, chosen at launch)

constexpr int MAX_POLY = 8;

// Uniform (n == 1) or Polynomial (n >= 2, highest power first) property
struct Poly {
    int n;
    double c[MAX_POLY];
};
struct MatSet {
    Poly rho, C, k;
};

struct FormArgs {
    const int* gather;    // [nelem*nloc]; essential DOFs are stored as ~index (negative)
    const double* geom;   // [nelem*10]: invJ[a][i] (9, row-major: a = reference axis) + detJ
    int geom_stride;      // 10, or 0 when every element shares geom[0..9]
    int geom_diag;        // metric invJ invJ^T is diagonal on every element (axis-aligned bricks)
    const double* geomq;  // non-affine (general multilinear) meshes: [nelem][NQ][10] per quadrature point, else null
    const double* x;      // input vector
    const double* z;      // frozen state (coefficients + grad z); may be null for CF_CONST/OP_LIN
    const double* zc;     // state the nodal property polynomials are evaluated on when it differs from z (explicit
                          // stages: properties frozen at the last UpdateMatProps), else null; generic kernel only
    double* y;            // output, accumulated with atomics (caller zeroes it)
    int nelem;
    int mask_in;          // treat x as zero on essential DOFs (eliminated columns)
    int mask_out;         // skip essential rows
    int diag_one;         // with mask_out: write y = x on essential rows from the kernel (unit diagonal)
    double cm, ck, cb;
    double m0, a0;        // CF_CONST
    double inv_rc;        // CF_K
    MatSet mat;
    long long nvec;       // length of x / z / y
    // element patches of the unique-pencil kernel (apply_patch.cuh, patches.cpp): null when the mesh is not (mostly) a lattice
    const int* patches;
    int n_patches;
    // device flag: the kernel returns at once when *skip != 0 (iterations enqueued ahead of a converged Krylov solve)
    const double* skip;
    // two zero-initialised counters of the dynamic tile scheduler of apply_patch2.cuh (tile counter, exit counter)
    unsigned int* sched;
    // non-null: the kernel adds sum_elements x_e . (A_e x_e) = (x, A x) over the rows it writes to *dot (two-phase patch
    // kernel only; the conjugate-gradient loop then needs no separate pass for (d, A d))
    double* dot;
    // Side stream of the conjugate-gradient loop (two-phase patch kernel, dot variant of the decoupled pipeline only;
    // side_chunk == 0: none).  The kernel is FP64-issue bound and leaves HBM idle, so the two vector passes of an iteration
    // that are off its critical path ride along: the block that processes tile t also does, over entries
    // [t * side_chunk, (t + 1) * side_chunk) of the vectors,
    //     side_x += *side_alpha * side_d      (the solution update x += alpha_{i-1} d_{i-1}, deferred by one apply)
    //     side_zero = 0                       (the output buffer of the NEXT apply; this launch's y was zeroed by the previous one)
    // n_patches * side_chunk >= nvec, side_chunk even and <= 2 * (threads - 32); all three vectors 16-byte aligned.
    double* side_x;
    const double* side_d;
    const double* side_alpha;
    double* side_zero;
    int side_chunk;
};

// the side stream of FormArgs as the solvers hand it down (ctx.hpp: Ctx::apply_form)
struct CgSide {
    double* x;
    const double* d;
    const double* alpha;    // device scalar
    double* zero;
    int ess_zero;           // the input is known to vanish on the essential rows: their share of (x, A x) needs no kernel
};

// 1-D tables; DZ = QZ = 1 turns the hex kernel into the quad kernel
template <int D, int Q, int DZ, int QZ>
struct Tables {
    double B[Q * D];       // B[q*D+i] = l_i(x_q)
    double G[Q * D];       // G[q*D+i] = l_i'(x_q)
    double Gq[Q * Q];      // collocated derivative at the quadrature points
    double W[Q];
    double Bz[QZ * DZ];
    double Gz[QZ * DZ];
    double Gqz[QZ * QZ];
    double Wz[QZ];
};

}  // namespace jots
