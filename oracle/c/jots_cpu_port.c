/*
 * CPU restatement (C + OpenMP) of the reference's nonlinear unsteady conduction step -- TEST / BASELINE
 * INFRASTRUCTURE ONLY (oracle/__init__.py explains the rules; the product never links or runs this).
 *
 * It follows the reference's algorithm class on the host, which is what `mpirun -np <cores> jots` does
 * inside MFEM/hypre (not buildable here):
 *   - ReducedSystemOperatorR::Mult        /root/reference/src/nl_conduction_operator.cpp:70-79
 *   - ReducedSystemOperatorR::GetGradient /root/reference/src/nl_conduction_operator.cpp:41-67,81-90
 *     = dense element Jacobians (JOTSNonlinearDiffusionIntegrator::AssembleElementGrad,
 *       /root/reference/src/jots_nlfis.cpp:24-40) -> global CSR, essential rows/cols eliminated,
 *       re-assembled at EVERY Newton iteration
 *   - mfem::NewtonSolver + CGSolver + HypreSmoother(Jacobi) on that matrix (SpMV based), MFEM stopping rules
 *   - BackwardEulerSolver::Step
 * Restricted to the BASELINE workload: structured brick, H1 order p hexes, uniform rho and C, uniform or
 * polynomial k, heat flux on x = 0, isothermal x = L, other faces adiabatic; p+2 Gauss points per
 * direction, nodal-interpolated k_h, k'_h.  Parity of this port is pinned against the NumPy oracle
 * (tests/test_oracle_c_port.py), which itself is pinned on the reference's analytic regressions.
 */
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 5
#define MAXQ 6
#define MAXP 8

static int D, Q, ND, NQ;
static double Bt[MAXQ][MAXD], Gt[MAXQ][MAXD], Wt[MAXQ], nodes[MAXD];

static void legendre(int n, long double x, long double* P, long double* dP) {
    long double p0 = 1, p1 = x;
    if (n == 0) { *P = 1; *dP = 0; return; }
    for (int k = 2; k <= n; ++k) { long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = pk; }
    *P = p1; *dP = n * (x * p1 - p0) / (x * x - 1);
}
static void gauss01(int n, double* x, double* w) {
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L)), P, dP;
        for (int it = 0; it < 100; ++it) { legendre(n, z, &P, &dP); long double dz = P / dP; z -= dz; if (fabsl(dz) < 1e-19L) break; }
        legendre(n, z, &P, &dP);
        long double wt = 2 / ((1 - z * z) * dP * dP);
        x[n - 1 - i] = (double)(0.5L * (1 + z)); x[i] = (double)(0.5L * (1 - z)); w[n - 1 - i] = w[i] = (double)(0.5L * wt);
    }
    if (n % 2) x[n / 2] = 0.5;
}
static void lobatto01(int n, double* x) {
    const long double pi = 3.141592653589793238462643383279502884L;
    const int m = n - 1;
    x[0] = 0; x[n - 1] = 1;
    for (int i = 1; i < n - 1; ++i) {
        if (2 * i == m) { x[i] = 0.5; continue; }
        long double z = -cosl(pi * i / m);
        for (int it = 0; it < 100; ++it) {
            long double P, dP; legendre(m, z, &P, &dP);
            long double d2P = (2 * z * dP - m * (m + 1) * P) / (1 - z * z), dz = dP / d2P;
            z -= dz; if (fabsl(dz) < 1e-19L) break;
        }
        x[i] = (double)(0.5L * (1 + z));
    }
    for (int i = 0; i < n / 2; ++i) { double a = 0.5 * (x[i] + (1 - x[n - 1 - i])); x[i] = a; x[n - 1 - i] = 1 - a; }
}
static void tables(int p) {
    double qx[MAXQ];
    D = p + 1; Q = p + 2; ND = D * D * D; NQ = Q * Q * Q;
    lobatto01(D, nodes); gauss01(Q, qx, Wt);
    for (int q = 0; q < Q; ++q)
        for (int d = 0; d < D; ++d) {
            long double den = 1, b = 1, g = 0;
            for (int m = 0; m < D; ++m) if (m != d) den *= (long double)nodes[d] - nodes[m];
            for (int m = 0; m < D; ++m) {
                if (m == d) continue;
                b *= (long double)qx[q] - nodes[m];
                long double t = 1;
                for (int r = 0; r < D; ++r) if (r != d && r != m) t *= (long double)qx[q] - nodes[r];
                g += t;
            }
            Bt[q][d] = (double)(b / den); Gt[q][d] = (double)(g / den);
        }
}

/* PolynomialProperty arithmetic (material_property.cpp:36-64) */
static int kn; static double kc[MAXP];
static double kval(double u) {
    if (kn == 1) return kc[0];
    double z1 = 0; for (int i = 0; i < kn; ++i) { double z2 = 1; for (int j = 0; j < kn - i - 1; ++j) z2 *= u; z1 += z2 * kc[i]; } return z1;
}
static double kder(double u) {
    if (kn == 1) return 0;
    double z1 = 0; for (int i = 0; i < kn - 1; ++i) { double z2 = kn - 1 - i; for (int j = 0; j < kn - i - 2; ++j) z2 *= u; z1 += z2 * kc[i]; } return z1;
}

typedef struct { long n; long* ptr; int* col; double* val; } CSR;

static long NX, NY, NZ, ndof, nelem;
static int nx, ny, nz, P;
static double hx, hy, hz;
static inline long gdof(int ex, int ey, int ez, int a, int b, int c) { return (long)(ex * P + a) + NX * ((long)(ey * P + b) + NY * (long)(ez * P + c)); }

static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

static void build_pattern(CSR* A) {
    A->n = ndof; A->ptr = (long*)malloc((ndof + 1) * sizeof(long));
    long* cnt = (long*)calloc(ndof + 1, sizeof(long));
    /* row i couples with the DOFs of every element containing it: a lattice box */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < ndof; ++i) {
        int ix = i % NX, iy = (i / NX) % NY, iz = i / (NX * NY);
        int lo[3], hi[3], idx[3] = {ix, iy, iz}, ne[3] = {nx, ny, nz};
        for (int a = 0; a < 3; ++a) {
            int e0 = idx[a] / P, e1 = e0;
            if (idx[a] % P == 0) { e0 = idx[a] / P - 1; e1 = idx[a] / P; }
            if (e0 < 0) e0 = 0; if (e1 > ne[a] - 1) e1 = ne[a] - 1;
            lo[a] = e0 * P; hi[a] = (e1 + 1) * P;
        }
        cnt[i + 1] = (long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
    }
    A->ptr[0] = 0;
    for (long i = 0; i < ndof; ++i) A->ptr[i + 1] = A->ptr[i] + cnt[i + 1];
    A->col = (int*)malloc(A->ptr[ndof] * sizeof(int));
    A->val = (double*)malloc(A->ptr[ndof] * sizeof(double));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < ndof; ++i) {
        int ix = i % NX, iy = (i / NX) % NY, iz = i / (NX * NY);
        int lo[3], hi[3], idx[3] = {ix, iy, iz}, ne[3] = {nx, ny, nz};
        for (int a = 0; a < 3; ++a) {
            int e0 = idx[a] / P, e1 = e0;
            if (idx[a] % P == 0) { e0 = idx[a] / P - 1; e1 = idx[a] / P; }
            if (e0 < 0) e0 = 0; if (e1 > ne[a] - 1) e1 = ne[a] - 1;
            lo[a] = e0 * P; hi[a] = (e1 + 1) * P;
        }
        long k = A->ptr[i];
        for (int c = lo[2]; c <= hi[2]; ++c) for (int b = lo[1]; b <= hi[1]; ++b) for (int a = lo[0]; a <= hi[0]; ++a)
            A->col[k++] = (int)(a + NX * (b + NY * (long)c));
    }
    free(cnt);
    (void)cmp_int;
}
static inline long find_col(const CSR* A, long row, int col) {
    long lo = A->ptr[row], hi = A->ptr[row + 1] - 1;
    while (lo <= hi) { long mid = (lo + hi) >> 1; int c = A->col[mid]; if (c == col) return mid; if (c < col) lo = mid + 1; else hi = mid - 1; }
    return -1;
}
static void spmv(const CSR* A, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < A->n; ++i) {
        double s = 0; for (long k = A->ptr[i]; k < A->ptr[i + 1]; ++k) s += A->val[k] * x[A->col[k]]; y[i] = s;
    }
}
static double dot(long n, const double* a, const double* b) {
    double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* element quadrature data at state z: alpha_q, alpha'_q (nodal k(z), k'(z) interpolated), grad z */
static void elem_qdata(const double* zl, double inv_rc, double* aq, double* daq, double* gz /*[NQ][3]*/) {
    double kn_[MAXD * MAXD * MAXD], dk_[MAXD * MAXD * MAXD];
    for (int l = 0; l < ND; ++l) { kn_[l] = kval(zl[l]); dk_[l] = kder(zl[l]); }
    for (int q = 0; q < NQ; ++q) {
        int q0 = q % Q, q1 = (q / Q) % Q, q2 = q / (Q * Q);
        double a = 0, da = 0, g0 = 0, g1 = 0, g2 = 0;
        for (int l = 0; l < ND; ++l) {
            int i0 = l % D, i1 = (l / D) % D, i2 = l / (D * D);
            double b0 = Bt[q0][i0], b1 = Bt[q1][i1], b2 = Bt[q2][i2], phi = b0 * b1 * b2;
            a += phi * kn_[l]; da += phi * dk_[l];
            g0 += Gt[q0][i0] * b1 * b2 * zl[l]; g1 += b0 * Gt[q1][i1] * b2 * zl[l]; g2 += b0 * b1 * Gt[q2][i2] * zl[l];
        }
        aq[q] = a * inv_rc; daq[q] = da * inv_rc;
        gz[3 * q] = g0 / hx; gz[3 * q + 1] = g1 / hy; gz[3 * q + 2] = g2 / hz;
    }
}

int main(int argc, char** argv) {
    if (argc < 20) {
        fprintf(stderr, "usage: %s nx ny nz lx ly lz order dt steps rho C T0 flux_x0 T_x1 rel abs maxit nrel nmaxit nthreads kn k0 [k1..] [--out file]\n", argv[0]);
        return 2;
    }
    int ai = 1;
    nx = atoi(argv[ai++]); ny = atoi(argv[ai++]); nz = atoi(argv[ai++]);
    double lx = atof(argv[ai++]), ly = atof(argv[ai++]), lz = atof(argv[ai++]);
    P = atoi(argv[ai++]);
    double dt = atof(argv[ai++]); int steps = atoi(argv[ai++]);
    double rho = atof(argv[ai++]), Cc = atof(argv[ai++]), T0 = atof(argv[ai++]), flux = atof(argv[ai++]), Tiso = atof(argv[ai++]);
    double rel = atof(argv[ai++]), abs_ = atof(argv[ai++]); int maxit = atoi(argv[ai++]);
    double nrel = atof(argv[ai++]); int nmaxit = atoi(argv[ai++]); int nthreads = atoi(argv[ai++]);
    kn = atoi(argv[ai++]);
    for (int i = 0; i < kn; ++i) kc[i] = atof(argv[ai++]);
    const char* outfile = NULL;
    if (ai + 1 < argc && !strcmp(argv[ai], "--out")) outfile = argv[ai + 1];
    if (nthreads > 0) omp_set_num_threads(nthreads);
    tables(P);
    NX = nx * P + 1; NY = ny * P + 1; NZ = nz * P + 1; ndof = NX * NY * NZ; nelem = (long)nx * ny * nz;
    hx = lx / nx; hy = ly / ny; hz = lz / nz;
    const double detJ = hx * hy * hz, inv_rc = 1.0 / (rho * Cc), nabs = 1e-16;

    double* u = (double*)malloc(ndof * sizeof(double));
    double* k = (double*)calloc(ndof, sizeof(double));
    double* z = (double*)malloc(ndof * sizeof(double));
    double* r = (double*)malloc(ndof * sizeof(double));
    double* c = (double*)malloc(ndof * sizeof(double));
    double* bvec = (double*)calloc(ndof, sizeof(double));
    double* cr = (double*)malloc(ndof * sizeof(double)); double* cz = (double*)malloc(ndof * sizeof(double));
    double* cd = (double*)malloc(ndof * sizeof(double)); double* dinv = (double*)malloc(ndof * sizeof(double));
    unsigned char* ess = (unsigned char*)calloc(ndof, 1);
    for (long i = 0; i < ndof; ++i) u[i] = T0;
    /* essential: x = L plane; Neumann vector b = int g/(rho C) phi on x = 0 with (p+1)/2+1 points */
    for (long c2 = 0; c2 < NZ; ++c2) for (long b2 = 0; b2 < NY; ++b2) ess[(NX - 1) + NX * (b2 + NY * c2)] = 1;
    {
        int nqf = (P + 1) / 2 + 1; double xf[MAXQ], wf[MAXQ], Bf[MAXQ][MAXD];
        gauss01(nqf, xf, wf);
        for (int q = 0; q < nqf; ++q) for (int d = 0; d < D; ++d) {
            long double den = 1, b = 1;
            for (int m = 0; m < D; ++m) if (m != d) { den *= (long double)nodes[d] - nodes[m]; b *= (long double)xf[q] - nodes[m]; }
            Bf[q][d] = (double)(b / den);
        }
        for (int ez = 0; ez < nz; ++ez) for (int ey = 0; ey < ny; ++ey)
            for (int cq = 0; cq < nqf; ++cq) for (int bq = 0; bq < nqf; ++bq)
                for (int cc = 0; cc < D; ++cc) for (int bb = 0; bb < D; ++bb)
                    bvec[gdof(0, ey, ez, 0, bb, cc)] += wf[bq] * wf[cq] * hy * hz * (flux * inv_rc) * Bf[bq][bb] * Bf[cq][cc];
    }
    CSR J; build_pattern(&J);
    /* unit mass element matrix (constant on a brick) */
    double* Me = (double*)calloc((size_t)ND * ND, sizeof(double));
    for (int q = 0; q < NQ; ++q) {
        int q0 = q % Q, q1 = (q / Q) % Q, q2 = q / (Q * Q); double w = Wt[q0] * Wt[q1] * Wt[q2] * detJ;
        for (int i = 0; i < ND; ++i) { double pi = Bt[q0][i % D] * Bt[q1][(i / D) % D] * Bt[q2][i / (D * D)];
            for (int j = 0; j < ND; ++j) Me[i * ND + j] += w * pi * Bt[q0][j % D] * Bt[q1][(j / D) % D] * Bt[q2][j / (D * D)]; }
    }
    long krylov = 0, newton_total = 0, resid_evals = 0;
    double t0 = omp_get_wtime();
    for (int step = 0; step < steps; ++step) {
        for (long i = 0; i < ndof; ++i) if (ess[i]) u[i] = Tiso;              /* UpdateAndApplyBCs */
        memset(k, 0, ndof * sizeof(double));                                   /* iterative_mode = false */
        double norm0 = 0, norm = 0; int it = 0;
        for (;; ++it) {
            /* ProcessNewState + residual R(k) = M k + kappa(z) - b, essential rows zeroed */
#pragma omp parallel for schedule(static)
            for (long i = 0; i < ndof; ++i) { z[i] = u[i] + dt * k[i]; r[i] = -bvec[i]; }
#pragma omp parallel
            {
                double zl[MAXD * MAXD * MAXD], kl[MAXD * MAXD * MAXD], el[MAXD * MAXD * MAXD];
                double* aq = (double*)malloc(NQ * sizeof(double)); double* daq = (double*)malloc(NQ * sizeof(double));
                double* gz = (double*)malloc(3 * NQ * sizeof(double));
#pragma omp for schedule(static)
                for (long e = 0; e < nelem; ++e) {
                    int ex = e % nx, ey = (e / nx) % ny, ez = e / ((long)nx * ny);
                    for (int l = 0; l < ND; ++l) { long g = gdof(ex, ey, ez, l % D, (l / D) % D, l / (D * D)); zl[l] = z[g]; kl[l] = ess[g] ? 0.0 : k[g]; el[l] = 0; }
                    elem_qdata(zl, inv_rc, aq, daq, gz);
                    for (int i = 0; i < ND; ++i) { double s = 0; for (int j = 0; j < ND; ++j) s += Me[i * ND + j] * kl[j]; el[i] = s; }
                    for (int q = 0; q < NQ; ++q) {
                        int q0 = q % Q, q1 = (q / Q) % Q, q2 = q / (Q * Q); double w = Wt[q0] * Wt[q1] * Wt[q2] * detJ * aq[q];
                        for (int i = 0; i < ND; ++i) {
                            int i0 = i % D, i1 = (i / D) % D, i2 = i / (D * D);
                            double d0 = Gt[q0][i0] * Bt[q1][i1] * Bt[q2][i2] / hx, d1 = Bt[q0][i0] * Gt[q1][i1] * Bt[q2][i2] / hy, d2 = Bt[q0][i0] * Bt[q1][i1] * Gt[q2][i2] / hz;
                            el[i] += w * (gz[3 * q] * d0 + gz[3 * q + 1] * d1 + gz[3 * q + 2] * d2);
                        }
                    }
                    for (int l = 0; l < ND; ++l) { long g = gdof(ex, ey, ez, l % D, (l / D) % D, l / (D * D));
#pragma omp atomic
                        r[g] += el[l]; }
                }
                free(aq); free(daq); free(gz);
            }
            for (long i = 0; i < ndof; ++i) if (ess[i]) r[i] = 0;
            ++resid_evals;
            norm = sqrt(dot(ndof, r, r));
            if (it == 0) norm0 = norm;
            double goal = fmax(nrel * norm0, nabs);
            if (norm <= goal || it >= nmaxit) break;
            /* GetGradient: dense element Jacobians -> CSR (re-assembled every Newton iteration) */
            memset(J.val, 0, J.ptr[ndof] * sizeof(double));
#pragma omp parallel
            {
                double zl[MAXD * MAXD * MAXD];
                double* aq = (double*)malloc(NQ * sizeof(double)); double* daq = (double*)malloc(NQ * sizeof(double));
                double* gz = (double*)malloc(3 * NQ * sizeof(double)); double* Je = (double*)malloc((size_t)ND * ND * sizeof(double));
                double* dphi = (double*)malloc((size_t)3 * ND * sizeof(double)); double* phi = (double*)malloc(ND * sizeof(double));
#pragma omp for schedule(static)
                for (long e = 0; e < nelem; ++e) {
                    int ex = e % nx, ey = (e / nx) % ny, ez = e / ((long)nx * ny);
                    long gl[MAXD * MAXD * MAXD];
                    for (int l = 0; l < ND; ++l) { gl[l] = gdof(ex, ey, ez, l % D, (l / D) % D, l / (D * D)); zl[l] = z[gl[l]]; }
                    elem_qdata(zl, inv_rc, aq, daq, gz);
                    for (int i = 0; i < ND * ND; ++i) Je[i] = Me[i];
                    for (int q = 0; q < NQ; ++q) {
                        int q0 = q % Q, q1 = (q / Q) % Q, q2 = q / (Q * Q); double w = Wt[q0] * Wt[q1] * Wt[q2] * detJ * dt;
                        for (int i = 0; i < ND; ++i) {
                            int i0 = i % D, i1 = (i / D) % D, i2 = i / (D * D);
                            phi[i] = Bt[q0][i0] * Bt[q1][i1] * Bt[q2][i2];
                            dphi[3 * i] = Gt[q0][i0] * Bt[q1][i1] * Bt[q2][i2] / hx; dphi[3 * i + 1] = Bt[q0][i0] * Gt[q1][i1] * Bt[q2][i2] / hy; dphi[3 * i + 2] = Bt[q0][i0] * Bt[q1][i1] * Gt[q2][i2] / hz;
                        }
                        for (int i = 0; i < ND; ++i) {
                            double gzi = gz[3 * q] * dphi[3 * i] + gz[3 * q + 1] * dphi[3 * i + 1] + gz[3 * q + 2] * dphi[3 * i + 2];
                            double wa = w * aq[q], wd = w * daq[q] * gzi;
                            for (int m = 0; m < ND; ++m)
                                Je[i * ND + m] += wa * (dphi[3 * i] * dphi[3 * m] + dphi[3 * i + 1] * dphi[3 * m + 1] + dphi[3 * i + 2] * dphi[3 * m + 2]) + wd * phi[m];
                        }
                    }
                    for (int i = 0; i < ND; ++i) {
                        if (ess[gl[i]]) continue;
                        for (int m = 0; m < ND; ++m) {
                            if (ess[gl[m]]) continue;
                            long pos = find_col(&J, gl[i], (int)gl[m]);
#pragma omp atomic
                            J.val[pos] += Je[i * ND + m];
                        }
                    }
                }
                free(aq); free(daq); free(gz); free(Je); free(dphi); free(phi);
            }
#pragma omp parallel for schedule(static)
            for (long i = 0; i < ndof; ++i) {
                if (ess[i]) J.val[find_col(&J, i, (int)i)] = 1.0;       /* EliminateRowsCols: diag one */
                dinv[i] = 1.0 / J.val[find_col(&J, i, (int)i)];          /* HypreSmoother::Jacobi */
            }
            /* mfem::CGSolver::Mult, x0 = 0, monitors (B r, r) */
            memset(c, 0, ndof * sizeof(double)); memcpy(cr, r, ndof * sizeof(double));
#pragma omp parallel for schedule(static)
            for (long i = 0; i < ndof; ++i) { cz[i] = dinv[i] * cr[i]; cd[i] = cz[i]; }
            double nom = dot(ndof, cd, cr), r0 = fmax(nom * rel * rel, abs_ * abs_);
            if (nom > r0) {
                spmv(&J, cd, cz); double den = dot(ndof, cz, cd);
                for (int ci = 1;; ) {
                    double alpha = nom / den, betanom = 0;
#pragma omp parallel for reduction(+ : betanom) schedule(static)
                    for (long i = 0; i < ndof; ++i) { c[i] += alpha * cd[i]; cr[i] -= alpha * cz[i]; cz[i] = dinv[i] * cr[i]; betanom += cr[i] * cz[i]; }
                    ++krylov;
                    if (betanom <= r0 || ++ci > maxit) break;
                    double beta = betanom / nom;
#pragma omp parallel for schedule(static)
                    for (long i = 0; i < ndof; ++i) cd[i] = cz[i] + beta * cd[i];
                    spmv(&J, cd, cz); den = dot(ndof, cd, cz); nom = betanom;
                }
            }
#pragma omp parallel for schedule(static)
            for (long i = 0; i < ndof; ++i) k[i] -= c[i];
            ++newton_total;
        }
#pragma omp parallel for schedule(static)
        for (long i = 0; i < ndof; ++i) u[i] += dt * k[i];                    /* BackwardEulerSolver::Step */
    }
    double elapsed = omp_get_wtime() - t0;
    double un = sqrt(dot(ndof, u, u));
    if (outfile) { FILE* f = fopen(outfile, "wb"); fwrite(u, sizeof(double), ndof, f); fclose(f); }
    printf("{\"ndof\": %ld, \"time\": %.6f, \"steps\": %d, \"krylov_iters\": %ld, \"newton_iters\": %ld, \"residual_evals\": %ld, \"threads\": %d, \"u_norm\": %.17g}\n",
           ndof, elapsed, steps, krylov, newton_total, resid_evals, omp_get_max_threads(), un);
    return 0;
}
