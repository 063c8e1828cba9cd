#include "solvers.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "halo.hpp"

namespace jots {

// ------------------------------------------------------------------------------------------------
// preconditioners
// ------------------------------------------------------------------------------------------------
void JacobiPrec::set_operator(LinOp* A) {
    dinv.alloc(c->n);
    A->assemble_diag(dinv.p);
    v_recip(c->stream, c->n, dinv.p, dinv.p);
    ++c->stats.kernel_launches;
}
void JacobiPrec::mult(const double* r, double* z) {
    v_mul(c->stream, c->n, dinv.p, r, z);
    ++c->stats.kernel_launches;
}

void lanczos_extreme_eigs(const std::vector<double>& alphas, const std::vector<double>& betas, double& lmax, double& lmin) {
    const int m = (int)alphas.size();
    if (m == 0) { lmax = lmin = 1.0; return; }
    std::vector<double> T((size_t)m * m, 0.0);
    for (int j = 0; j < m; ++j) {
        T[j * m + j] = 1.0 / alphas[j] + (j > 0 ? betas[j - 1] / alphas[j - 1] : 0.0);
        if (j + 1 < m) T[j * m + j + 1] = T[(j + 1) * m + j] = std::sqrt(betas[j]) / alphas[j];
    }
    // cyclic Jacobi eigenvalue iteration on the small symmetric tridiagonal
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < m; ++i) for (int j = i + 1; j < m; ++j) off += T[i * m + j] * T[i * m + j];
        if (off < 1e-300) break;
        for (int p = 0; p < m; ++p)
            for (int q = p + 1; q < m; ++q) {
                const double apq = T[p * m + q];
                if (apq == 0.0) continue;
                const double theta = (T[q * m + q] - T[p * m + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < m; ++k) {
                    const double akp = T[k * m + p], akq = T[k * m + q];
                    T[k * m + p] = cs * akp - sn * akq; T[k * m + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < m; ++k) {
                    const double apk = T[p * m + k], aqk = T[q * m + k];
                    T[p * m + k] = cs * apk - sn * aqk; T[q * m + k] = sn * apk + cs * aqk;
                }
            }
    }
    lmax = lmin = T[0];
    for (int i = 1; i < m; ++i) { lmax = std::max(lmax, T[i * m + i]); lmin = std::min(lmin, T[i * m + i]); }
}

void ChebyshevPrec::ahat(const double* v, double* out, double* tmp) {
    // out = ds * A (ds * v)
    v_mul(c->stream, c->n, ds.p, v, tmp);
    A->mult(tmp, out);
    v_mul(c->stream, c->n, ds.p, out, out);
    c->stats.kernel_launches += 2;
}

void ChebyshevPrec::set_operator(LinOp* op) {
    A = op;
    const int64_t n = c->n;
    ds.alloc(n); t0.alloc(n); t1.alloc(n); t2.alloc(n); t3.alloc(n);
    A->assemble_diag(ds.p);
    v_rsqrt(c->stream, n, ds.p, ds.p);
    // CG-Lanczos on the scaled operator
    double* r = t0.p; double* p = t1.p; double* Ap = t2.p; double* tmp = t3.p;
    v_hash_unit(c->stream, n, c->halo ? halo_global_index(*c) : nullptr, 0, r);
    c->stats.kernel_launches += 2;
    c->copy(r, p);
    double rr = c->dot(r, r);
    std::vector<double> alphas, betas;
    for (int it = 0; it < eig_iters; ++it) {
        ahat(p, Ap, tmp);
        const double pAp = c->dot(p, Ap);
        if (pAp <= 0 || rr == 0) break;
        const double alpha = rr / pAp;
        v_axpy(c->stream, n, -alpha, Ap, r);
        const double rr_new = c->dot(r, r);
        const double beta = rr_new / rr;
        alphas.push_back(alpha); betas.push_back(beta);
        v_waxpby(c->stream, n, 1.0, r, beta, p, p);
        c->stats.kernel_launches += 2;
        rr = rr_new;
    }
    lanczos_extreme_eigs(alphas, betas, lmax, lmin);
    const double ub = 1.1 * lmax, lb = lmin + fraction * (ub - lmin);
    const double theta = 0.5 * (ub + lb), delta = 0.5 * (ub - lb);
    // hypre_ParCSRRelax_Cheby_Setup: cheby_order = order - 1 ("we are using the order of p(A)"); MFEM's default
    // poly_order = 2 is therefore the degree-1 polynomial, one operator application per smoother call
    order = poly_order - 1;
    coef[0] = coef[1] = coef[2] = 0.0;
    if (order <= 0) { order = 0; coef[0] = 1.0 / theta; }
    else if (order == 1) {
        const double den = theta * theta + delta * theta;
        coef[0] = (delta + 2 * theta) / den;
        coef[1] = -1.0 / den;
    } else {
        order = 2;
        const double den = 2 * delta * theta * theta - delta * delta * theta - delta * delta * delta + 2 * theta * theta * theta;
        coef[0] = (4 * delta * theta - delta * delta + 6 * theta * theta) / den;
        coef[1] = -(2 * delta + 6 * theta) / den;
        coef[2] = 2.0 / den;
    }
}

void ChebyshevPrec::mult(const double* r, double* z) {
    // z = ds * (c0 + c1 Ahat + c2 Ahat^2) (ds * r), Horner
    const int64_t n = c->n;
    double* rs = t0.p; double* u = t1.p; double* au = t2.p; double* tmp = t3.p;
    v_mul2(c->stream, n, ds.p, r, coef[order], rs, u);
    ++c->stats.kernel_launches;
    for (int i = order - 1; i >= 0; --i) {
        // u = ds * A (ds * u) + c_i * rs
        v_mul(c->stream, n, ds.p, u, tmp);
        A->mult(tmp, au);
        v_mul_axpby(c->stream, n, ds.p, 1.0, au, coef[i], rs, u);
        c->stats.kernel_launches += 2;
    }
    v_mul(c->stream, n, ds.p, u, z);
    ++c->stats.kernel_launches;
}

// ------------------------------------------------------------------------------------------------
// Krylov
// ------------------------------------------------------------------------------------------------
Krylov::Krylov(Ctx* ctx, int kind_, int prec_kind) : c(ctx), kind(kind_) {
    if (prec_kind == PREC_JACOBI) prec.reset(new JacobiPrec(ctx));
    else if (prec_kind == PREC_CHEBYSHEV) prec.reset(new ChebyshevPrec(ctx));
    if (const char* e = std::getenv("JOTS_CG_HOST_SCALARS")) device_scalars = !(e[0] == '1');
    if (const char* e = std::getenv("JOTS_CG_FUSED_DOT")) fused_dot = !(e[0] == '0');
    if (const char* e = std::getenv("JOTS_CG_SIDE")) side_stream = !(e[0] == '0');
}

Krylov::~Krylov() {
    if (cgd.h) cudaFreeHost(cgd.h);
    for (auto& e : cgd.ev) if (e) cudaEventDestroy(e);
}

double* Krylov::w(int i) {
    while ((int)work.size() <= i) work.emplace_back(new DArr<double>());
    if (work[i]->n != (size_t)c->n) work[i]->alloc(c->n);
    return work[i]->p;
}

void Krylov::mult(const double* b, double* x) {
    NvtxRange range("jots:krylov");
    info = SolveInfo();
    if (!A) throw std::runtime_error("Krylov solver has no operator");
    if (kind == SOLVER_CG) cg(b, x);
    else gmres(b, x, kind == SOLVER_FGMRES);
    c->stats.krylov_iters += info.iterations;
    c->stats.last_linear_iters = info.iterations;
    c->stats.last_linear_norm = info.final_norm;
    c->stats.last_linear_converged = info.converged;
    // mfem::IterativeSolver print options (CGSolver / GMRESSolver / FGMRESSolver::Mult epilogue), rank 0 only
    if (print != 0 && c->rank() == 0 && !quiet()) {
        const char* name = kind == SOLVER_CG ? "PCG" : kind == SOLVER_GMRES ? "GMRES" : "FGMRES";
        if ((print & PRINT_FIRST_AND_LAST) && !(print & PRINT_ITERATIONS))
            std::printf("   Iteration : %3d  final norm = %g\n", info.iterations, info.final_norm);
        if ((print & PRINT_SUMMARY) || ((print & PRINT_ITERATIONS) && !info.converged))
            std::printf("%s: Number of iterations: %d\n", name, info.iterations);
        if ((print & PRINT_WARNINGS) && !info.converged) std::printf("%s: No convergence!\n", name);
    }
}

// mfem::CGSolver::Mult, preconditioned: monitors (B r, r)
void Krylov::cg(const double* b, double* x) {
    const int64_t n = c->n;
    cudaStream_t st = c->stream;
    double* r = w(0); double* z = w(1); double* d = w(2);
    c->zero(x);
    c->copy(b, r);
    const double* dinv = prec ? prec->jacobi_dinv() : nullptr;
    if (prec) prec->mult(r, z); else c->copy(r, z);
    c->copy(z, d);
    double nom = c->dot(d, r);
    if (nom < 0) return;
    const double r0 = std::max(nom * rel_tol * rel_tol, abs_tol * abs_tol);
    if (nom <= r0) { info.converged = true; info.final_norm = std::sqrt(nom); return; }
    apply(d, z);
    double den = c->dot(z, d);
    if (den == 0.0) { info.final_norm = std::sqrt(nom); return; }
    info.iterations = max_iter;
    double betanom = nom;
    const bool it_print = (print & PRINT_ITERATIONS) && c->rank() == 0 && !quiet();
    if (it_print || ((print & PRINT_FIRST_AND_LAST) && c->rank() == 0 && !quiet())) std::printf("   Iteration : %3d  (B r, r) = %g\n", 0, nom);
    // per-iteration output needs every (B r, r) on the host: the host-sequenced loop
    if ((!prec || dinv) && device_scalars && !(print & PRINT_ITERATIONS)) { cg_device_loop(x, r, z, d, dinv, nom, den, r0); return; }
    int i = 1;
    while (true) {
        const double alpha = nom / den;
        if (!prec || dinv) {
            // x += alpha d ; r -= alpha z ; z = dinv r ; betanom = (r, z)   -- one pass
            v_cg_update(st, n, c->n_owned, alpha, d, z, x, r, dinv, z, c->dot_scratch(), c->d_scal.p);
            ++c->stats.kernel_launches;
            c->fetch_scalars(1, &betanom);
        } else {
            v_axpy(st, n, alpha, d, x);
            v_axpy(st, n, -alpha, z, r);
            c->stats.kernel_launches += 2;
            prec->mult(r, z);
            betanom = c->dot(r, z);
        }
        if (it_print) std::printf("   Iteration : %3d  (B r, r) = %g\n", i, betanom);
        if (betanom < 0) { info.iterations = i; break; }
        if (betanom <= r0) { info.converged = true; info.iterations = i; break; }
        if (++i > max_iter) break;
        const double beta = betanom / nom;
        v_waxpby(st, n, 1.0, prec ? z : r, beta, d, d);
        ++c->stats.kernel_launches;
        apply(d, z);
        den = c->dot(d, z);
        if (den == 0.0) { info.iterations = i; break; }
        nom = betanom;
    }
    info.final_norm = std::sqrt(std::fabs(betanom));
}

// The iteration loop of cg() with every scalar in device memory (vec.cu: k_cgd_*): alpha = nom / den, the convergence test
// and beta = betanom / nom are evaluated by the kernels themselves, so the host only enqueues iterations -- in chunks, at most
// two chunks ahead of the last stop flag it has seen -- and never waits for a scalar.  Kernels enqueued past the stopping
// iteration return at once (the element kernel through FormArgs::skip), so x, the iteration count and the final norm are
// exactly those of the host-sequenced loop.
void Krylov::cg_device_loop(double* x, double* r, double* z, double* d, const double* dinv, double nom, double den, double r0) {
    const int64_t n = c->n;
    cudaStream_t st = c->stream;
    constexpr int CHUNK = 8, NSLOT = 4;
    if (!cgd.S.p) {
        cgd.S.alloc(8);
        JOTS_CUDA(cudaMallocHost((void**)&cgd.h, sizeof(double) * 8 * (NSLOT + 1)));
        for (int k = 0; k < NSLOT; ++k) JOTS_CUDA(cudaEventCreateWithFlags(&cgd.ev[k], cudaEventDisableTiming));
    }
    double* S = cgd.S.p;
    double* init = cgd.h + 8 * NSLOT;
    init[0] = nom; init[1] = 0.0; init[2] = den; init[3] = r0; init[4] = 0.0; init[5] = 0.0; init[6] = nom; init[7] = 0.0;
    JOTS_CUDA(cudaMemcpyAsync(S, init, 8 * sizeof(double), cudaMemcpyHostToDevice, st));
    c->apply_skip = S + 4;
    // Side-stream loop: the element kernel is FP64-issue bound and leaves HBM idle, so the two vector passes that are off the
    // critical path of an iteration ride along with it (FormArgs::side_*): x += alpha_i d_i is deferred to the apply of d_{i+1}
    // (d ping-pongs between two buffers so that d_i is still there), and that apply also zeroes the buffer the one after it
    // accumulates into (A d / z ping-pong as well).  The arithmetic of every entry is unchanged, so x, r, the iteration count
    // and the norms are those of the plain loop bit for bit; only the update that is still pending when the iteration stops is
    // applied after the loop.  Vector traffic per iteration: 8 passes instead of 12 (update 5, direction 3).
    const bool side = fused_dot && side_stream && A->side_chunk() > 0;
    double* D[2] = {d, side ? w(3) : d};
    double* Z[2] = {z, side ? w(4) : z};
    if (side) c->zero(Z[1]);
    // r vanishes on the essential rows (the right-hand sides of the reference's operators do: rhs[ess] = 0) => so does every
    // direction, exactly (unit diagonal rows: r -= alpha * d keeps a zero), and their share of (d, A d) needs no kernel
    const bool ess_zero = side && c->ess_rows_are_zero(r);
    int launched = 0, chunk = 0;
    bool stopped = false;
    while (launched < max_iter && !stopped) {
        if (chunk >= 2) {
            // the stop flag as it was after chunk - 2
            JOTS_CUDA(cudaEventSynchronize(cgd.ev[(chunk - 2) % NSLOT]));
            if (cgd.h[8 * ((chunk - 2) % NSLOT) + 4] != 0.0) { stopped = true; break; }
        }
        for (int k = 0; k < CHUNK && launched < max_iter; ++k) {
            const int i = ++launched;
            const int p = side ? ((i - 1) & 1) : 0;       // d_i is in D[p], A d_i (then z_i) in Z[p]
            v_cgd_update(st, n, c->n_owned, S, i, D[p], Z[p], side ? nullptr : x, r, dinv, Z[p], c->dot_scratch());
            c->allreduce_scalars(S + (i & 1), 1);
            const bool last = (i == max_iter);
            // the direction kernel zeroes the (d, A d) slot; an operator that cannot add to it during the apply leaves it to
            // the separate pass, which overwrites the slot
            v_cgd_direction(st, n, S, i, dinv ? Z[p] : r, D[p], D[side ? 1 - p : p], !last, fused_dot);
            c->stats.kernel_launches += 2;
            if (last) break;
            ++info.applies;
            if (side) {
                const CgSide sj{x, D[p], S + 7, Z[p], ess_zero ? 1 : 0};
                A->mult_dot_side(D[1 - p], Z[1 - p], S + 2, sj);
            } else if (!(fused_dot ? A->mult_dot(d, z, S + 2) : (A->mult(d, z), false))) {
                v_cgd_dot(st, c->n_owned, S, d, z, c->dot_scratch());
                ++c->stats.kernel_launches;
            }
            c->allreduce_scalars(S + 2, 1);
            ++c->stats.dots;
        }
        JOTS_CUDA(cudaMemcpyAsync(cgd.h + 8 * (chunk % NSLOT), S, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        JOTS_CUDA(cudaEventRecord(cgd.ev[chunk % NSLOT], st));
        ++chunk;
    }
    c->apply_skip = nullptr;
    JOTS_CUDA(cudaMemcpyAsync(cgd.h, S, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    JOTS_CUDA(cudaStreamSynchronize(st));
    const double* h = cgd.h;
    const int code = (int)h[4];
    double betanom;
    if (code != 0) { info.iterations = (int)h[5]; betanom = h[6]; info.converged = (code == 1); }
    else { info.iterations = max_iter; betanom = h[6]; }
    info.final_norm = std::sqrt(std::fabs(betanom));
    if (side && code != 3 && info.iterations >= 1) {
        // the update of the stopping iteration: its apply was skipped (or never enqueued)
        v_axpy(st, n, h[7], D[(info.iterations - 1) & 1], x);
        ++c->stats.kernel_launches;
    }
}

static void gen_rot(double dx, double dy, double& cs, double& sn) {
    if (dy == 0.0) { cs = 1.0; sn = 0.0; }
    else if (std::fabs(dy) > std::fabs(dx)) { const double t = dx / dy; sn = 1.0 / std::sqrt(1.0 + t * t); cs = t * sn; }
    else { const double t = dy / dx; cs = 1.0 / std::sqrt(1.0 + t * t); sn = t * cs; }
}
static void app_rot(double& dx, double& dy, double cs, double sn) {
    const double t = cs * dx + sn * dy;
    dy = -sn * dx + cs * dy;
    dx = t;
}

// mfem::GMRESSolver (left preconditioning, monitors ||B r||) and mfem::FGMRESSolver (right /
// flexible preconditioning, true residual); m = 50, modified Gram-Schmidt
void Krylov::gmres(const double* b, double* x, bool flexible) {
    const int64_t n = c->n;
    cudaStream_t st = c->stream;
    const int m = M;
    // work layout: 0 = r, 1 = tmp, 2..2+m = V, then Z for FGMRES
    double* r = w(0); double* tmp = w(1);
    auto V = [&](int i) { return w(2 + i); };
    auto Z = [&](int i) { return w(2 + m + 1 + i); };
    auto P = [&](const double* in, double* out) { if (prec) prec->mult(in, out); else c->copy(in, out); };
    // JOTS_MGS_SMALL=0 keeps the chained per-vector launches (the path large and partitioned problems take)
    static const bool sweep_ok = [] { const char* e = std::getenv("JOTS_MGS_SMALL"); return !e || e[0] != '0'; }();
    const bool small_mgs = sweep_ok && !c->halo && n <= 32768 && m + 2 <= 64;     // d_scal holds 64 scalars
    bool vptr_valid = false;
    c->zero(x);
    if (flexible) c->copy(b, r); else P(b, r);
    double beta = c->norm(r);
    const double final_norm = std::max(rel_tol * beta, abs_tol);
    if (beta <= final_norm) { info.converged = true; info.final_norm = beta; return; }
    std::vector<double> H((size_t)(m + 1) * m, 0.0), s(m + 1, 0.0), cs(m + 1, 0.0), sn(m + 1, 0.0), yv(m + 2, 0.0);
    auto Hh = [&](int i, int j) -> double& { return H[(size_t)i * m + j]; };
    auto update = [&](int k) {
        for (int i = 0; i <= k; ++i) yv[i] = s[i];
        for (int i = k; i >= 0; --i) {
            yv[i] /= Hh(i, i);
            for (int j = i - 1; j >= 0; --j) yv[j] -= Hh(j, i) * yv[i];
        }
        for (int j = 0; j <= k; ++j) { v_axpy(st, n, yv[j], flexible ? Z(j) : V(j), x); ++c->stats.kernel_launches; }
    };
    int j = 1;
    while (j <= max_iter) {
        v_scale(st, n, 1.0 / beta, r, V(0));
        ++c->stats.kernel_launches;
        std::fill(s.begin(), s.end(), 0.0);
        s[0] = beta;
        int i = 0;
        while (i < m && j <= max_iter) {
            if (flexible) { P(V(i), Z(i)); apply(Z(i), r); }
            else { apply(V(i), tmp); P(tmp, r); }
            // modified Gram-Schmidt, device-chained: step k subtracts the previous projection and forms
            // the next inner product in one pass; the i+2 scalars come back in a single copy
            if (small_mgs) {
                // small problems on one GPU: the whole sweep in one block (vec.cu: k_mgs_sweep_small)
                if (i == 0 || !vptr_valid) {
                    std::vector<const double*> hp(m + 1);
                    for (int k = 0; k <= m; ++k) hp[k] = V(k);
                    d_vptr.upload(hp);
                    vptr_valid = true;
                }
                v_mgs_sweep_small(st, (int)n, r, d_vptr.p, i + 1, c->d_scal.p);
                ++c->stats.kernel_launches;
            } else
            for (int k = 0; k <= i + 1; ++k) {
                v_mgs_step(st, n, c->n_owned, r, k > 0 ? V(k - 1) : nullptr, k > 0 ? c->d_scal.p + (k - 1) : nullptr,
                           k <= i ? V(k) : nullptr, c->dot_scratch(), c->d_scal.p + k);
                ++c->stats.kernel_launches;
                c->allreduce_scalars(c->d_scal.p + k, 1);
            }
            c->copy_scalars(i + 2, yv.data());
            c->stats.dots += i + 2;
            for (int k = 0; k <= i; ++k) Hh(k, i) = yv[k];
            Hh(i + 1, i) = std::sqrt(yv[i + 1]);
            v_scale(st, n, 1.0 / Hh(i + 1, i), r, V(i + 1));
            ++c->stats.kernel_launches;
            for (int k = 0; k < i; ++k) app_rot(Hh(k, i), Hh(k + 1, i), cs[k], sn[k]);
            gen_rot(Hh(i, i), Hh(i + 1, i), cs[i], sn[i]);
            app_rot(Hh(i, i), Hh(i + 1, i), cs[i], sn[i]);
            app_rot(s[i], s[i + 1], cs[i], sn[i]);
            const double resid = std::fabs(s[i + 1]);
            if ((print & PRINT_ITERATIONS) && c->rank() == 0 && !quiet())
                std::printf("   Pass : %2d   Iteration : %3d  || %sr || = %g\n", (j - 1) / m + 1, j, flexible ? "" : "B ", resid);
            if (resid <= final_norm) {
                update(i);
                info.converged = true; info.final_norm = resid; info.iterations = j;
                return;
            }
            ++i; ++j;
        }
        update(i - 1);
        apply(x, tmp);
        if (flexible) v_waxpby(st, n, 1.0, b, -1.0, tmp, r);
        else { v_waxpby(st, n, 1.0, b, -1.0, tmp, tmp); P(tmp, r); }
        ++c->stats.kernel_launches;
        beta = c->norm(r);
        if (beta <= final_norm) { info.converged = true; info.final_norm = beta; info.iterations = j; return; }
    }
    info.final_norm = beta; info.iterations = max_iter;
}

// ------------------------------------------------------------------------------------------------
// Newton (mfem::NewtonSolver::Mult + JOTSNewtonSolver::ProcessNewState)
// ------------------------------------------------------------------------------------------------
void Newton::mult(NewtonOp& op, Krylov& lin, const double* b, double* x) {
    NvtxRange range("jots:newton");
    const int64_t n = c->n;
    r.alloc(n); cvec.alloc(n);
    op.new_state(x);
    op.mult(x, r.p);
    ++c->stats.residual_evals;
    if (b) { v_axpy(c->stream, n, -1.0, b, r.p); ++c->stats.kernel_launches; }
    double norm = c->norm(r.p);
    const double norm0 = norm;
    const double goal = std::max(rel_tol * norm, abs_tol);
    int it = 0;
    while (true) {
        if ((print & PRINT_ITERATIONS) && c->rank() == 0 && !quiet()) {
            std::printf("Newton iteration %2d : ||r|| = %g", it, norm);
            if (it > 0) std::printf(", ||r||/||r_0|| = %g", norm0 > 0 ? norm / norm0 : 0.0);
            std::printf("\n");
        }
        if (!(norm == norm)) { converged = false; break; }
        if (norm <= goal) { converged = true; break; }
        if (it >= max_iter) { converged = false; break; }
        LinOp* J = op.gradient(x);
        lin.set_operator(J);
        lin.mult(r.p, cvec.p);
        v_axpy(c->stream, n, -1.0, cvec.p, x);
        ++c->stats.kernel_launches;
        op.new_state(x);
        op.mult(x, r.p);
        ++c->stats.residual_evals;
        if (b) { v_axpy(c->stream, n, -1.0, b, r.p); ++c->stats.kernel_launches; }
        norm = c->norm(r.p);
        ++it;
    }
    iterations = it;
    final_norm = norm;
    if (print != 0 && c->rank() == 0 && !quiet()) {
        if ((print & PRINT_SUMMARY) || (!converged && (print & (PRINT_WARNINGS | PRINT_ITERATIONS))))
            std::printf("Newton: Number of iterations: %d\n   ||r|| = %g\n", it, norm);
        if (!converged && (print & (PRINT_SUMMARY | PRINT_WARNINGS))) std::printf("Newton: No convergence!\n");
    }
    c->stats.newton_iters += it;
    c->stats.last_newton_iters = it;
    c->stats.last_newton_norm = norm;
    c->stats.last_newton_converged = converged;
}

}  // namespace jots
