// Crank-Nicolson diffusion of every energy bin on the masked grid.
//
// Replaces qpsim/solver.py:1428-1452 (_apply_diffusion: rhs = B@u + dt*D*s ; u = splu(A).solve(rhs)) and the
// operator build of solver.py:152-212, 221-232, 235-321.  The reference factors the unsplit 2-D operator
// A = I - a*L with SuperLU; here the same linear system is solved to a max-norm residual tolerance with
// batched tridiagonal x/y sweeps:
//   H = I/2 + a*Gx,  V = I/2 + a*Gy,  A = H + V   (Gx, Gy = minus the x / y parts of the 5-point Laplacian)
//   (H + r) u* = b - (V - r) u          x sweep, r = shift of this iteration
//   u+ = u + 2 r (V + r)^-1 (u* - u)    y sweep (algebraically the second Peaceman-Rachford half step)
// One-cell-thick geometries are solved by a single direct sweep (A itself is tridiagonal there).
// The x sweep also evaluates the true residual b - A u of its input, so convergence is decided on
// ||b - A u||_inf <= tol * ||u||_inf, which bounds the error because ||A^-1||_inf <= 1.
//
// This file holds the generic kernels (any mask, any boundary kinds, uniform or per-cell D): one thread per
// line, pivots by division.  qpb_sweep_fast.cu holds the chunked, table-driven kernels used for uniform D.
#include "qpb_internal.h"
#include "qpb_faces.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

// b = (I + a L) u + dt*D*s       (solver.py:1440, 1451)
template <bool VARD>
__global__ void k_build_rhs(int ne, int ny, int nx, const double *__restrict__ S, double *__restrict__ B,
                            const uint8_t *__restrict__ flags, const double *__restrict__ bcx,
                            const double *__restrict__ bcy, const double *__restrict__ a_bin,
                            const double *__restrict__ srccoef, const double *__restrict__ src,
                            const double *__restrict__ ex, const double *__restrict__ ey,
                            const double *__restrict__ gbx, const double *__restrict__ gby) {
    const int ncd = ny * nx;
    const int bin = blockIdx.y;   // grid.y = energy bin: no per-element index division
    const long long off = (long long)bin * ncd;
    const double *u = S + off;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncd; c += gridDim.x * blockDim.x) {
        const long long g = off + c;
        const unsigned fl = flags[c];
        if (!(fl & QPB_IN)) {
            B[g] = 0.0;
            continue;
        }
        const double a = VARD ? 0.0 : a_bin[bin];
        Faces f = load_faces<VARD>(c, nx, fl, a, bcx, bcy, VARD ? ex + off : nullptr, VARD ? ey + off : nullptr,
                                   VARD ? gbx + off : nullptr, VARD ? gby + off : nullptr);
        const double uc = u[c];
        double acc = 0.0;
        if (fl & QPB_LK_L) acc += f.eL * (u[c - 1] - uc);
        if (fl & QPB_LK_R) acc += f.eR * (u[c + 1] - uc);
        if (fl & QPB_LK_U) acc += f.eU * (u[c - nx] - uc);
        if (fl & QPB_LK_D) acc += f.eD * (u[c + nx] - uc);
        acc -= (f.gbx + f.gby) * uc;
        const double s = VARD ? src[off + c] : srccoef[bin] * src[c];
        B[g] = uc + acc + s;
    }
}

// One thread per (bin, line).  mode 0: PR sweep along x, 1: PR sweep along y, 2: direct solve along `dir`.
template <bool VARD>
__global__ void k_sweep_generic(int ne, int ny, int nx, int dir, int mode, int iter, const double *__restrict__ tol,
                                double *__restrict__ S, const double *__restrict__ B, double *__restrict__ T1,
                                double *__restrict__ T2, const uint8_t *__restrict__ flags,
                                const double *__restrict__ bcx, const double *__restrict__ bcy,
                                const double *__restrict__ a_bin, const double *__restrict__ shift,
                                const int *__restrict__ jlen, int jmax, const double *__restrict__ ex,
                                const double *__restrict__ ey, const double *__restrict__ gbx,
                                const double *__restrict__ gby, unsigned long long *__restrict__ res,
                                unsigned long long *__restrict__ unorm, int *__restrict__ done,
                                int *__restrict__ iters_out) {
    const int ncd = ny * nx;
    const int nlines = dir == 0 ? ny : nx;
    const int n = dir == 0 ? nx : ny;
    const int sk = dir == 0 ? 1 : nx;   // stride along the line
    const int sl = dir == 0 ? nx : 1;   // stride between lines
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= (long long)ne * nlines) return;
    const int bin = (int)(gid / nlines);
    const int line = (int)(gid - (long long)bin * nlines);
    const long long off = (long long)bin * ncd;

    if (mode != 2) {
        if (done[bin]) return;
        if (mode == 1) {
            const double r = __longlong_as_double((long long)res[(long long)iter * ne + bin]);
            const double un = __longlong_as_double((long long)unorm[(long long)iter * ne + bin]);
            (void)un;
            if (r <= 0.0) {  // the input of this iteration satisfies the system in every cell (componentwise bound)
                if (line == 0) {
                    done[bin] = 1;
                    iters_out[bin] = iter;
                }
                return;
            }
        }
    }
    const double a = VARD ? 0.0 : a_bin[bin];
    const double rho = mode == 2 ? 0.0 : shift[(long long)bin * jmax + (iter % jlen[bin])];
    const double sigma = mode == 2 ? 1.0 : 0.5;
    const double *u = S + off;
    const double *b = B + off;
    double *t1 = T1 + off;
    double *t2 = T2 + off;
    const double *pex = VARD ? ex + off : nullptr;
    const double *pey = VARD ? ey + off : nullptr;
    const double *pgx = VARD ? gbx + off : nullptr;
    const double *pgy = VARD ? gby + off : nullptr;

    double yprev = 0.0, gprev = 0.0;
    double rmax = 0.0, umax = 0.0;
    const int c0 = line * sl;
    for (int k = 0; k < n; ++k) {
        const int c = c0 + k * sk;
        const unsigned fl = flags[c];
        double eM = 0.0, eP = 0.0, diag = sigma + rho, d = 0.0;
        if (fl & QPB_IN) {
            Faces f = load_faces<VARD>(c, nx, fl, a, bcx, bcy, pex, pey, pgx, pgy);
            double eCm, eCp, gl, gc;
            int sc;
            if (dir == 0) { eM = f.eL; eP = f.eR; eCm = f.eU; eCp = f.eD; gl = f.gbx; gc = f.gby; sc = nx; }
            else          { eM = f.eU; eP = f.eD; eCm = f.eL; eCp = f.eR; gl = f.gby; gc = f.gbx; sc = 1; }
            diag += eM + eP + gl;
            const double uc = u[c];
            if (mode == 0) {
                // cross operator on u, and the along-line operator for the residual
                const double au = fabs(uc);
                double cross = gc * uc, wsum = (fabs(gc) + fabs(gl)) * au;
                if (eCm != 0.0) { cross += eCm * (uc - u[c - sc]); wsum += eCm * (au + fabs(u[c - sc])); }
                if (eCp != 0.0) { cross += eCp * (uc - u[c + sc]); wsum += eCp * (au + fabs(u[c + sc])); }
                double along = gl * uc;
                if (eM != 0.0) { along += eM * (uc - u[c - sk]); wsum += eM * (au + fabs(u[c - sk])); }
                if (eP != 0.0) { along += eP * (uc - u[c + sk]); wsum += eP * (au + fabs(u[c + sk])); }
                const double bc_ = b[c];
                d = bc_ + (rho - 0.5) * uc - cross;
                // componentwise stop test: excess of |b - A u| over tol (|A||u| + |b|) in this cell
                rmax = fmax(rmax, fabs(bc_ - uc - cross - along) - tol[bin] * (fabs(bc_) + au + wsum));
                umax = fmax(umax, au);
            } else if (mode == 1) {
                d = t1[c] - uc;
            } else {
                diag += gc;
                d = b[c];
            }
        }
        const double m = 1.0 / (diag - eM * gprev);
        const double y = (d + eM * yprev) * m;
        const double g = eP * m;
        t1[c] = y;
        t2[c] = g;
        yprev = y;
        gprev = g;
    }
    double xn = 0.0;
    for (int k = n - 1; k >= 0; --k) {
        const int c = c0 + k * sk;
        const double x = t1[c] + t2[c] * xn;
        xn = x;
        if (mode == 0) t1[c] = x;
        else if (mode == 1) S[off + c] = u[c] + 2.0 * rho * x;
        else S[off + c] = x;
    }
    if (mode == 0) {
        atomicMax(&res[(long long)iter * ne + bin], (unsigned long long)__double_as_longlong(rmax));
        atomicMax(&unorm[(long long)iter * ne + bin], (unsigned long long)__double_as_longlong(umax));
    }
}

}  // namespace

int qpbk_build_rhs(qpb_ctx *c, DiffSlot &s) {
    const auto &cf = c->cfg;
    const bool vard = cf.flags & QPB_F_VARIABLE_D;
    const int threads = 256;
    const dim3 blocks((unsigned)std::max(1, std::min((c->ncd + threads - 1) / threads, 256)), (unsigned)cf.ne);
    if (vard)
        k_build_rhs<true><<<blocks, threads, 0, c->stream>>>(cf.ne, cf.ny, cf.nx, c->d_S, c->d_B, c->d_flags,
                                                             c->d_bcx, c->d_bcy, s.d_a, nullptr, s.d_src, s.d_ex,
                                                             s.d_ey, s.d_gbx, s.d_gby);
    else
        k_build_rhs<false><<<blocks, threads, 0, c->stream>>>(cf.ne, cf.ny, cf.nx, c->d_S, c->d_B, c->d_flags,
                                                              c->d_bcx, c->d_bcy, s.d_a, s.d_src, c->d_srcgeom,
                                                              nullptr, nullptr, nullptr, nullptr);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_sweep_generic(qpb_ctx *c, DiffSlot &s, int dir, int iter, int mode) {
    const auto &cf = c->cfg;
    const bool vard = cf.flags & QPB_F_VARIABLE_D;
    const int nlines = dir == 0 ? cf.ny : cf.nx;
    const long long total = (long long)cf.ne * nlines;
    const int threads = 128;
    const int blocks = (int)ceil_div64(total, threads);
    const double *tol = s.d_tol;
    ScopedTimer tm(c, dir == 0 ? 0 : 1);
    if (vard)
        k_sweep_generic<true><<<blocks, threads, 0, c->stream>>>(
            cf.ne, cf.ny, cf.nx, dir, mode, iter, tol, c->d_S, c->d_B, c->d_T1, c->d_T2, c->d_flags, c->d_bcx,
            c->d_bcy, s.d_a, s.d_shift, s.d_jlen, s.jmax, s.d_ex, s.d_ey, s.d_gbx, s.d_gby, c->d_res, c->d_unorm,
            c->d_done, c->d_done + cf.ne);
    else
        k_sweep_generic<false><<<blocks, threads, 0, c->stream>>>(
            cf.ne, cf.ny, cf.nx, dir, mode, iter, tol, c->d_S, c->d_B, c->d_T1, c->d_T2, c->d_flags, c->d_bcx,
            c->d_bcy, s.d_a, s.d_shift, s.d_jlen, s.jmax, nullptr, nullptr, nullptr, nullptr, c->d_res,
            c->d_unorm, c->d_done, c->d_done + cf.ne);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

// Solve the CN system of every bin for the prepared step length (state in c->d_S, in place).
int qpbk_diffuse(qpb_ctx *c, DiffSlot &s) {
    const auto &cf = c->cfg;
    const int ne = cf.ne;
    int rc;
    if (s.mode == 0 && !s.spectral && !s.krylov && s.res.ok) {
        // bin-resident solve (qpb_resident.cu): right-hand side, iteration and stop test in one launch
        std::vector<int> hd;
        rc = qpbr_solve(c, s, hd);
        if (rc != QPB_OK && rc != QPB_E_NOCONV) return rc;
        int kmax = 0;
        long long bs = 0;
        for (int b = 0; b < ne; ++b) {
            kmax = std::max(kmax, hd[ne + b]);
            bs += 2 * hd[ne + b] + 1;
        }
        c->diag.sweeps += 2 * kmax + 1;
        c->diag.bin_sweeps += bs;
        if (rc == QPB_E_NOCONV) {
            // same fall-back as the launched iteration below: this and every later step of the slot by the Krylov solve
            if (getenv("QPB_NO_KRYLOV") && getenv("QPB_NO_KRYLOV")[0] == '1') {
                qpb_set_error("Crank-Nicolson sweep iteration did not reach tolerance %.3g in %d iterations", cf.diff_tol,
                              c->maxit);
                return QPB_E_NOCONV;
            }
            QPB_CUDA(cudaMemcpyAsync(c->d_S, c->d_B, sizeof(double) * (size_t)ne * c->ncd, cudaMemcpyDeviceToDevice,
                                     c->stream));
            s.krylov = true;
            return qpbk_diffuse_krylov(c, s);
        }
        s.known_iters = kmax;
        s.solves++;
        QPB_CUDA(cudaMemcpyAsync(s.d_known, hd.data() + ne, sizeof(int) * ne, cudaMemcpyHostToDevice, c->stream));
        c->diag.pr_iterations += kmax;
        return QPB_OK;
    }
    if (s.mode == 0 && s.spectral && s.sp_fused) return qpbk_diffuse_spectral(c, s);   // forms its right-hand side itself
    if ((rc = qpbk_build_rhs(c, s)) != QPB_OK) return rc;
    if (s.mode != 0) {
        const int dir = s.mode == 1 ? 0 : 1;
        rc = s.fast ? qpbk_sweep_fast(c, s, dir, 0, 2) : qpbk_sweep_generic(c, s, dir, 0, 2);
        if (rc != QPB_OK) return rc;
        c->diag.sweeps += 1;
        c->diag.bin_sweeps += ne;
        return QPB_OK;
    }
    if (s.spectral) return qpbk_diffuse_spectral(c, s);
    if (s.krylov) return qpbk_diffuse_krylov(c, s);
    QPB_CUDA(cudaMemsetAsync(c->d_res, 0, sizeof(unsigned long long) * (size_t)c->maxit * ne, c->stream));
    QPB_CUDA(cudaMemsetAsync(c->d_unorm, 0, sizeof(unsigned long long) * (size_t)c->maxit * ne, c->stream));
    QPB_CUDA(cudaMemsetAsync(c->d_done, 0, sizeof(int) * 2 * (size_t)ne, c->stream));
    std::vector<int> h_done(2 * (size_t)ne);
    int it = 0;
    int batch = std::max(1, s.launch_iters);
    bool all = false;
    while (it < c->maxit) {
        const int upto = std::min(c->maxit, it + batch);
        for (; it < upto; ++it) {
            // The residual is measured only where convergence can happen: a bin whose previous solve needed k
            // iterations runs its first k - 2 iterations unchecked (consecutive time steps need about as many; both
            // pipelined kernels must be in use: only their y sweep can ignore the residual record).  Every 16th
            // solve checks from the start, so a solve that gets easier over time is noticed.
            const bool both = s.fast && s.pipe.x_ok && s.pipe.y_ok;
            const bool check = !(both && s.known_iters > 0 && (s.solves % 16) != 0);   // false: per-bin decision on the device
            rc = s.fast ? qpbk_sweep_fast(c, s, 0, it, 0, check) : qpbk_sweep_generic(c, s, 0, it, 0);
            if (rc != QPB_OK) return rc;
            rc = s.fast ? qpbk_sweep_fast(c, s, 1, it, 1, check) : qpbk_sweep_generic(c, s, 1, it, 1);
            if (rc != QPB_OK) return rc;
        }
        QPB_CUDA(cudaMemcpyAsync(h_done.data(), c->d_done, sizeof(int) * 2 * (size_t)ne, cudaMemcpyDeviceToHost,
                                 c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));
        all = true;
        for (int b = 0; b < ne; ++b) all = all && h_done[b];
        if (all) break;
        batch = 1;
    }
    if (!all) {
        // The sweep iteration stalled (it has no convergence guarantee on non-commuting geometries): solve this and
        // every later step of the slot with the preconditioned Krylov method, from b as the initial guess (the
        // iterate may have grown without bound).
        if (getenv("QPB_NO_KRYLOV") && getenv("QPB_NO_KRYLOV")[0] == '1') {
            qpb_set_error("Crank-Nicolson sweep iteration did not reach tolerance %.3g in %d iterations", cf.diff_tol,
                          c->maxit);
            return QPB_E_NOCONV;
        }
        QPB_CUDA(cudaMemcpyAsync(c->d_S, c->d_B, sizeof(double) * (size_t)ne * c->ncd, cudaMemcpyDeviceToDevice,
                                 c->stream));
        s.krylov = true;
        c->diag.sweeps += 2 * it;
        return qpbk_diffuse_krylov(c, s);
    }
    if (const char *dbg = getenv("QPB_DEBUG_RES")) {   // residual history of the first and last bin (diagnostics)
        if (dbg[0] == '1') {
            std::vector<unsigned long long> hr((size_t)it * ne), hu((size_t)it * ne);
            cudaMemcpy(hr.data(), c->d_res, sizeof(unsigned long long) * hr.size(), cudaMemcpyDeviceToHost);
            cudaMemcpy(hu.data(), c->d_unorm, sizeof(unsigned long long) * hu.size(), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[qpb] iterations per bin:");
            for (int b = 0; b < ne; ++b) fprintf(stderr, " %d", h_done[ne + b]);
            fprintf(stderr, "  (launched %d)\n", it);
            for (int b : {0, ne - 1}) {
                fprintf(stderr, "[qpb] bin %d done at %d:", b, h_done[ne + b]);
                for (int k = 0; k < it; ++k) {
                    double r, u;
                    memcpy(&r, &hr[(size_t)k * ne + b], 8);
                    memcpy(&u, &hu[(size_t)k * ne + b], 8);
                    fprintf(stderr, " %.1e", u > 0 ? r / u : -1.0);
                }
                fprintf(stderr, "\n");
            }
        }
    }
    int kmax = 0;
    long long bs = 0;
    for (int b = 0; b < ne; ++b) {
        const int k = h_done[ne + b];  // iteration whose input was converged: k y-sweeps, k+1 x-sweeps
        kmax = std::max(kmax, k);
        bs += 2 * k + 1;
    }
    // next call: launch exactly what was needed this time (+1 detects convergence) before the first host check
    s.launch_iters = kmax + 1;
    s.known_iters = kmax;
    s.solves++;
    if (s.d_known)
        QPB_CUDA(cudaMemcpyAsync(s.d_known, h_done.data() + ne, sizeof(int) * ne, cudaMemcpyHostToDevice, c->stream));
    c->diag.pr_iterations += kmax;
    c->diag.sweeps += 2 * it;
    c->diag.bin_sweeps += bs;
    return QPB_OK;
}
