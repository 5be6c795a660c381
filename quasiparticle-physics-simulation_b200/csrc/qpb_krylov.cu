// Stiff Crank-Nicolson steps on geometries whose split operators do not commute (slotted / curved masks).
//
// The reference solves (I - a L) u' = b with SuperLU for any step length (qpsim/solver.py:221-232, 1441, 1452).  The
// Peaceman-Rachford iteration of qpb_diffusion.cu reaches the same answer in ~16 iterations at a = dt D / (2 dx^2) of
// order one, but on a mask where Gx and Gy do not commute its cyclic-shift error operator stops contracting once
// hi/lo = 1 + 8a reaches a few hundred (measured: 57 iterations at a = 40, no convergence at a = 150).  From there on
// the SAME two tridiagonal line solves are used as a preconditioner instead of as an iteration:
//
//     M^-1 = 2 r (V + r)^-1 (H + r)^-1,   r = sqrt(lo hi)       (one PR step from u = 0)
//
// inside BiCGStab, one independent system per energy bin, all bins advanced together with their own scalars kept on
// the device (26 iterations at a = 40, 73 at a = 2500 on the slotted test mask; CPU study in DESIGN.md).  The stop test
// is the one of the sweep iteration, ||b - A u||_inf <= tol_bin ||u||_inf, re-checked on the TRUE residual before the
// solve returns.  Sums are reduced through a fixed number of per-block partials, so results are run-to-run identical.
// Works on the uniform-D and the per-cell-D (non-uniform gap) operator alike.  This is a robustness path: one thread
// per line, no TMA pipeline - the pipelined sweeps stay the hot path for the step lengths the solver is used at.
#include "qpb_internal.h"
#include "qpb_faces.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

constexpr int KBLK = 64;    // blocks per bin of the element-wise kernels (fixed: deterministic partial sums)
constexpr int KTHR = 256;
constexpr int NSCAL = 8;    // per-bin scalars: rho, alpha, omega, beta, and spares

struct KryGeom {
    int ne, ny, nx;
    const uint8_t *flags;
    const double *bcx, *bcy, *a_bin, *ex, *ey, *gbx, *gby;
    const int *done;
};

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < KTHR / 32; ++k) t += sh[k];   // fixed order
    return t;
}

__device__ __forceinline__ double block_max(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < KTHR / 32; ++k) t = fmax(t, sh[k]);
    return t;
}

// max that does not swallow a NaN (fmax would): a broken-down recurrence must never look converged
__device__ __forceinline__ double nmax(double m, double v) { return v != v ? INFINITY : fmax(m, v); }

// A v at cell c (A = I + a(Gx + Gy) with the boundary diagonals); cells outside the mask give 0
template <bool VARD>
__device__ __forceinline__ double apply_A(const KryGeom &G, int bin, long long off, int c, const double *v) {
    const unsigned fl = G.flags[c];
    if (!(fl & QPB_IN)) return 0.0;
    const double a = VARD ? 0.0 : G.a_bin[bin];
    const Faces f = load_faces<VARD>(c, G.nx, fl, a, G.bcx, G.bcy, VARD ? G.ex + off : nullptr,
                                     VARD ? G.ey + off : nullptr, VARD ? G.gbx + off : nullptr,
                                     VARD ? G.gby + off : nullptr);
    const double vc = v[c];
    double acc = (f.gbx + f.gby) * vc;
    if (fl & QPB_LK_L) acc += f.eL * (vc - v[c - 1]);
    if (fl & QPB_LK_R) acc += f.eR * (vc - v[c + 1]);
    if (fl & QPB_LK_U) acc += f.eU * (vc - v[c - G.nx]);
    if (fl & QPB_LK_D) acc += f.eD * (vc - v[c + G.nx]);
    return vc + acc;
}

// MODE 0: out = A in;  partial[0] = sum out*w0, partial[1] = sum out*out  (w0 may equal `in`'s sibling vectors)
// MODE 1: out = b - A in (true residual), out2 = out;  partial[0] = max|out|, partial[1] = max|in|
template <bool VARD, int MODE>
__global__ void __launch_bounds__(KTHR) k_kry_matvec(KryGeom G, const double *__restrict__ in, double *__restrict__ out,
                                                     const double *__restrict__ w0, const double *__restrict__ w1,
                                                     double *__restrict__ out2, double *__restrict__ partial) {
    __shared__ double sh[KTHR / 32];
    const int bin = blockIdx.y;
    if (G.done[bin]) return;
    const int ncd = G.ny * G.nx;
    const long long off = (long long)bin * ncd;
    double s0 = 0.0, s1 = 0.0;
    for (int c = blockIdx.x * KTHR + threadIdx.x; c < ncd; c += KBLK * KTHR) {
        const double av = apply_A<VARD>(G, bin, off, c, in + off);
        if (MODE == 0) {
            out[off + c] = av;
            s0 += av * w0[off + c];
            s1 += av * (w1 ? w1[off + c] : av);
        } else {
            const bool inside = G.flags[c] & QPB_IN;
            const double r = inside ? w0[off + c] - av : 0.0;
            out[off + c] = r;
            out2[off + c] = r;
            s0 = nmax(s0, fabs(r));
            s1 = nmax(s1, inside ? fabs(in[off + c]) : 0.0);
        }
    }
    const double t0 = MODE == 0 ? block_sum(s0, sh) : block_max(s0, sh);
    const double t1 = MODE == 0 ? block_sum(s1, sh) : block_max(s1, sh);
    if (threadIdx.x == 0) {
        partial[((long long)bin * KBLK + blockIdx.x) * 2 + 0] = t0;
        partial[((long long)bin * KBLK + blockIdx.x) * 2 + 1] = t1;
    }
}

__global__ void __launch_bounds__(KTHR) k_kry_dot(int ncd, const int *__restrict__ done, const double *__restrict__ a,
                                                  const double *__restrict__ b, double *__restrict__ partial) {
    __shared__ double sh[KTHR / 32];
    const int bin = blockIdx.y;
    if (done[bin]) return;
    const long long off = (long long)bin * ncd;
    double s = 0.0;
    for (int c = blockIdx.x * KTHR + threadIdx.x; c < ncd; c += KBLK * KTHR) s += a[off + c] * b[off + c];
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0) {
        partial[((long long)bin * KBLK + blockIdx.x) * 2 + 0] = t;
        partial[((long long)bin * KBLK + blockIdx.x) * 2 + 1] = 0.0;
    }
}

// per-bin scalar bookkeeping between the vector kernels (one thread per bin)
//   stage 1: rho1 = <rh, r>;  beta = (rho1/rho)(alpha/omega);  rho = rho1
//   stage 2: alpha = rho / <rh, v>
//   stage 3: omega = <t, s> / <t, t>
//   stage 4: stop test on (max|r|, max|x|)
__global__ void k_kry_scalars(int ne, int stage, int iter, double slack, const double *__restrict__ partial,
                              double *__restrict__ scal, const double *__restrict__ tol, int *__restrict__ done,
                              int *__restrict__ iters_out) {
    const int bin = blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= ne || done[bin]) return;
    const double *p = partial + (long long)bin * KBLK * 2;
    double *s = scal + (long long)bin * NSCAL;
    double d0 = 0.0, d1 = 0.0;
    if (stage == 4) {
        for (int k = 0; k < KBLK; ++k) {
            d0 = fmax(d0, p[2 * k]);
            d1 = fmax(d1, p[2 * k + 1]);
        }
        if (d0 <= slack * tol[bin] * d1 && d1 < INFINITY) {
            done[bin] = 1;
            iters_out[bin] = iter;
        }
        return;
    }
    for (int k = 0; k < KBLK; ++k) {
        d0 += p[2 * k];
        d1 += p[2 * k + 1];
    }
    if (stage == 1) {
        const double rho = s[0], alpha = s[1], omega = s[2];
        const double beta = (rho != 0.0 && omega != 0.0) ? (d0 / rho) * (alpha / omega) : 0.0;
        s[3] = beta;
        s[0] = d0;
    } else if (stage == 2) {
        s[1] = d0 != 0.0 ? s[0] / d0 : 0.0;
    } else if (stage == 3) {
        s[2] = d1 != 0.0 ? d0 / d1 : 0.0;
    }
}

// which 0: p = r + beta (p - omega v)      1: r = r - alpha v  (the vector "s" of the method, kept in r)
template <int WHICH>
__global__ void __launch_bounds__(KTHR) k_kry_axpy(int ncd, const int *__restrict__ done, const double *__restrict__ scal,
                                                   double *__restrict__ r, double *__restrict__ p,
                                                   const double *__restrict__ v) {
    const int bin = blockIdx.y;
    if (done[bin]) return;
    const long long off = (long long)bin * ncd;
    const double *s = scal + (long long)bin * NSCAL;
    const double alpha = s[1], omega = s[2], beta = s[3];
    for (int c = blockIdx.x * KTHR + threadIdx.x; c < ncd; c += KBLK * KTHR) {
        if (WHICH == 0) p[off + c] = r[off + c] + beta * (p[off + c] - omega * v[off + c]);
        else r[off + c] = r[off + c] - alpha * v[off + c];
    }
}

// x += alpha y + omega z;  r = s - omega t;  partial = (max|r|, max|x|)
__global__ void __launch_bounds__(KTHR) k_kry_update(int ncd, const int *__restrict__ done, const double *__restrict__ scal,
                                                     double *__restrict__ x, double *__restrict__ r,
                                                     const double *__restrict__ y, const double *__restrict__ z,
                                                     const double *__restrict__ t, double *__restrict__ partial) {
    __shared__ double sh[KTHR / 32];
    const int bin = blockIdx.y;
    if (done[bin]) return;
    const long long off = (long long)bin * ncd;
    const double *s = scal + (long long)bin * NSCAL;
    const double alpha = s[1], omega = s[2];
    double rm = 0.0, xm = 0.0;
    for (int c = blockIdx.x * KTHR + threadIdx.x; c < ncd; c += KBLK * KTHR) {
        const double xn = x[off + c] + alpha * y[off + c] + omega * z[off + c];
        const double rn = r[off + c] - omega * t[off + c];
        x[off + c] = xn;
        r[off + c] = rn;
        rm = nmax(rm, fabs(rn));
        xm = nmax(xm, fabs(xn));
    }
    const double t0 = block_max(rm, sh), t1 = block_max(xm, sh);
    if (threadIdx.x == 0) {
        partial[((long long)bin * KBLK + blockIdx.x) * 2 + 0] = t0;
        partial[((long long)bin * KBLK + blockIdx.x) * 2 + 1] = t1;
    }
}

// out = scale (1/2 + r + a G_dir)^-1 in along every line of direction dir (0: x, 1: y); scale = 2 r when dir == 1.
// One thread per (bin, line); `g` holds the eliminated super-diagonal between the two passes.
template <bool VARD>
__global__ void k_kry_line(KryGeom G, int dir, const double *__restrict__ shift, const double *__restrict__ in,
                           double *__restrict__ out, double *__restrict__ g) {
    const int ncd = G.ny * G.nx;
    const int nlines = dir == 0 ? G.ny : G.nx;
    const int n = dir == 0 ? G.nx : G.ny;
    const int sk = dir == 0 ? 1 : G.nx, sl = dir == 0 ? G.nx : 1;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= (long long)G.ne * nlines) return;
    const int bin = (int)(gid / nlines);
    if (G.done[bin]) return;
    const int line = (int)(gid - (long long)bin * nlines);
    const long long off = (long long)bin * ncd;
    const double a = VARD ? 0.0 : G.a_bin[bin];
    const double rho = shift[bin];
    const double scale = dir == 1 ? 2.0 * rho : 1.0;
    double yprev = 0.0, gprev = 0.0;
    const int c0 = line * sl;
    for (int k = 0; k < n; ++k) {
        const int c = c0 + k * sk;
        const unsigned fl = G.flags[c];
        double eM = 0.0, eP = 0.0, diag = 0.5 + rho, d = 0.0;
        if (fl & QPB_IN) {
            const Faces f = load_faces<VARD>(c, G.nx, fl, a, G.bcx, G.bcy, VARD ? G.ex + off : nullptr,
                                             VARD ? G.ey + off : nullptr, VARD ? G.gbx + off : nullptr,
                                             VARD ? G.gby + off : nullptr);
            if (dir == 0) { eM = f.eL; eP = f.eR; diag += f.eL + f.eR + f.gbx; }
            else          { eM = f.eU; eP = f.eD; diag += f.eU + f.eD + f.gby; }
            d = in[off + c];
        }
        const double m = 1.0 / (diag - eM * gprev);
        const double y = (d + eM * yprev) * m;
        const double gg = eP * m;
        out[off + c] = y;
        g[off + c] = gg;
        yprev = y;
        gprev = gg;
    }
    double xn = 0.0;
    for (int k = n - 1; k >= 0; --k) {
        const int c = c0 + k * sk;
        const double x = out[off + c] + g[off + c] * xn;
        xn = x;
        out[off + c] = (G.flags[c] & QPB_IN) ? scale * x : 0.0;
    }
}

}  // namespace

void qpbk_free_krylov(qpb_ctx *c) {
    for (auto &p : c->d_kry) {
        if (p) qpb_dev_free(p);
        p = nullptr;
    }
    if (c->d_kry_small) qpb_dev_free(c->d_kry_small);
    c->d_kry_small = nullptr;
}

// Solve A u = b for every bin; b in c->d_B (already built), initial guess and result in c->d_S.
int qpbk_diffuse_krylov(qpb_ctx *c, DiffSlot &s) {
    const auto &cf = c->cfg;
    const int ne = cf.ne, ncd = c->ncd;
    const bool vard = cf.flags & QPB_F_VARIABLE_D;
    const size_t nst = (size_t)ne * ncd;
    if (!c->d_kry[0]) {
        for (auto &p : c->d_kry) QPB_CUDA(qpb_dev_malloc((void **)&p, sizeof(double) * nst));
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_kry_small,
                                sizeof(double) * ((size_t)ne * KBLK * 2 + (size_t)ne * NSCAL)));
    }
    double *R = c->d_kry[0], *RH = c->d_kry[1], *P = c->d_kry[2], *V = c->d_kry[3], *Y = c->d_kry[4],
           *Z = c->d_kry[5], *T = c->d_kry[6];
    double *partial = c->d_kry_small, *scal = partial + (size_t)ne * KBLK * 2;
    double *W = c->d_T1, *Gs = c->d_T2;   // line-solve intermediates
    int *done = c->d_done, *iters = c->d_done + ne;
    KryGeom G{ne, cf.ny, cf.nx, c->d_flags, c->d_bcx, c->d_bcy, s.d_a, s.d_ex, s.d_ey, s.d_gbx, s.d_gby, done};
    const dim3 egrid(KBLK, ne);
    const int sblocks = (ne + 127) / 128;
    auto lines = [&](int dir, const double *in, double *out) {
        const long long total = (long long)ne * (dir == 0 ? cf.ny : cf.nx);
        const int blocks = (int)ceil_div64(total, 128);
        if (vard) k_kry_line<true><<<blocks, 128, 0, c->stream>>>(G, dir, s.d_kshift, in, out, Gs);
        else k_kry_line<false><<<blocks, 128, 0, c->stream>>>(G, dir, s.d_kshift, in, out, Gs);
        c->diag.kernel_launches++;
    };
    auto precond = [&](const double *in, double *out) {   // out = M^-1 in
        lines(0, in, W);
        lines(1, W, out);
    };
    auto matvec = [&](const double *in, double *out, const double *w0, const double *w1) {
        if (vard) k_kry_matvec<true, 0><<<egrid, KTHR, 0, c->stream>>>(G, in, out, w0, w1, nullptr, partial);
        else k_kry_matvec<false, 0><<<egrid, KTHR, 0, c->stream>>>(G, in, out, w0, w1, nullptr, partial);
        c->diag.kernel_launches++;
    };
    auto residual = [&]() {   // R = RH = b - A x ; partial = (max|r|, max|x|)
        if (vard) k_kry_matvec<true, 1><<<egrid, KTHR, 0, c->stream>>>(G, c->d_S, R, c->d_B, nullptr, RH, partial);
        else k_kry_matvec<false, 1><<<egrid, KTHR, 0, c->stream>>>(G, c->d_S, R, c->d_B, nullptr, RH, partial);
        c->diag.kernel_launches++;
    };
    auto scalars = [&](int stage, int it, double slack = 1.0) {
        k_kry_scalars<<<sblocks, 128, 0, c->stream>>>(ne, stage, it, slack, partial, scal, s.d_tolk, done, iters);
        c->diag.kernel_launches++;
    };
    std::vector<int> h_done(2 * (size_t)ne);
    std::vector<double> init((size_t)ne * NSCAL, 0.0);
    const int maxit = 600, check_every = 4;
    long long total_it = 0;
    bool all = false;
    for (int round = 0; round < 4 && !all; ++round) {
        // (re)start from the current x: true residual, fresh shadow residual and scalars
        QPB_CUDA(cudaMemsetAsync(done, 0, sizeof(int) * 2 * (size_t)ne, c->stream));
        for (int b = 0; b < ne; ++b) init[(size_t)b * NSCAL + 0] = init[(size_t)b * NSCAL + 1] = init[(size_t)b * NSCAL + 2] = 1.0;
        QPB_CUDA(cudaMemcpyAsync(scal, init.data(), sizeof(double) * init.size(), cudaMemcpyHostToDevice, c->stream));
        QPB_CUDA(cudaMemsetAsync(P, 0, sizeof(double) * nst, c->stream));
        QPB_CUDA(cudaMemsetAsync(V, 0, sizeof(double) * nst, c->stream));
        residual();
        scalars(4, 0);   // bins that already satisfy the system are done
        int it = 0;
        while (it < maxit) {
            const int upto = std::min(maxit, it + check_every);
            for (; it < upto; ++it) {
                k_kry_dot<<<egrid, KTHR, 0, c->stream>>>(ncd, done, RH, R, partial);
                scalars(1, it);
                k_kry_axpy<0><<<egrid, KTHR, 0, c->stream>>>(ncd, done, scal, R, P, V);
                precond(P, Y);
                matvec(Y, V, RH, RH);                    // v = A y ; <v, rh>
                scalars(2, it);
                k_kry_axpy<1><<<egrid, KTHR, 0, c->stream>>>(ncd, done, scal, R, P, V);   // s -> R
                precond(R, Z);
                matvec(Z, T, R, nullptr);                // t = A z ; <t, s>, <t, t>
                scalars(3, it);
                k_kry_update<<<egrid, KTHR, 0, c->stream>>>(ncd, done, scal, c->d_S, R, Y, Z, T, partial);
                scalars(4, it + 1);
                c->diag.kernel_launches += 4;
            }
            QPB_CHECK_LAUNCH();
            QPB_CUDA(cudaMemcpyAsync(h_done.data(), done, sizeof(int) * 2 * (size_t)ne, cudaMemcpyDeviceToHost, c->stream));
            QPB_CUDA(cudaStreamSynchronize(c->stream));
            bool alld = true;
            for (int b = 0; b < ne; ++b) alld = alld && h_done[b];
            if (alld) break;
        }
        total_it += it;
        // the recurrence residual said "converged": confirm on the true residual, restart the bins that disagree
        QPB_CUDA(cudaMemsetAsync(done, 0, sizeof(int) * 2 * (size_t)ne, c->stream));
        residual();
        scalars(4, 0, 4.0);   // rounding of the recurrence against the true residual: a factor of slack, not a restart
        QPB_CUDA(cudaMemcpyAsync(h_done.data(), done, sizeof(int) * 2 * (size_t)ne, cudaMemcpyDeviceToHost, c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));
        all = true;
        for (int b = 0; b < ne; ++b) all = all && h_done[b];
    }
    c->diag.pr_iterations += total_it;
    c->diag.sweeps += 4 * total_it;
    c->diag.bin_sweeps += 4 * total_it * ne;
    if (!all) {
        qpb_set_error("Crank-Nicolson solve (preconditioned BiCGStab) did not reach tolerance %.3g", cf.diff_tol);
        return QPB_E_NOCONV;
    }
    return QPB_OK;
}
