// Direct Crank-Nicolson solve on full rectangles whose left / right walls are reflective (or carry a flux): BASELINE
// configs 2 and 3 (2048^2 x 256 bins, 1024^2 x 512 bins).
//
// The reference factors A = I - a L per bin with SuperLU (qpsim/solver.py:221-232) and solves (I - aL) u' = b
// (:1441, :1452).  On such a grid the x part of the operator, Gx (second difference with zero-flux ends), is the same
// in every row and is diagonalised by the cosine transform: Gx v_k = lam_k v_k, v_k(j) = cos(pi k (j + 1/2) / nx),
// lam_k = 4 sin^2(pi k / (2 nx)).  In that basis the 2-D system falls apart into one tridiagonal system along y per
// mode k and bin:
//
//      b^ = DCT_x(b)                                   rows, two at a time as one complex FFT in shared memory
//      (1 + a lam_k + a Gy) u^_k = b^_k                Thomas along y, one thread per (bin, k), coalesced over k
//      u' = DCT_x^-1(u^)
//
// - a direct solve (no iteration, no tolerance): three passes over the bin instead of the 22 line sweeps the
// Peaceman-Rachford iteration needs at 2048^2.  The top / bottom walls may be of any kind that is uniform along the
// wall (their diagonal terms enter Gy row by row), sources of any shape enter through b.  Everything else (masks,
// walls with a diagonal term on the left / right, row lengths that are not a power of two) stays with the sweep
// iteration.
#include "qpb_internal.h"

#include <cmath>
#include <cstdlib>
#include <vector>

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// Shared-memory slot of element e.  The transforms are fed in bit-reversed order: consecutive lanes write elements that
// differ in bits 6.. of e, i.e. 16-byte slots a multiple of 1 KB apart - all in the same banks (ncu on the first
// version: 130 M bank conflicts per 16 bins, shared-memory pipe 92 % busy).  XOR-ing those bits into the low three
// spreads such a warp over all eight 128-byte bank groups and leaves runs of consecutive elements conflict free.
__device__ __forceinline__ int slot(int e) { return e ^ ((e >> 6) & 7); }

// in-place decimation-in-time passes over z[0..N) (input in bit-reversed order), two radix-2 stages fused per trip
// through shared memory (four elements in registers: half the passes and barriers of a plain radix-2 walk).  tw is
// packed by stage: tw[half + pos] = exp(-2 pi i pos / (2 half)) for pos < half (contiguous reads in every stage),
// conjugated for the inverse transform.
template <bool INV>
__device__ __forceinline__ void fft_passes(double2 *z, const double2 *tw, int N, int logN) {
    int s = 0;
    if (logN & 1) {   // odd number of stages: the first one (twiddle 1) on its own
        __syncthreads();
        for (int q = threadIdx.x; q < N / 2; q += blockDim.x) {
            const int ia = slot(2 * q), ib = slot(2 * q + 1);
            const double2 a = z[ia], t = z[ib];
            z[ia] = make_double2(a.x + t.x, a.y + t.y);
            z[ib] = make_double2(a.x - t.x, a.y - t.y);
        }
        s = 1;
    }
    for (; s < logN; s += 2) {
        const int h = 1 << s;
        __syncthreads();
        for (int q = threadIdx.x; q < N / 4; q += blockDim.x) {
            const int pos = q & (h - 1);
            const int i0 = ((q >> s) << (s + 2)) + pos;
            double2 w1 = tw[h + pos], w2 = tw[2 * h + pos];
            if (INV) {
                w1.y = -w1.y;
                w2.y = -w2.y;
            }
            // twiddle of the second pair of the later stage: exp(-+ i pi / 2) times w2
            const double2 w3 = INV ? make_double2(-w2.y, w2.x) : make_double2(w2.y, -w2.x);
            const int s0 = slot(i0), s1 = slot(i0 + h), s2 = slot(i0 + 2 * h), s3 = slot(i0 + 3 * h);
            const double2 a0 = z[s0], a1 = z[s1], a2 = z[s2], a3 = z[s3];
            double2 t = cmul(w1, a1);
            const double2 b0 = make_double2(a0.x + t.x, a0.y + t.y), b1 = make_double2(a0.x - t.x, a0.y - t.y);
            t = cmul(w1, a3);
            const double2 b2 = make_double2(a2.x + t.x, a2.y + t.y), b3 = make_double2(a2.x - t.x, a2.y - t.y);
            t = cmul(w2, b2);
            z[s0] = make_double2(b0.x + t.x, b0.y + t.y);
            z[s2] = make_double2(b0.x - t.x, b0.y - t.y);
            t = cmul(w3, b3);
            z[s1] = make_double2(b1.x + t.x, b1.y + t.y);
            z[s3] = make_double2(b1.x - t.x, b1.y - t.y);
        }
    }
    __syncthreads();
}

// position of sample j of a row in the reordered sequence of the cosine transform (even samples ascending, then the
// odd ones descending)
__device__ __forceinline__ int reorder(int j, int N) { return (j & 1) ? N - 1 - (j >> 1) : (j >> 1); }

// forward: out[row][k] = sum_j in[row][j] cos(pi k (j + 1/2) / N) for the rows 2p, 2p+1 of bin blockIdx.y
__global__ void k_dct_forward(int ny, int N, int logN, const double *__restrict__ in, double *__restrict__ out,
                              const double2 *__restrict__ twg, const double2 *__restrict__ tw2) {
    extern __shared__ __align__(16) double2 smz[];
    double2 *z = smz;
    const double2 *tw = twg;   // twiddles straight from L1 / L2 (32 KB shared by every CTA): half the shared memory, twice the CTAs per SM
    const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
    const bool two = r1 < ny;
    const size_t base = (size_t)blockIdx.y * ny * N;
    const double *a = in + base + (size_t)r0 * N, *b = in + base + (size_t)r1 * N;
    // four row elements per thread in flight before the first shared-memory store (N is a multiple of 4 blockDim.x
    // for the row lengths the launch uses: 128 threads below 1024, 256 from there on)
    for (int j0 = threadIdx.x; j0 < N; j0 += 4 * blockDim.x) {
        double va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * blockDim.x;
            va[u] = j < N ? a[j] : 0.0;
            vb[u] = (two && j < N) ? b[j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * blockDim.x;
            if (j < N) z[slot(__brev((unsigned)reorder(j, N)) >> (32 - logN))] = make_double2(va[u], vb[u]);
        }
    }
    fft_passes<false>(z, tw, N, logN);
    double *oa = out + base + (size_t)r0 * N, *ob = out + base + (size_t)r1 * N;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double2 zk = z[slot(k)], zr = z[slot((N - k) & (N - 1))];
        // spectra of the two real rows packed into one complex transform
        const double2 va = make_double2(0.5 * (zk.x + zr.x), 0.5 * (zk.y - zr.y));
        const double2 vb = make_double2(0.5 * (zk.y + zr.y), 0.5 * (zr.x - zk.x));
        const double2 w = tw2[k];                       // exp(-i pi k / (2N))
        oa[k] = fma(w.x, va.x, -w.y * va.y);
        if (two) ob[k] = fma(w.x, vb.x, -w.y * vb.y);
    }
}

// inverse: out[row][j] = (1/N) (c_0 + 2 sum_{k>=1} c_k cos(pi k (j + 1/2) / N))
__global__ void k_dct_inverse(int ny, int N, int logN, const double *__restrict__ in, double *__restrict__ out,
                              const double2 *__restrict__ twg, const double2 *__restrict__ tw2) {
    extern __shared__ __align__(16) double2 smz[];
    double2 *z = smz;
    const double2 *tw = twg;   // twiddles straight from L1 / L2 (32 KB shared by every CTA): half the shared memory, twice the CTAs per SM
    const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
    const bool two = r1 < ny;
    const size_t base = (size_t)blockIdx.y * ny * N;
    const double *a = in + base + (size_t)r0 * N, *b = in + base + (size_t)r1 * N;
#pragma unroll 2
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double ca = a[k], car = k ? a[N - k] : 0.0;
        const double cb = two ? b[k] : 0.0, cbr = (two && k) ? b[N - k] : 0.0;
        double2 w = tw2[k];
        w.y = -w.y;                                     // exp(+i pi k / (2N))
        double2 va = cmul(w, make_double2(ca, -car)), vb = cmul(w, make_double2(cb, -cbr));
        if (k == 0) {
            va = make_double2(ca, 0.0);
            vb = make_double2(cb, 0.0);
        }
        // Z = VA + i VB
        z[slot(__brev((unsigned)k) >> (32 - logN))] = make_double2(va.x - vb.y, va.y + vb.x);
    }
    fft_passes<true>(z, tw, N, logN);
    const double inv = 1.0 / N;
    double *oa = out + base + (size_t)r0 * N, *ob = out + base + (size_t)r1 * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const double2 v = z[slot(reorder(j, N))];
        oa[j] = v.x * inv;
        if (two) ob[j] = v.y * inv;
    }
}

// (1 + a lam_k + a Gy) x = b along y for every mode k of every bin, in place in `v`; `g` keeps the eliminated
// super-diagonal between the two passes.  One thread per (bin, k): consecutive threads are consecutive k (coalesced).
__global__ void k_thomas_modes(int ne, int ny, int N, double *__restrict__ v, double *__restrict__ g,
                               const double *__restrict__ a_bin, const double *__restrict__ lam,
                               const double *__restrict__ bcy_row) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int bin = blockIdx.y;
    if (k >= N) return;
    const double a = a_bin[bin];
    const double d0 = fma(a, lam[k], 1.0);
    double *col = v + (size_t)bin * ny * N + k, *gc = g + (size_t)bin * ny * N + k;
    double yp = 0.0, gp = 0.0;
#pragma unroll 4
    for (int t = 0; t < ny; ++t) {
        const double deg = (t > 0 ? 1.0 : 0.0) + (t < ny - 1 ? 1.0 : 0.0);
        const double diag = fma(a, deg + bcy_row[t], d0);
        const double m = 1.0 / fma(-a, gp, diag);
        const double y = fma(a, yp, col[(size_t)t * N]) * m;
        const double gg = t < ny - 1 ? a * m : 0.0;
        col[(size_t)t * N] = y;
        gc[(size_t)t * N] = gg;
        yp = y;
        gp = gg;
    }
    double xn = 0.0;
#pragma unroll 4
    for (int t = ny - 1; t >= 0; --t) {
        xn = fma(gc[(size_t)t * N], xn, col[(size_t)t * N]);
        col[(size_t)t * N] = xn;
    }
}

// Pivots of the Thomas recurrence do not depend on the data: m_0 = 1 / (d + a (1 + wall_0)), m_t = 1 / (d + 2a -
// a^2 m_{t-1}) in the interior, d = 1 + a lam_k.  The recurrence contracts by (a m)^2 per row, so after T rows (T from
// the slowest bin and mode, chosen on the host) m_t has reached its fixed point to below an ulp: the first T pivots of
// every (bin, mode) go into a table once per prepared step length, every later row uses the last of them.
__global__ void k_thomas_pivots(int T, int N, double *__restrict__ tab, const double *__restrict__ a_bin,
                                const double *__restrict__ lam, double wall0) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int bin = blockIdx.y;
    if (k >= N) return;
    const double a = a_bin[bin];
    const double d0 = fma(a, lam[k], 1.0);
    double m = 1.0 / fma(a, 1.0 + wall0, d0);
    double *col = tab + (size_t)bin * T * N + k;
    col[0] = m;
    for (int t = 1; t < T; ++t) {
        m = 1.0 / fma(-a * a, m, fma(a, 2.0, d0));
        col[(size_t)t * N] = m;
    }
}

// The same solve as k_thomas_modes with the tabulated / frozen pivots: no division and no second array between the
// passes - b^ in and y out on the way down, y in and u^ out on the way up (32 B per cell and bin).
__global__ void k_thomas_frozen(int ny, int N, int T, double *__restrict__ v, const double *__restrict__ tab,
                                const double *__restrict__ a_bin, const double *__restrict__ lam, double wall_last) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int bin = blockIdx.y;
    if (k >= N) return;
    const double a = a_bin[bin];
    double *col = v + (size_t)bin * ny * N + k;
    const double *piv = tab + (size_t)bin * T * N + k;
    const double mf = piv[(size_t)(T - 1) * N];
    const double m_last = 1.0 / fma(-a * a, mf, fma(a, 1.0 + wall_last, fma(a, lam[k], 1.0)));
    double yp = 0.0;
    for (int t = 0; t < T; ++t) {
        yp = fma(a, yp, col[(size_t)t * N]) * piv[(size_t)t * N];
        col[(size_t)t * N] = yp;
    }
    // The rows below are a first-order recurrence: one dependent FMA + MUL per row, but the loads do not depend on it.
    // The compiler cannot move a load of row t+1 above the store of row t (the row stride is a run-time value), so the
    // walk is pipelined by hand: the next PF rows are in registers before the current ones are stored - PF loads in
    // flight per thread instead of one (the first version spent two memory latencies per row and thread).
    constexpr int PF = 8;
    {
        double cur[PF], nxt[PF];
        int t = T;
#pragma unroll
        for (int u = 0; u < PF; ++u) cur[u] = t + u < ny - 1 ? col[(size_t)(t + u) * N] : 0.0;
        for (; t < ny - 1; t += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) nxt[u] = t + PF + u < ny - 1 ? col[(size_t)(t + PF + u) * N] : 0.0;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (t + u < ny - 1) {
                    yp = fma(a, yp, cur[u]) * mf;
                    col[(size_t)(t + u) * N] = yp;
                }
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) cur[u] = nxt[u];
        }
    }
    double xn = fma(a, yp, col[(size_t)(ny - 1) * N]) * m_last;
    col[(size_t)(ny - 1) * N] = xn;
    const double gf = a * mf;
    {
        double cur[PF], nxt[PF];
        int t = ny - 2;
#pragma unroll
        for (int u = 0; u < PF; ++u) cur[u] = t - u >= T - 1 ? col[(size_t)(t - u) * N] : 0.0;
        for (; t >= T - 1; t -= PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) nxt[u] = t - PF - u >= T - 1 ? col[(size_t)(t - PF - u) * N] : 0.0;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (t - u >= T - 1) {
                    xn = fma(gf, xn, cur[u]);
                    col[(size_t)(t - u) * N] = xn;
                }
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) cur[u] = nxt[u];
        }
    }
    for (int t = T - 2; t >= 0; --t) {
        xn = fma(a * piv[(size_t)t * N], xn, col[(size_t)t * N]);
        col[(size_t)t * N] = xn;
    }
}

// k_thomas_frozen with the right-hand side formed on the way down: the input is the transform of u itself, not of
// b = (I + a L) u + dt D s.  In the cosine basis the x part of L is the number lam_k, so
//     b^_k(t) = (1 - a (lam_k + links_t + wall_t)) u^_k(t) + a (u^_k(t-1) + u^_k(t+1)) + dt D s^_k(t)
// (links_t = 2 inside, 1 in the first and last row; wall terms only there; s^ = transform of the boundary sources, absent
// when they are all zero) needs nothing but the neighbouring rows the walk already holds in registers: the k_build_rhs
// pass over the bin (16 B per cell and bin) and the array b fall away.
__global__ void k_thomas_fused(int ny, int N, int T, double *__restrict__ v, const double *__restrict__ tab,
                               const double *__restrict__ a_bin, const double *__restrict__ lam, double wall0,
                               double wall_last, const double *__restrict__ srchat, const double *__restrict__ srccoef) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int bin = blockIdx.y;
    if (k >= N) return;
    const double a = a_bin[bin], lk = lam[k];
    const double sc = srchat ? srccoef[bin] : 0.0;
    double *col = v + (size_t)bin * ny * N + k;
    const double *piv = tab + (size_t)bin * T * N + k;
    const double *sh = srchat ? srchat + k : nullptr;
    const double mf = piv[(size_t)(T - 1) * N];
    const double m_last = 1.0 / fma(-a * a, mf, fma(a, 1.0 + wall_last, fma(a, lk, 1.0)));
    const double c_in = 1.0 - a * (lk + 2.0), c_first = 1.0 - a * (lk + 1.0 + wall0), c_last = 1.0 - a * (lk + 1.0 + wall_last);
    constexpr int PF = 8;
    double yp = 0.0, um = 0.0;   // y of the previous row, u^ of the previous row
    {
        double cur[PF + 1], nxt[PF];   // cur[PF]: the row behind the batch (needed as the lower neighbour of its last row)
        int t = 0;
#pragma unroll
        for (int u = 0; u <= PF; ++u) cur[u] = t + u < ny ? col[(size_t)(t + u) * N] : 0.0;
        for (; t < ny - 1; t += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) nxt[u] = t + PF + 1 + u < ny ? col[(size_t)(t + PF + 1 + u) * N] : 0.0;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int r = t + u;
                if (r < ny - 1) {
                    double rhs = fma(r == 0 ? c_first : c_in, cur[u], a * (um + cur[u + 1]));
                    if (sh) rhs = fma(sc, sh[(size_t)r * N], rhs);
                    const double m = r < T ? piv[(size_t)r * N] : mf;
                    yp = fma(a, yp, rhs) * m;
                    col[(size_t)r * N] = yp;
                    um = cur[u];
                }
            }
            cur[0] = cur[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) cur[u + 1] = nxt[u];
        }
        // last row: its u^ is the entry of the window that follows the last processed row
        // (after the loop t >= ny - 1 and cur[0] holds row t; row ny - 1 is cur[ny - 1 - t + ... ] only when t == ny - 1)
    }
    double ul = col[(size_t)(ny - 1) * N];
    double rhs = fma(c_last, ul, a * um);
    if (sh) rhs = fma(sc, sh[(size_t)(ny - 1) * N], rhs);
    double xn = fma(a, yp, rhs) * m_last;
    col[(size_t)(ny - 1) * N] = xn;
    const double gf = a * mf;
    {
        double cur[PF], nxt[PF];
        int t = ny - 2;
#pragma unroll
        for (int u = 0; u < PF; ++u) cur[u] = t - u >= 0 ? col[(size_t)(t - u) * N] : 0.0;
        for (; t >= 0; t -= PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) nxt[u] = t - PF - u >= 0 ? col[(size_t)(t - PF - u) * N] : 0.0;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int r = t - u;
                if (r >= 0) {
                    xn = fma(r < T - 1 ? a * piv[(size_t)r * N] : gf, xn, cur[u]);
                    col[(size_t)r * N] = xn;
                }
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) cur[u] = nxt[u];
        }
    }
}

}  // namespace

static int configure_dct() {
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(k_dct_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        QPB_CUDA(cudaFuncSetAttribute(k_dct_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        configured = true;
    }
    return QPB_OK;
}

// Decide whether the prepared solve can take the spectral path and build its tables.
int qpbk_prepare_spectral(qpb_ctx *c, DiffSlot &s) {
    const auto &cf = c->cfg;
    s.spectral = false;
    if (getenv("QPB_NO_SPECTRAL") && getenv("QPB_NO_SPECTRAL")[0] == '1') return QPB_OK;
    if ((cf.flags & QPB_F_VARIABLE_D) || s.mode != 0 || cf.ncell != c->ncd) return QPB_OK;
    const int nx = cf.nx, ny = cf.ny;
    int logN = 0;
    while ((1 << logN) < nx) ++logN;
    if ((1 << logN) != nx || nx < 64 || nx > 4096 || ny < 2) return QPB_OK;
    std::vector<double> bcy_row(ny);
    for (int y = 0; y < ny; ++y) {
        bcy_row[y] = c->h_bcy[(size_t)y * nx];
        for (int x = 0; x < nx; ++x) {
            if (c->h_bcx[(size_t)y * nx + x] != 0.0) return QPB_OK;          // a wall term on the left / right
            if (c->h_bcy[(size_t)y * nx + x] != bcy_row[y]) return QPB_OK;    // wall kinds that change along the wall
        }
        if (bcy_row[y] < 0.0) return QPB_OK;
    }
    const double pi = 3.14159265358979323846;
    std::vector<double2> tw(nx, make_double2(1.0, 0.0)), tw2(nx);
    std::vector<double> lam(nx);
    for (int half = 1; half < nx; half <<= 1)   // packed by stage: tw[half + pos] = exp(-2 pi i pos / (2 half))
        for (int pos = 0; pos < half; ++pos)
            tw[half + pos] = make_double2(std::cos(pi * pos / half), -std::sin(pi * pos / half));
    for (int k = 0; k < nx; ++k) {
        tw2[k] = make_double2(std::cos(pi * k / (2.0 * nx)), -std::sin(pi * k / (2.0 * nx)));
        const double sn = std::sin(pi * k / (2.0 * nx));
        lam[k] = 4.0 * sn * sn;
    }
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_sp_tw, sizeof(double2) * tw.size()));
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_sp_tw2, sizeof(double2) * tw2.size()));
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_sp_lam, sizeof(double) * nx));
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_sp_bcy, sizeof(double) * ny));
    QPB_CUDA(cudaMemcpy(s.d_sp_tw, tw.data(), sizeof(double2) * tw.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(s.d_sp_tw2, tw2.data(), sizeof(double2) * tw2.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(s.d_sp_lam, lam.data(), sizeof(double) * nx, cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(s.d_sp_bcy, bcy_row.data(), sizeof(double) * ny, cudaMemcpyHostToDevice));
    s.sp_logn = logN;
    // Tabulated pivots of the Thomas pass (k_thomas_pivots): possible when only the first and last row carry a wall
    // term.  T rows until the slowest pivot recurrence (mode 0 of the bin with the largest a) is within 1e-18 of its
    // fixed point; past 128 rows (very stiff steps) or on short grids the pass keeps its on-the-fly form.
    s.sp_T = 0;
    bool interior_plain = true;
    for (int y = 1; y < ny - 1; ++y) interior_plain = interior_plain && bcy_row[y] == 0.0;
    if (interior_plain && !(getenv("QPB_NO_PIVOT_TABLE") && getenv("QPB_NO_PIVOT_TABLE")[0] == '1')) {
        int T = 4;
        for (int i = 0; i < cf.ne; ++i) {
            const double a = s.a_bin[i], d = 1.0 + 2.0 * a;
            const double ms = (d - std::sqrt(d * d - 4.0 * a * a)) / (2.0 * a * a);
            const double ratio = (a * ms) * (a * ms);
            const int need = ratio > 0.0 ? (int)std::ceil(std::log(1e-18) / std::log(ratio)) + 3 : 4;
            T = std::max(T, need);
        }
        if (T <= 128 && T <= ny - 2) {
            QPB_CUDA(qpb_dev_malloc((void **)&s.d_sp_piv, sizeof(double) * (size_t)cf.ne * T * nx));
            const dim3 grid((unsigned)((nx + 127) / 128), (unsigned)cf.ne);
            k_thomas_pivots<<<grid, 128, 0, c->stream>>>(T, nx, s.d_sp_piv, s.d_a, s.d_sp_lam, bcy_row[0]);
            QPB_CHECK_LAUNCH();
            QPB_CUDA(cudaStreamSynchronize(c->stream));
            s.sp_T = T;
            s.sp_wall_last = bcy_row[ny - 1];
            s.sp_wall0 = bcy_row[0];
            // with tabulated pivots the Thomas pass also forms the right-hand side (k_thomas_fused); it needs the
            // transform of the boundary sources when there are any
            if (!(getenv("QPB_NO_SPECTRAL_FUSED") && getenv("QPB_NO_SPECTRAL_FUSED")[0] == '1')) {
                bool any_src = false;
                for (size_t p = 0; p < c->h_src.size() && !any_src; ++p) any_src = c->h_src[p] != 0.0;
                if (any_src) {
                    int rc = configure_dct();
                    if (rc != QPB_OK) return rc;
                    QPB_CUDA(qpb_dev_malloc((void **)&s.d_sp_srchat, sizeof(double) * (size_t)ny * nx));
                    k_dct_forward<<<dim3((unsigned)((ny + 1) / 2), 1), nx >= 1024 ? 256 : 128,
                                    sizeof(double2) * (size_t)nx, c->stream>>>(
                        ny, nx, logN, c->d_srcgeom, s.d_sp_srchat, (const double2 *)s.d_sp_tw, (const double2 *)s.d_sp_tw2);
                    QPB_CHECK_LAUNCH();
                    QPB_CUDA(cudaStreamSynchronize(c->stream));
                }
                s.sp_fused = true;
            }
        }
    }
    s.spectral = true;
    return QPB_OK;
}

// u' = A^-1 b for every bin: b in c->d_B (already built), result into c->d_S.
int qpbk_diffuse_spectral(qpb_ctx *c, DiffSlot &s) {
    const auto &cf = c->cfg;
    const int ne = cf.ne, ny = cf.ny, nx = cf.nx;
    const size_t smem = sizeof(double2) * (size_t)nx;
    {
        int rc = configure_dct();
        if (rc != QPB_OK) return rc;
    }
    const dim3 rgrid((unsigned)((ny + 1) / 2), (unsigned)ne);
    const int fthreads = nx >= 1024 ? 256 : 128;
    {
        ScopedTimer tm(c, 0);
        k_dct_forward<<<rgrid, fthreads, smem, c->stream>>>(ny, nx, s.sp_logn, s.sp_fused ? c->d_S : c->d_B, c->d_T1,
                                                            (const double2 *)s.d_sp_tw, (const double2 *)s.d_sp_tw2);
        c->diag.kernel_launches++;
    }
    {
        ScopedTimer tm(c, 1);
        const dim3 tgrid((unsigned)((nx + 127) / 128), (unsigned)ne);
        if (s.sp_fused)
            k_thomas_fused<<<tgrid, 128, 0, c->stream>>>(ny, nx, s.sp_T, c->d_T1, s.d_sp_piv, s.d_a, s.d_sp_lam, s.sp_wall0,
                                                         s.sp_wall_last, s.d_sp_srchat, s.d_src);
        else if (s.sp_T > 0)
            k_thomas_frozen<<<tgrid, 128, 0, c->stream>>>(ny, nx, s.sp_T, c->d_T1, s.d_sp_piv, s.d_a, s.d_sp_lam,
                                                          s.sp_wall_last);
        else
            k_thomas_modes<<<tgrid, 128, 0, c->stream>>>(ne, ny, nx, c->d_T1, c->d_T2, s.d_a, s.d_sp_lam, s.d_sp_bcy);
        c->diag.kernel_launches++;
    }
    {
        ScopedTimer tm(c, 0);
        k_dct_inverse<<<rgrid, fthreads, smem, c->stream>>>(ny, nx, s.sp_logn, c->d_T1, c->d_S,
                                                            (const double2 *)s.d_sp_tw, (const double2 *)s.d_sp_tw2);
        c->diag.kernel_launches++;
    }
    QPB_CHECK_LAUNCH();
    // three passes over every bin, each reading and writing it once: counted as three directional sweeps
    c->diag.sweeps += 3;
    c->diag.bin_sweeps += 3LL * ne;
    return QPB_OK;
}
