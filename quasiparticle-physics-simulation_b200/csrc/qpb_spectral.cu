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

// in-place radix-2 decimation-in-time passes over z[0..N) (input in bit-reversed order).  tw is packed by stage:
// tw[half + pos] = exp(-2 pi i pos / (2 half)) for pos < half (contiguous reads in every stage), conjugated for the
// inverse transform
template <bool INV>
__device__ __forceinline__ void fft_passes(double2 *z, const double2 *tw, int N, int logN) {
    for (int s = 0; s < logN; ++s) {
        const int half = 1 << s;
        __syncthreads();
        for (int q = threadIdx.x; q < N / 2; q += blockDim.x) {
            const int pos = q & (half - 1);
            const int i = ((q >> s) << (s + 1)) + pos;
            double2 w = tw[half + pos];
            if (INV) w.y = -w.y;
            const int ia = slot(i), ib = slot(i + half);
            const double2 a = z[ia], t = cmul(w, z[ib]);
            z[ia] = make_double2(a.x + t.x, a.y + t.y);
            z[ib] = make_double2(a.x - t.x, a.y - t.y);
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
    double2 *z = smz, *tw = smz + N;
    const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
    const bool two = r1 < ny;
    const size_t base = (size_t)blockIdx.y * ny * N;
    const double *a = in + base + (size_t)r0 * N, *b = in + base + (size_t)r1 * N;
    for (int k = threadIdx.x; k < N; k += blockDim.x) tw[k] = twg[k];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const int n = reorder(j, N);
        z[slot(__brev((unsigned)n) >> (32 - logN))] = make_double2(a[j], two ? b[j] : 0.0);
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
    double2 *z = smz, *tw = smz + N;
    const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
    const bool two = r1 < ny;
    const size_t base = (size_t)blockIdx.y * ny * N;
    const double *a = in + base + (size_t)r0 * N, *b = in + base + (size_t)r1 * N;
    for (int k = threadIdx.x; k < N; k += blockDim.x) tw[k] = twg[k];
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

}  // namespace

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
    s.spectral = true;
    return QPB_OK;
}

// u' = A^-1 b for every bin: b in c->d_B (already built), result into c->d_S.
int qpbk_diffuse_spectral(qpb_ctx *c, DiffSlot &s) {
    const auto &cf = c->cfg;
    const int ne = cf.ne, ny = cf.ny, nx = cf.nx;
    const size_t smem = sizeof(double2) * ((size_t)nx + nx);
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(k_dct_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        QPB_CUDA(cudaFuncSetAttribute(k_dct_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        configured = true;
    }
    const dim3 rgrid((unsigned)((ny + 1) / 2), (unsigned)ne);
    const int fthreads = nx >= 1024 ? 256 : 128;
    {
        ScopedTimer tm(c, 0);
        k_dct_forward<<<rgrid, fthreads, smem, c->stream>>>(ny, nx, s.sp_logn, c->d_B, c->d_T1,
                                                            (const double2 *)s.d_sp_tw, (const double2 *)s.d_sp_tw2);
        c->diag.kernel_launches++;
    }
    {
        ScopedTimer tm(c, 1);
        const dim3 tgrid((unsigned)((nx + 127) / 128), (unsigned)ne);
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
    c->diag.sweeps += 2;
    c->diag.bin_sweeps += 2LL * ne;
    return QPB_OK;
}
