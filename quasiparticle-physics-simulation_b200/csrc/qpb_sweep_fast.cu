// Chunked, table-driven tridiagonal sweeps for uniform D (the hot diffusion kernels).
//
// Same linear algebra as qpb_diffusion.cu (one Peaceman-Rachford half step per launch, or the direct solve of a
// one-cell-thick geometry), organised for the B200 memory system:
//
//  * A line (row for the x sweep, column for the y sweep) is cut into chunks of S cells; one thread owns one
//    chunk in registers.  The LU pivots of a line depend only on (bin, shift, geometry of the line), never on
//    the data, so they are factored once per prepared step length into a table that is shared by all lines
//    with the same geometry ("class"): the sweeps contain no division.
//  * With the pivots known, forward elimination  y_k = m_k d_k + f_k y_{k-1}  and back substitution
//    x_k = y_k + g_k x_{k+1}  are first-order linear recurrences.  Every chunk evaluates its recurrence with a
//    zero carry, the carries are resolved across the chunks of a line by a scan over affine maps (warp shuffles
//    for the x sweep where a line's chunks sit in adjacent lanes, shared memory for the y sweep where they sit
//    in different warps), and the chunk is re-evaluated with the true carry.
//  * x sweep: the right-hand side b - (V - rho)u (with the y stencil) and the residual b - A u are computed while
//    the tile is staged into shared memory with fully coalesced loads; chunks are read back with conflict-free
//    128-bit loads through an XOR swizzle.  y sweep: lanes run along x, so global accesses are coalesced directly.
//
// Algorithmic traffic: 16 B per cell*bin per sweep (read u, write u'); this implementation also reads b (x sweep)
// or u* (y sweep): 24 B, plus the pivot table (8 B per cell of a CLASS, shared by all its lines, L2 resident).
#include "qpb_internal.h"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

#define QPB_BCXNZ 32u  // flags: the cell has a non-zero x boundary diagonal
#define QPB_BCYNZ 64u

// ---- pivot factorisation ---------------------------------------------------------------------------------
// One thread per (bin, shift index, class).  geometry tables per class: lk[pos] (bit 0 = linked to pos-1, bit 1 = cell
// inside the mask), bc[pos].
// m_k = 1 / (sigma + rho + e_k + e_{k+1} + a*bc_k - e_k^2 m_{k-1}),  e_k = a*lk_k;  m_k = 0 outside the mask.
// g_k = e_{k+1} m_k (the multiplier of the scaled recurrences in qpb_sweep_pipe.cu), same layout, optional.
// interleave > 0 stores position k at ((k%S)/2 * Q + k/S)*2 + k%2 (16-byte units of a chunk side by side across
// chunks, for coalesced 128-bit loads by lanes that own adjacent chunks).
__global__ void k_factor(int ne, int jmax, int nclass, int npad, int S, int interleave, double sigma,
                         const double *__restrict__ a_bin, const double *__restrict__ shift,
                         const int *__restrict__ jlen, const uint8_t *__restrict__ lk, const double *__restrict__ bc,
                         double *__restrict__ tab, double *__restrict__ tabg, unsigned long long *__restrict__ amax) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)ne * jmax * nclass;
    if (gid >= total) return;
    const int cls = (int)(gid % nclass);
    const int j = (int)((gid / nclass) % jmax);
    const int bin = (int)(gid / ((long long)nclass * jmax));
    if (j >= jlen[bin]) return;
    const double a = a_bin[bin];
    const double rho = shift[(long long)bin * jmax + j];
    const uint8_t *l = lk + (size_t)cls * (npad + 1);
    const double *g = bc + (size_t)cls * npad;
    double *out = tab + (size_t)gid * npad;
    double *outg = tabg ? tabg + (size_t)gid * npad : nullptr;
    const int Q = npad / S;
    double mprev = 0.0, prod = 1.0, pmax = 0.0;
    for (int k = 0; k < npad; ++k) {
        const double e = (l[k] & 1) ? a : 0.0, en = (l[k + 1] & 1) ? a : 0.0;
        const double m = (l[k] & 2) ? 1.0 / (sigma + rho + e + en + a * g[k] - e * e * mprev) : 0.0;
        mprev = m;
        const int pos = interleave ? (((k % S) / 2) * Q + k / S) * 2 + (k & 1) : k;
        out[pos] = m;
        if (outg) outg[pos] = en * m;
        prod *= en * m;                       // coupling of a chunk to its neighbour: product of its g
        if (k % S == S - 1) {
            pmax = fmax(pmax, prod);
            prod = 1.0;
        }
    }
    if (amax) atomicMax(amax, (unsigned long long)__double_as_longlong(pmax));
}

// ---- affine-map scans ------------------------------------------------------------------------------------
// carry_out = A * carry_in + B for every chunk; returns the carry entering this chunk (0 for the first).
// Chunks of one line occupy WIDTH adjacent lanes (WIDTH a power of two <= 32).
template <int WIDTH, bool REVERSE>
__device__ __forceinline__ double warp_carry(double A, double B, int q) {
    // inclusive Kogge-Stone scan of affine maps in chunk order (or reverse chunk order)
#pragma unroll
    for (int off = 1; off < WIDTH; off <<= 1) {
        const double Ao = REVERSE ? __shfl_down_sync(0xffffffffu, A, off, WIDTH) : __shfl_up_sync(0xffffffffu, A, off, WIDTH);
        const double Bo = REVERSE ? __shfl_down_sync(0xffffffffu, B, off, WIDTH) : __shfl_up_sync(0xffffffffu, B, off, WIDTH);
        const bool has = REVERSE ? (q + off < WIDTH) : (q >= off);
        if (has) {
            B = fma(A, Bo, B);  // apply the earlier map first: y -> A*(Ao*y + Bo) + B
            A = A * Ao;
        }
    }
    const double prev = REVERSE ? __shfl_down_sync(0xffffffffu, B, 1, WIDTH) : __shfl_up_sync(0xffffffffu, B, 1, WIDTH);
    const bool first = REVERSE ? (q == WIDTH - 1) : (q == 0);
    return first ? 0.0 : prev;
}

struct SweepArgs {
    int ne, ny, nx, iter, jmax;
    const double *tol;   // [ne]
    double *S;          // state u (dense)
    const double *B;    // rhs b
    double *T1;         // u*
    const uint8_t *flags;
    const double *bcx, *bcy;
    const double *a_bin, *shift;
    const int *jlen;
    const int *cls;     // class of every line
    const uint8_t *lk;  // [nclass][npad+1]
    const double *tab;  // [ne][jmax][nclass][npad]
    int nclass, npad, Q;
    unsigned long long *res, *unorm;
    int *done, *iters_out;
    // per-cell coefficient fields of a non-uniform gap (dense [ne][ncd]): links to the left / upper neighbour, boundary
    // diagonals; tab then holds one line of pivots per grid line (x: [ne][jmax][ny][npad] chunk-interleaved,
    // y: [ne][jmax][npad][nx], position major so that lanes along x read neighbours)
    const double *ex, *ey, *gbx, *gby;
};

// per-thread solve of one chunk given its rhs d[S] (in place -> x), pivots m[S], link bits
template <int S, bool VARD = false>
struct Chunk {
    double v[S];
    double m[S];
    unsigned lkbits;   // bit t: element t linked to element t-1 (bit 0: to the previous chunk), bit S: next chunk
    unsigned lknext;
    double a;
    double ev[VARD ? S + 1 : 1];   // per-cell D: coupling of element t to element t-1 (0 where there is no link)

    __device__ __forceinline__ double e(int t) const {
        if (VARD) return ev[t];
        return ((t < 32 ? (lkbits >> t) : lknext) & 1u) ? a : 0.0;
    }

    // forward with zero carry: returns (A = prod f, B = y_last)
    __device__ __forceinline__ void fwd_probe(double &A, double &B) const {
        double y = 0.0, P = 1.0;
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const double f = e(t) * m[t];
            y = fma(f, y, v[t] * m[t]);
            P *= f;
        }
        A = P;
        B = y;
    }
    __device__ __forceinline__ void fwd_apply(double yin) {
        double y = yin;
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const double f = e(t) * m[t];
            y = fma(f, y, v[t] * m[t]);
            v[t] = y;
        }
    }
    __device__ __forceinline__ void bwd_probe(double &A, double &B) const {
        double x = 0.0, P = 1.0;
#pragma unroll
        for (int t = S - 1; t >= 0; --t) {
            const double g = e(t + 1) * m[t];
            x = fma(g, x, v[t]);
            P *= g;
        }
        A = P;
        B = x;
    }
    __device__ __forceinline__ void bwd_apply(double xin) {
        double x = xin;
#pragma unroll
        for (int t = S - 1; t >= 0; --t) {
            const double g = e(t + 1) * m[t];
            x = fma(g, x, v[t]);
            v[t] = x;
        }
    }
};

__device__ __forceinline__ bool bin_active(const SweepArgs &A, int bin, int mode, bool leader) {
    if (mode == 2) return true;
    if (A.done[bin]) return false;
    if (mode == 1) {
        const double r = __longlong_as_double((long long)A.res[(long long)A.iter * A.ne + bin]);
        const double un = __longlong_as_double((long long)A.unorm[(long long)A.iter * A.ne + bin]);
        (void)un;
        if (r <= 0.0) {   // no cell exceeded its componentwise bound (recorded by the x sweep)
            if (leader) {
                A.done[bin] = 1;
                A.iters_out[bin] = A.iter;
            }
            return false;
        }
    }
    return true;
}

// ---- x sweep ---------------------------------------------------------------------------------------------
// CTA = LINES rows x QP chunks (QP = chunks per row padded to a power of two <= 32), one thread per chunk.
// MODE 0: PR x half step (rhs with y stencil, residual);  MODE 2: direct solve of rows (rhs = b).
template <int S, int QP, int MODE>
__global__ void __launch_bounds__(256) k_sweep_x(SweepArgs A) {
    constexpr int THREADS = 256;
    constexpr int LINES = THREADS / QP;
    constexpr int ROWLEN = QP * S;  // padded row length in shared memory
    extern __shared__ double sm[];  // [LINES][ROWLEN]
    const int tid = threadIdx.x;
    const int tiles = (A.ny + LINES - 1) / LINES;
    const int bin = blockIdx.x / tiles;
    const int y0 = (blockIdx.x - bin * tiles) * LINES;
    if (!bin_active(A, bin, MODE, false)) return;
    const int nx = A.nx, ncd = A.ny * A.nx;
    const double a = A.a_bin[bin];
    const double rho = MODE == 2 ? 0.0 : A.shift[(long long)bin * A.jmax + (A.iter % A.jlen[bin])];
    const double *u = A.S + (long long)bin * ncd;
    const double *b = A.B + (long long)bin * ncd;

    // ---- stage the right-hand side, coalesced ----
    double rmax = 0.0, umax = 0.0;
    for (int e = tid; e < LINES * ROWLEN; e += THREADS) {
        const int g = e / ROWLEN, x = e - g * ROWLEN;
        const int y = y0 + g;
        double d = 0.0;
        if (y < A.ny && x < nx) {
            const int c = y * nx + x;
            const unsigned fl = A.flags[c];
            if (fl & QPB_IN) {
                const double bc_ = b[c];
                if (MODE == 0) {
                    const double uc = u[c];
                    const double au = fabs(uc);
                    double cross = (fl & QPB_BCYNZ) ? A.bcy[c] * uc : 0.0;
                    double wsum = ((fl & QPB_BCYNZ) ? fabs(A.bcy[c]) * au : 0.0) + ((fl & QPB_BCXNZ) ? fabs(A.bcx[c]) * au : 0.0);
                    if (fl & QPB_LK_U) { cross += uc - u[c - nx]; wsum += au + fabs(u[c - nx]); }
                    if (fl & QPB_LK_D) { cross += uc - u[c + nx]; wsum += au + fabs(u[c + nx]); }
                    double along = (fl & QPB_BCXNZ) ? A.bcx[c] * uc : 0.0;
                    if (fl & QPB_LK_L) { along += uc - u[c - 1]; wsum += au + fabs(u[c - 1]); }
                    if (fl & QPB_LK_R) { along += uc - u[c + 1]; wsum += au + fabs(u[c + 1]); }
                    d = fma(rho - 0.5, uc, bc_) - a * cross;
                    // componentwise stop test: excess of |b - A u| over tol (|A||u| + |b|) in this cell
                    rmax = fmax(rmax, fabs(bc_ - uc - a * (cross + along)) - A.tol[bin] * (fabs(bc_) + au + a * wsum));
                    umax = fmax(umax, au);
                } else {
                    d = bc_;
                }
            }
        }
        const int q = x / S, w = x - q * S;
        const int unit = (w >> 1) ^ (q & 7);
        sm[g * ROWLEN + q * S + unit * 2 + (w & 1)] = d;
    }
    if (MODE == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
            umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        }
        if ((tid & 31) == 0) {
            atomicMax(&A.res[(long long)A.iter * A.ne + bin], (unsigned long long)__double_as_longlong(rmax));
            atomicMax(&A.unorm[(long long)A.iter * A.ne + bin], (unsigned long long)__double_as_longlong(umax));
        }
    }
    __syncthreads();

    // ---- one chunk per thread ----
    const int g = tid / QP, q = tid - g * QP;
    const int y = min(y0 + g, A.ny - 1);
    const int cls = A.cls[y];
    Chunk<S> ch;
    ch.a = a;
    {
        const double2 *src = reinterpret_cast<const double2 *>(sm + g * ROWLEN + q * S);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) {
            const double2 t = src[un ^ (q & 7)];
            ch.v[2 * un] = t.x;
            ch.v[2 * un + 1] = t.y;
        }
        const double2 *mt = reinterpret_cast<const double2 *>(
            A.tab + ((((size_t)bin * A.jmax + (MODE == 2 ? 0 : A.iter % A.jlen[bin])) * A.nclass + cls) * A.npad));
        const int Qt = A.Q;  // chunks per line in the table (== QP when the row is exactly padded)
        const int qq = min(q, Qt - 1);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) {
            const double2 t = mt[un * Qt + qq];
            ch.m[2 * un] = t.x;
            ch.m[2 * un + 1] = t.y;
        }
        const uint8_t *lkp = A.lk + (size_t)cls * (A.npad + 1) + (size_t)qq * S;
        unsigned bits = 0;
#pragma unroll
        for (int t = 0; t < S && t < 32; ++t) bits |= (unsigned)(lkp[t] & 1u) << t;
        ch.lkbits = q < Qt ? bits : 0u;
        ch.lknext = q < Qt ? (lkp[S] & 1u) : 0u;
        if constexpr (S < 32) ch.lkbits |= ch.lknext << S;
        if (q >= Qt) {
#pragma unroll
            for (int t = 0; t < S; ++t) ch.m[t] = 1.0;
        }
    }
    double Am, Bm;
    ch.fwd_probe(Am, Bm);
    const double yin = warp_carry<QP, false>(Am, Bm, q);
    ch.fwd_apply(yin);
    ch.bwd_probe(Am, Bm);
    const double xin = warp_carry<QP, true>(Am, Bm, q);
    ch.bwd_apply(xin);
    __syncthreads();  // everyone has read its chunk of d
    {
        double2 *dst = reinterpret_cast<double2 *>(sm + g * ROWLEN + q * S);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) dst[un ^ (q & 7)] = make_double2(ch.v[2 * un], ch.v[2 * un + 1]);
    }
    __syncthreads();
    // ---- drain, coalesced ----
    double *out = (MODE == 0 ? A.T1 : A.S) + (long long)bin * ncd;
    for (int e = tid; e < LINES * ROWLEN; e += THREADS) {
        const int gg = e / ROWLEN, x = e - gg * ROWLEN;
        const int yy = y0 + gg;
        if (yy < A.ny && x < nx) {
            const int qx = x / S, w = x - qx * S;
            const int unit = (w >> 1) ^ (qx & 7);
            out[yy * nx + x] = sm[gg * ROWLEN + qx * S + unit * 2 + (w & 1)];
        }
    }
}

// ---- x sweep, TMA staged -----------------------------------------------------------------------------------
// Same solve as k_sweep_x, with the tiles moved by the tensor memory accelerator: the u tile (LINES rows plus one
// halo row above and below, zero filled outside the grid) and the b tile arrive in shared memory as 128-byte
// swizzled boxes of a 4-D view (16 doubles | nx/16 chunks | ny rows | ne bins) of the dense arrays, so the chunk
// owner reads everything it needs for its right-hand side with conflict-free 128-bit loads; the solved chunks go
// back through shared memory and one TMA store.  Requires nx % 16 == 0 and nx <= 512 (S = 16).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct TmaMaps {
    CUtensorMap u;    // state, box rows = LINES + 2
    CUtensorMap b;    // rhs b, box rows = LINES
    CUtensorMap out;  // T1 (mode 0) or state (mode 2), box rows = LINES
};

template <int QP, int MODE>
__global__ void __launch_bounds__(256, 2)
k_sweep_x_tma(SweepArgs A, const __grid_constant__ TmaMaps maps) {
    constexpr int S = 16;
    constexpr int THREADS = 256;
    constexpr int LINES = THREADS / QP;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int Q = A.Q;                       // real chunks per row (<= QP)
    double *su = reinterpret_cast<double *>(smraw);                                   // [(LINES+2)][Q][16]
    const int u_bytes = ((LINES + 2) * Q * 128 + 1023) / 1024 * 1024;
    double *sb = reinterpret_cast<double *>(smraw + (MODE == 0 ? u_bytes : 0));       // [LINES][Q][16], reused for x
    const int b_bytes = (LINES * Q * 128 + 1023) / 1024 * 1024;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw + (MODE == 0 ? u_bytes : 0) + b_bytes);
    const int tid = threadIdx.x;
    const int tiles = (A.ny + LINES - 1) / LINES;
    const int bin = blockIdx.x / tiles;
    const int y0 = (blockIdx.x - bin * tiles) * LINES;
    if (!bin_active(A, bin, MODE, false)) return;
    const uint32_t bar_a = smem_u32(bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (MODE == 0 ? (LINES + 2) * Q * 128 : 0) + LINES * Q * 128;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        if (MODE == 0) tma_load_4d(smem_u32(su), &maps.u, bar_a, 0, 0, y0 - 1, bin);
        tma_load_4d(smem_u32(sb), &maps.b, bar_a, 0, 0, y0, bin);
    }
    const int g = tid / QP, q = tid - g * QP;
    const int y = y0 + g;
    const bool rowok = y < A.ny && q < Q;
    const int yc = min(y, A.ny - 1), qc = min(q, Q - 1);
    const int nx = A.nx;
    const double a = A.a_bin[bin];
    const int jidx = MODE == 2 ? 0 : A.iter % A.jlen[bin];
    const double rho = MODE == 2 ? 0.0 : A.shift[(long long)bin * A.jmax + jidx];
    const int cls = A.cls[yc];
    // geometry of this chunk while the tiles are in flight
    const uint4 fl4 = *reinterpret_cast<const uint4 *>(A.flags + (size_t)yc * nx + qc * S);
    const unsigned flw[4] = {fl4.x, fl4.y, fl4.z, fl4.w};
    const double2 *mt = reinterpret_cast<const double2 *>(A.tab + ((((size_t)bin * A.jmax + jidx) * A.nclass + cls) * A.npad));
    mbar_wait(bar_a, 0);

    Chunk<S> ch;
    ch.a = a;
    double rmax = 0.0, umax = 0.0;
    const double tolb = A.tol[bin];
    {
        // chunk linear index inside a tile decides the swizzle phase: unit' = unit ^ ((row*Q + q) & 7)
        const int rb = g * Q + qc;            // b tile / output tile
        const double2 *pb = reinterpret_cast<const double2 *>(sb + (size_t)rb * S);
        const int swb = rb & 7;
        if (MODE == 0) {
            const int ru = (g + 1) * Q + qc;  // u tile has one halo row on top
            const double2 *pc = reinterpret_cast<const double2 *>(su + (size_t)ru * S);
            const double2 *pu = reinterpret_cast<const double2 *>(su + (size_t)(ru - Q) * S);
            const double2 *pd = reinterpret_cast<const double2 *>(su + (size_t)(ru + Q) * S);
            const int swc = ru & 7, swu = (ru - Q) & 7, swd = (ru + Q) & 7;
            double uc[S];
#pragma unroll
            for (int un = 0; un < S / 2; ++un) {
                const double2 t = pc[un ^ swc];
                uc[2 * un] = t.x;
                uc[2 * un + 1] = t.y;
            }
            // neighbours across the chunk boundary come from the adjacent lanes (same row: QP consecutive lanes)
            double ul = __shfl_up_sync(0xffffffffu, uc[S - 1], 1, QP);
            double ur = __shfl_down_sync(0xffffffffu, uc[0], 1, QP);
#pragma unroll
            for (int un = 0; un < S / 2; ++un) {
                const double2 tu = pu[un ^ swu], td = pd[un ^ swd], tb = pb[un ^ swb];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = 2 * un + h;
                    const unsigned fl = (flw[t >> 2] >> (8 * (t & 3))) & 0xffu;
                    const double u0 = uc[t];
                    const double uu = h == 0 ? tu.x : tu.y, ud = h == 0 ? td.x : td.y, bv = h == 0 ? tb.x : tb.y;
                    const double left = t == 0 ? ul : uc[t - 1], right = t == S - 1 ? ur : uc[t + 1];
                    const double au = fabs(u0);
                    double cross = 0.0, along = 0.0, wsum = 0.0;
                    if (fl & QPB_LK_U) { cross += u0 - uu; wsum += au + fabs(uu); }
                    if (fl & QPB_LK_D) { cross += u0 - ud; wsum += au + fabs(ud); }
                    if (fl & QPB_LK_L) { along += u0 - left; wsum += au + fabs(left); }
                    if (fl & QPB_LK_R) { along += u0 - right; wsum += au + fabs(right); }
                    if (fl & (QPB_BCXNZ | QPB_BCYNZ)) {
                        const size_t c = (size_t)yc * nx + qc * S + t;
                        if (fl & QPB_BCYNZ) { cross = fma(A.bcy[c], u0, cross); wsum = fma(fabs(A.bcy[c]), au, wsum); }
                        if (fl & QPB_BCXNZ) { along = fma(A.bcx[c], u0, along); wsum = fma(fabs(A.bcx[c]), au, wsum); }
                    }
                    const bool in = (fl & QPB_IN) && rowok;
                    const double d = fma(rho - 0.5, u0, bv) - a * cross;
                    ch.v[t] = in ? d : 0.0;
                    if (in) {   // componentwise stop test: excess of |b - A u| over tol (|A||u| + |b|) in this cell
                        rmax = fmax(rmax, fabs(bv - u0 - a * (cross + along)) - tolb * (fabs(bv) + au + a * wsum));
                        umax = fmax(umax, au);
                    }
                }
            }
        } else {
#pragma unroll
            for (int un = 0; un < S / 2; ++un) {
                const double2 tb = pb[un ^ swb];
                const unsigned f0 = (flw[(2 * un) >> 2] >> (8 * ((2 * un) & 3))) & 0xffu;
                const unsigned f1 = (flw[(2 * un + 1) >> 2] >> (8 * ((2 * un + 1) & 3))) & 0xffu;
                ch.v[2 * un] = ((f0 & QPB_IN) && rowok) ? tb.x : 0.0;
                ch.v[2 * un + 1] = ((f1 & QPB_IN) && rowok) ? tb.y : 0.0;
            }
        }
    }
    if (MODE == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
            umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        }
        if ((tid & 31) == 0) {
            atomicMax(&A.res[(long long)A.iter * A.ne + bin], (unsigned long long)__double_as_longlong(rmax));
            atomicMax(&A.unorm[(long long)A.iter * A.ne + bin], (unsigned long long)__double_as_longlong(umax));
        }
    }
    // pivots and links (links along x are bit 0 of the flag bytes; the link to the next chunk is LK_R of the last)
    {
#pragma unroll
        for (int un = 0; un < S / 2; ++un) {
            const double2 t = mt[un * Q + qc];
            ch.m[2 * un] = q < Q ? t.x : 1.0;
            ch.m[2 * un + 1] = q < Q ? t.y : 1.0;
        }
        unsigned bits = 0;
#pragma unroll
        for (int t = 0; t < S; ++t) bits |= (((flw[t >> 2] >> (8 * (t & 3))) & QPB_LK_L) ? 1u : 0u) << t;
        const unsigned last = (flw[3] >> 24) & 0xffu;
        ch.lknext = (last & QPB_LK_R) ? 1u : 0u;
        if (!rowok) { bits = 0; ch.lknext = 0; }
        ch.lkbits = bits | (ch.lknext << S);
    }
    double Am, Bm;
    ch.fwd_probe(Am, Bm);
    const double yin = warp_carry<QP, false>(Am, Bm, q);
    ch.fwd_apply(yin);
    ch.bwd_probe(Am, Bm);
    const double xin = warp_carry<QP, true>(Am, Bm, q);
    ch.bwd_apply(xin);
    __syncthreads();  // all reads of the b tile are done; it becomes the output tile
    if (q < Q) {
        const int rb = g * Q + q;
        double2 *dst = reinterpret_cast<double2 *>(sb + (size_t)rb * S);
        const int swb = rb & 7;
#pragma unroll
        for (int un = 0; un < S / 2; ++un) dst[un ^ swb] = make_double2(ch.v[2 * un], ch.v[2 * un + 1]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        tma_store_4d(&maps.out, smem_u32(sb), 0, 0, y0, bin);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---- y sweep ---------------------------------------------------------------------------------------------
// CTA = 32 columns (lanes) x QW chunks (warps).  MODE 1: PR y half step, rhs = u* - u, out = u + 2 rho x.
// MODE 2: direct solve of columns (rhs = b, out = x).
template <int S, int MODE>
__global__ void __launch_bounds__(512) k_sweep_y(SweepArgs A, int QW) {
    extern __shared__ double sm[];  // [2][QW][32] carries
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int tiles = (A.nx + 31) / 32;
    const int bin = blockIdx.x / tiles;
    const int x0 = (blockIdx.x - bin * tiles) * 32;
    if (!bin_active(A, bin, MODE, blockIdx.x == bin * tiles && threadIdx.x == 0)) return;
    const int nx = A.nx, ny = A.ny, ncd = ny * nx;
    const int x = x0 + lane;
    const bool live = x < nx;
    const int xc = live ? x : nx - 1;
    const double a = A.a_bin[bin];
    const int jidx = MODE == 2 ? 0 : A.iter % A.jlen[bin];
    const double rho = MODE == 2 ? 0.0 : A.shift[(long long)bin * A.jmax + jidx];
    double *u = A.S + (long long)bin * ncd;
    const double *p = (MODE == 1 ? A.T1 : A.B) + (long long)bin * ncd;
    const int cls = A.cls[xc];
    Chunk<S> ch;
    ch.a = a;
    const int r0 = q * S;
    double uold[MODE == 1 ? S : 1];
    {
        const double *mt = A.tab + (((size_t)bin * A.jmax + jidx) * A.nclass + cls) * A.npad + r0;
        const uint8_t *lkp = A.lk + (size_t)cls * (A.npad + 1) + r0;
        unsigned bits = 0;
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const int r = r0 + t;
            double d = 0.0;
            if (live && r < ny) {
                const int c = r * nx + x;
                if (MODE == 1) {
                    const double uc = u[c];
                    uold[t] = uc;
                    d = p[c] - uc;
                } else {
                    d = p[c];
                }
            } else if (MODE == 1) {
                uold[t] = 0.0;
            }
            ch.v[t] = d;
            ch.m[t] = mt[t];
            if (t < 32) bits |= (unsigned)(lkp[t] & 1u) << t;
        }
        ch.lkbits = bits;
        ch.lknext = lkp[S] & 1u;
        if constexpr (S < 32) ch.lkbits |= ch.lknext << S;
    }
    double *cA = sm, *cB = sm + QW * 32;
    double Am, Bm;
    ch.fwd_probe(Am, Bm);
    cA[q * 32 + lane] = Am;
    cB[q * 32 + lane] = Bm;
    __syncthreads();
    double carry = 0.0;
    for (int k = 0; k < q; ++k) carry = fma(cA[k * 32 + lane], carry, cB[k * 32 + lane]);
    ch.fwd_apply(carry);
    ch.bwd_probe(Am, Bm);
    __syncthreads();
    cA[q * 32 + lane] = Am;
    cB[q * 32 + lane] = Bm;
    __syncthreads();
    carry = 0.0;
    for (int k = QW - 1; k > q; --k) carry = fma(cA[k * 32 + lane], carry, cB[k * 32 + lane]);
    ch.bwd_apply(carry);
    if (live) {
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const int r = r0 + t;
            if (r < ny) {
                const int c = r * nx + x;
                u[c] = MODE == 1 ? fma(2.0 * rho, ch.v[t], uold[t]) : ch.v[t];
            }
        }
    }
}

// ---- per-cell D (non-uniform gap): the same chunked solves with one line of pivots per grid line ----------------------
// One thread per (bin, shift index, line).  e_k = coupling of cell k to cell k-1 along the line (ex / ey, 0 without a
// link), gb_k the boundary diagonal of the line's direction:  m_k = 1 / (1/2 + rho + e_k + e_{k+1} + gb_k - e_k^2 m_{k-1}).
__global__ void k_factor_vard(int ne, int jmax, int ny, int nx, int dir, int npad, int S, const double *__restrict__ shift,
                              const int *__restrict__ jlen, const uint8_t *__restrict__ flags, const double *__restrict__ ef,
                              const double *__restrict__ gb, double *__restrict__ tab) {
    const int nlines = dir == 0 ? ny : nx, n = dir == 0 ? nx : ny;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)ne * jmax * nlines;
    if (gid >= total) return;
    const int line = (int)(gid % nlines);
    const int j = (int)((gid / nlines) % jmax);
    const int bin = (int)(gid / ((long long)nlines * jmax));
    if (j >= jlen[bin]) return;
    const double rho = shift[(long long)bin * jmax + j];
    const long long off = (long long)bin * ny * nx;
    const int sk = dir == 0 ? 1 : nx, c0 = dir == 0 ? line * nx : line;
    const int Q = npad / S;
    double *out = dir == 0 ? tab + (size_t)gid * npad : tab + ((size_t)bin * jmax + j) * (size_t)npad * nx + line;
    double mprev = 0.0;
    for (int k = 0; k < npad; ++k) {
        double m = 0.0;
        if (k < n) {
            const int c = c0 + k * sk;
            if (flags[c] & QPB_IN) {
                const double e = ef[off + c];
                const double en = (k + 1 < n && (flags[c + sk] & QPB_IN)) ? ef[off + c + sk] : 0.0;
                m = 1.0 / (0.5 + rho + e + en + gb[off + c] - e * e * mprev);
            }
        }
        mprev = m;
        if (dir == 0) out[(((k % S) / 2) * Q + k / S) * 2 + (k & 1)] = m;
        else out[(size_t)k * nx] = m;
    }
}

// x sweep, per-cell D.  Same structure as k_sweep_x<S, QP, 0>: the right-hand side b - (V - rho) u and the componentwise
// stop test are formed while the tile is staged (coalesced), one thread then solves one chunk.
template <int S, int QP>
__global__ void __launch_bounds__(256) k_sweep_x_vard(SweepArgs A) {
    constexpr int THREADS = 256;
    constexpr int LINES = THREADS / QP;
    constexpr int ROWLEN = QP * S;
    extern __shared__ double sm[];  // [LINES][ROWLEN]
    const int tid = threadIdx.x;
    const int tiles = (A.ny + LINES - 1) / LINES;
    const int bin = blockIdx.x / tiles;
    const int y0 = (blockIdx.x - bin * tiles) * LINES;
    if (!bin_active(A, bin, 0, false)) return;
    const int nx = A.nx, ncd = A.ny * A.nx;
    const long long off = (long long)bin * ncd;
    const double rho = A.shift[(long long)bin * A.jmax + (A.iter % A.jlen[bin])];
    const double *u = A.S + off, *b = A.B + off;
    const double *ex = A.ex + off, *ey = A.ey + off, *gbx = A.gbx + off, *gby = A.gby + off;
    const double tol = A.tol[bin];

    double rmax = 0.0, umax = 0.0;
#pragma unroll 4
    for (int e = tid; e < LINES * ROWLEN; e += THREADS) {
        const int g = e / ROWLEN, x = e - g * ROWLEN;
        const int y = y0 + g;
        double d = 0.0;
        if (y < A.ny && x < nx) {
            const int c = y * nx + x;
            const unsigned fl = A.flags[c];
            // The face arrays hold 0 where there is no link, so the stencil needs no link tests: every load below is
            // independent of the flags (one memory round trip per cell instead of flags -> branch -> neighbour).
            const int cu = y > 0 ? c - nx : c, cd = y + 1 < A.ny ? c + nx : c, cl = x > 0 ? c - 1 : c, cr = x + 1 < nx ? c + 1 : c;
            const double bc_ = b[c], uc = u[c], uU = u[cu], uD = u[cd], uL = u[cl], uR = u[cr];
            const double eU = ey[c], eD = cd != c ? ey[cd] : 0.0, eL = ex[c], eR = cr != c ? ex[cr] : 0.0;
            if (fl & QPB_IN) {
                const double au = fabs(uc);
                // boundary diagonals are zero away from absorbing / Dirichlet / Robin walls: flagged, not streamed
                const double gx = (fl & QPB_BCXNZ) ? gbx[c] : 0.0, gy = (fl & QPB_BCYNZ) ? gby[c] : 0.0;
                double cross = gy * uc, along = gx * uc, wsum = (fabs(gy) + fabs(gx)) * au;
                cross = fma(eU, uc - uU, cross); wsum = fma(eU, au + fabs(uU), wsum);
                cross = fma(eD, uc - uD, cross); wsum = fma(eD, au + fabs(uD), wsum);
                along = fma(eL, uc - uL, along); wsum = fma(eL, au + fabs(uL), wsum);
                along = fma(eR, uc - uR, along); wsum = fma(eR, au + fabs(uR), wsum);
                d = fma(rho - 0.5, uc, bc_) - cross;
                rmax = fmax(rmax, fabs(bc_ - uc - cross - along) - tol * (fabs(bc_) + au + wsum));
                umax = fmax(umax, au);
            }
        }
        const int q = x / S, w = x - q * S;
        const int unit = (w >> 1) ^ (q & 7);
        sm[g * ROWLEN + q * S + unit * 2 + (w & 1)] = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
    }
    if ((tid & 31) == 0) {
        atomicMax(&A.res[(long long)A.iter * A.ne + bin], (unsigned long long)__double_as_longlong(rmax));
        atomicMax(&A.unorm[(long long)A.iter * A.ne + bin], (unsigned long long)__double_as_longlong(umax));
    }
    __syncthreads();

    const int g = tid / QP, q = tid - g * QP;
    const int y = min(y0 + g, A.ny - 1);
    Chunk<S, true> ch;
    ch.a = 0.0;
    ch.lkbits = ch.lknext = 0u;
    const int Qt = A.Q;
    const bool act = q < Qt;
    {
        const double2 *src = reinterpret_cast<const double2 *>(sm + g * ROWLEN + q * S);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) {
            const double2 t = src[un ^ (q & 7)];
            ch.v[2 * un] = t.x;
            ch.v[2 * un + 1] = t.y;
        }
        const double2 *mt = reinterpret_cast<const double2 *>(
            A.tab + ((((size_t)bin * A.jmax + (A.iter % A.jlen[bin])) * A.ny + y) * A.npad));
        const int qq = min(q, Qt - 1);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) {
            const double2 t = mt[un * Qt + qq];
            ch.m[2 * un] = act ? t.x : 1.0;
            ch.m[2 * un + 1] = act ? t.y : 1.0;
        }
        // couplings of the chunk's cells to their left neighbours (and of the next chunk's first cell to my last)
        const double *er = ex + (size_t)y * nx;
#pragma unroll
        for (int t = 0; t <= S; ++t) {
            const int x = q * S + t;
            ch.ev[t] = (act && x < nx) ? er[x] : 0.0;   // 0 where the cell has no left neighbour in the mask
        }
    }
    double Am, Bm;
    ch.fwd_probe(Am, Bm);
    const double yin = warp_carry<QP, false>(Am, Bm, q);
    ch.fwd_apply(yin);
    ch.bwd_probe(Am, Bm);
    const double xin = warp_carry<QP, true>(Am, Bm, q);
    ch.bwd_apply(xin);
    __syncthreads();
    {
        double2 *dst = reinterpret_cast<double2 *>(sm + g * ROWLEN + q * S);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) dst[un ^ (q & 7)] = make_double2(ch.v[2 * un], ch.v[2 * un + 1]);
    }
    __syncthreads();
    double *out = A.T1 + off;
    for (int e = tid; e < LINES * ROWLEN; e += THREADS) {
        const int gg = e / ROWLEN, x = e - gg * ROWLEN;
        const int yy = y0 + gg;
        if (yy < A.ny && x < nx) {
            const int qx = x / S, w = x - qx * S;
            const int unit = (w >> 1) ^ (qx & 7);
            out[yy * nx + x] = sm[gg * ROWLEN + qx * S + unit * 2 + (w & 1)];
        }
    }
}

// y sweep, per-cell D: CTA = 32 columns (lanes) x QW chunks (warps); rhs = u* - u, out = u + 2 rho x.
template <int S>
__global__ void __launch_bounds__(512) k_sweep_y_vard(SweepArgs A, int QW) {
    extern __shared__ double sm[];  // [2][QW][32] carries
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int tiles = (A.nx + 31) / 32;
    const int bin = blockIdx.x / tiles;
    const int x0 = (blockIdx.x - bin * tiles) * 32;
    if (!bin_active(A, bin, 1, blockIdx.x == bin * tiles && threadIdx.x == 0)) return;
    const int nx = A.nx, ny = A.ny, ncd = ny * nx;
    const int x = x0 + lane;
    const bool live = x < nx;
    const int xc = live ? x : nx - 1;
    const int jidx = A.iter % A.jlen[bin];
    const double rho = A.shift[(long long)bin * A.jmax + jidx];
    const long long off = (long long)bin * ncd;
    double *u = A.S + off;
    const double *p = A.T1 + off, *ey = A.ey + off;
    Chunk<S, true> ch;
    ch.a = 0.0;
    ch.lkbits = ch.lknext = 0u;
    const int r0 = q * S;
    double uold[S];
    {
        const double *mt = A.tab + (((size_t)bin * A.jmax + jidx) * A.npad + r0) * nx + xc;
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const int r = r0 + t;
            double d = 0.0;
            uold[t] = 0.0;
            ch.ev[t] = 0.0;
            if (live && r < ny) {
                const int c = r * nx + x;
                const double uc = u[c];
                uold[t] = uc;
                d = p[c] - uc;
                ch.ev[t] = ey[c];   // 0 where the cell has no upper neighbour in the mask
            }
            ch.v[t] = d;
            ch.m[t] = mt[(size_t)t * nx];
        }
        const int rn = r0 + S;
        ch.ev[S] = (live && rn < ny) ? ey[rn * nx + x] : 0.0;
    }
    double *cA = sm, *cB = sm + QW * 32;
    double Am, Bm;
    ch.fwd_probe(Am, Bm);
    cA[q * 32 + lane] = Am;
    cB[q * 32 + lane] = Bm;
    __syncthreads();
    double carry = 0.0;
    for (int k = 0; k < q; ++k) carry = fma(cA[k * 32 + lane], carry, cB[k * 32 + lane]);
    ch.fwd_apply(carry);
    ch.bwd_probe(Am, Bm);
    __syncthreads();
    cA[q * 32 + lane] = Am;
    cB[q * 32 + lane] = Bm;
    __syncthreads();
    carry = 0.0;
    for (int k = QW - 1; k > q; --k) carry = fma(cA[k * 32 + lane], carry, cB[k * 32 + lane]);
    ch.bwd_apply(carry);
    if (live) {
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const int r = r0 + t;
            if (r < ny) u[r * nx + x] = fma(2.0 * rho, ch.v[t], uold[t]);
        }
    }
}

// ---- host: classes and tables ------------------------------------------------------------------------------
struct ClassInfo {
    std::vector<int> cls;          // per line
    std::vector<uint8_t> lk;       // [nclass][npad+1]
    std::vector<double> bc;        // [nclass][npad]
    int nclass = 0;
};

// dir 0: lines are rows (along x); dir 1: lines are columns.  add_cross: direct mode puts the cross bc on the diagonal.
ClassInfo build_classes(const qpb_ctx *c, int dir, int npad, bool add_cross) {
    const int ny = c->cfg.ny, nx = c->cfg.nx;
    const int nlines = dir == 0 ? ny : nx, n = dir == 0 ? nx : ny;
    ClassInfo ci;
    ci.cls.resize(nlines);
    std::map<std::string, int> seen;
    std::vector<uint8_t> lk(npad + 1);
    std::vector<double> bc(npad);
    for (int l = 0; l < nlines; ++l) {
        std::fill(lk.begin(), lk.end(), 0);
        std::fill(bc.begin(), bc.end(), 0.0);
        for (int k = 0; k < n; ++k) {
            const int p = dir == 0 ? l * nx + k : k * nx + l;
            const unsigned f = c->h_flags[p];
            if (!(f & QPB_IN)) continue;
            lk[k] = ((f & (dir == 0 ? QPB_LK_L : QPB_LK_U)) ? 1 : 0) | 2;   // bit 1: inside the mask
            bc[k] = dir == 0 ? c->h_bcx[p] : c->h_bcy[p];
            if (add_cross) bc[k] += dir == 0 ? c->h_bcy[p] : c->h_bcx[p];
        }
        std::string key((const char *)lk.data(), lk.size());
        key.append((const char *)bc.data(), bc.size() * sizeof(double));
        auto it = seen.find(key);
        if (it == seen.end()) {
            it = seen.emplace(key, ci.nclass++).first;
            ci.lk.insert(ci.lk.end(), lk.begin(), lk.end());
            ci.bc.insert(ci.bc.end(), bc.begin(), bc.end());
        }
        ci.cls[l] = it->second;
    }
    return ci;
}

struct FastDev {
    uint8_t *d_lk = nullptr;
};

// x sweep: up to 32 chunks per row (one warp scans them); y sweep: up to 16 chunks per column (512 threads)
int pick_S(int n, int dir) { return dir == 0 ? (n <= 512 ? 16 : 32) : (n <= 256 ? 16 : 32); }

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

// The per-direction device geometry tables are stored behind d_cls: [cls ints | lk bytes], located by offsets
// kept in FastDir (n = line length, S, Q, npad, nclass).
static size_t cls_bytes(int nlines) { return ((sizeof(int) * (size_t)nlines + 255) / 256) * 256; }

static int setup_dir(qpb_ctx *c, DiffSlot &s, DiffSlot::FastDir &fd, int dir, bool direct, bool interleave,
                     bool pipe = false) {
    const auto &cf = c->cfg;
    const int n = dir == 0 ? cf.nx : cf.ny;
    const int nlines = dir == 0 ? cf.ny : cf.nx;
    fd.n = n;
    fd.S = pipe ? qpbp_chunk(n, dir) : pick_S(n, dir);
    fd.Q = (n + fd.S - 1) / fd.S;
    fd.npad = fd.Q * fd.S;
    ClassInfo ci = build_classes(c, dir, fd.npad, direct);
    fd.nclass = ci.nclass;
    const size_t tab_elems = (size_t)cf.ne * s.jmax * ci.nclass * fd.npad;
    if (tab_elems * sizeof(double) * (pipe ? 2 : 1) > ((size_t)6 << 30)) return 1;  // irregular geometry: not worth tabulating
    const size_t cb = cls_bytes(nlines);
    char *blob = nullptr;
    QPB_CUDA(qpb_dev_malloc((void **)&blob, cb + ci.lk.size()));
    fd.d_cls = (int *)blob;
    QPB_CUDA(cudaMemcpy(blob, ci.cls.data(), sizeof(int) * nlines, cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(blob + cb, ci.lk.data(), ci.lk.size(), cudaMemcpyHostToDevice));
    double *d_bc = nullptr;
    QPB_CUDA(qpb_dev_malloc((void **)&d_bc, sizeof(double) * ci.bc.size()));
    QPB_CUDA(cudaMemcpy(d_bc, ci.bc.data(), sizeof(double) * ci.bc.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(qpb_dev_malloc((void **)&fd.d_tab, sizeof(double) * tab_elems));
    if (pipe) QPB_CUDA(qpb_dev_malloc((void **)&fd.d_tabg, sizeof(double) * tab_elems));
    const long long total = (long long)cf.ne * s.jmax * ci.nclass;
    unsigned long long *d_amax = nullptr;
    QPB_CUDA(qpb_dev_malloc((void **)&d_amax, sizeof(unsigned long long)));
    QPB_CUDA(cudaMemsetAsync(d_amax, 0, sizeof(unsigned long long), c->stream));
    k_factor<<<(int)ceil_div64(total, 64), 64, 0, c->stream>>>(cf.ne, s.jmax, ci.nclass, fd.npad, fd.S, interleave ? 1 : 0,
                                                              direct ? 1.0 : 0.5, s.d_a, s.d_shift, s.d_jlen,
                                                              (const uint8_t *)(blob + cb), d_bc, fd.d_tab, fd.d_tabg, d_amax);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    unsigned long long bits = 0;
    QPB_CUDA(cudaMemcpyAsync(&bits, d_amax, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    qpb_dev_free(d_amax);
    {   // how many neighbouring chunks a carry can reach before it drops below one part in 1e18
        double am;
        memcpy(&am, &bits, sizeof(am));
        int depth = fd.Q;
        if (am <= 0.0) depth = 1;
        else if (am < 1.0) depth = (int)std::ceil(std::log(1e-18) / std::log(am));
        fd.carry_depth = std::max(1, std::min(depth, fd.Q));
    }
    qpb_dev_free(d_bc);
    return 0;
}

// ---- tensor maps --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 4-D view of a dense [ne][ny][nx] float64 array: (16 | nx/16 | ny | ne), box (16 | Q | rows | 1), 128-byte swizzle.
static bool make_map(CUtensorMap *m, double *base, int ne, int ny, int nx, int Q, int rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {16, (cuuint64_t)(nx / 16), (cuuint64_t)ny, (cuuint64_t)ne};
    const cuuint64_t strides[3] = {128, (cuuint64_t)nx * 8, (cuuint64_t)ny * nx * 8};
    const cuuint32_t box[4] = {16, (cuuint32_t)Q, (cuuint32_t)rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void setup_tma(qpb_ctx *c, DiffSlot &s) {
    DiffSlot::FastDir &fd = s.fx;
    const auto &cf = c->cfg;
    fd.use_tma = false;
    if (getenv("QPB_NO_TMA") && getenv("QPB_NO_TMA")[0] == '1') return;
    if (fd.S != 16 || cf.nx % 16 != 0 || cf.nx > 512 || fd.Q > 32) return;
    const int qp = next_pow2(fd.Q);
    const int lines = 256 / qp;
    if (lines + 2 > 256) return;
    TmaMaps maps;
    const bool direct = s.mode != 0;
    if (!make_map(&maps.u, c->d_S, cf.ne, cf.ny, cf.nx, fd.Q, lines + 2)) return;
    if (!make_map(&maps.b, c->d_B, cf.ne, cf.ny, cf.nx, fd.Q, lines)) return;
    if (!make_map(&maps.out, direct ? c->d_S : c->d_T1, cf.ne, cf.ny, cf.nx, fd.Q, lines)) return;
    fd.tma.resize(sizeof(TmaMaps));
    memcpy(fd.tma.data(), &maps, sizeof(TmaMaps));
    fd.use_tma = true;
}

// per-cell D: one line of pivots per grid line and shift (the chunked kernels k_sweep_x_vard / k_sweep_y_vard)
static int setup_dir_vard(qpb_ctx *c, DiffSlot &s, DiffSlot::FastDir &fd, int dir) {
    const auto &cf = c->cfg;
    fd.n = dir == 0 ? cf.nx : cf.ny;
    fd.S = 16;
    fd.Q = (fd.n + fd.S - 1) / fd.S;
    fd.npad = fd.Q * fd.S;
    fd.nclass = dir == 0 ? cf.ny : cf.nx;
    fd.carry_depth = fd.Q;
    fd.use_tma = false;
    const size_t elems = (size_t)cf.ne * s.jmax * (dir == 0 ? (size_t)cf.ny * fd.npad : (size_t)fd.npad * cf.nx);
    if (elems * sizeof(double) > ((size_t)4 << 30)) return 1;
    QPB_CUDA(qpb_dev_malloc((void **)&fd.d_tab, sizeof(double) * elems));
    const long long total = (long long)cf.ne * s.jmax * fd.nclass;
    k_factor_vard<<<(int)ceil_div64(total, 64), 64, 0, c->stream>>>(cf.ne, s.jmax, cf.ny, cf.nx, dir, fd.npad, fd.S, s.d_shift,
                                                                   s.d_jlen, c->d_flags, dir == 0 ? s.d_ex : s.d_ey,
                                                                   dir == 0 ? s.d_gbx : s.d_gby, fd.d_tab);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return 0;
}

int qpbk_prepare_fast(qpb_ctx *c, DiffSlot &s) {
    s.fast = false;
    const auto &cf = c->cfg;
    if (getenv("QPB_FORCE_GENERIC") && getenv("QPB_FORCE_GENERIC")[0] == '1') return QPB_OK;
    if (cf.flags & QPB_F_VARIABLE_D) {
        // non-uniform gap: chunked sweeps with per-line pivot tables where a line fits one CTA (rows up to 512 cells,
        // columns up to 256); one-cell-thick geometries and longer lines stay with the one-thread-per-line kernels
        if (s.mode != 0 || cf.nx > 512 || cf.ny > 256) return QPB_OK;
        if (getenv("QPB_NO_VARD_FAST") && getenv("QPB_NO_VARD_FAST")[0] == '1') return QPB_OK;
        {   // flag bits for non-zero boundary diagonals (read by the x sweep)
            std::vector<uint8_t> fl = c->h_flags;
            for (int p = 0; p < c->ncd; ++p) {
                if (c->h_bcx[p] != 0.0) fl[p] |= QPB_BCXNZ;
                if (c->h_bcy[p] != 0.0) fl[p] |= QPB_BCYNZ;
            }
            QPB_CUDA(cudaMemcpy(c->d_flags, fl.data(), c->ncd, cudaMemcpyHostToDevice));
        }
        int rc;
        if ((rc = setup_dir_vard(c, s, s.fx, 0)) != 0) return rc < 0 ? rc : QPB_OK;
        if ((rc = setup_dir_vard(c, s, s.fy, 1)) != 0) return rc < 0 ? rc : QPB_OK;
        s.pipe = PipePlan();
        s.fast = true;
        return QPB_OK;
    }
    // lines beyond the older chunked kernels: only the segmented pipelined kernels (iterated solves) take them
    const bool long_lines = cf.nx > 1024 || cf.ny > 512;
    if (long_lines && (s.mode != 0 || cf.nx % 16 != 0 || cf.ne > 2048 ||
                       (getenv("QPB_NO_PIPE") && getenv("QPB_NO_PIPE")[0] == '1')))
        return QPB_OK;
    // flag bits for non-zero boundary diagonals (read by the x sweep)
    static_assert(QPB_BCXNZ == 32u && QPB_BCYNZ == 64u, "flag bits");
    std::vector<uint8_t> fl = c->h_flags;
    for (int p = 0; p < c->ncd; ++p) {
        if (c->h_bcx[p] != 0.0) fl[p] |= QPB_BCXNZ;
        if (c->h_bcy[p] != 0.0) fl[p] |= QPB_BCYNZ;
    }
    QPB_CUDA(cudaMemcpy(c->d_flags, fl.data(), c->ncd, cudaMemcpyHostToDevice));
    {   // per-cell doubles of the select-free sweeps: linked neighbours + boundary diagonal, 0 outside the mask
        std::vector<double> cx(c->ncd, 0.0), cy(c->ncd, 0.0);
        for (int p = 0; p < c->ncd; ++p) {
            const unsigned f = c->h_flags[p];
            if (!(f & QPB_IN)) continue;
            cx[p] = ((f & QPB_LK_L) ? 1.0 : 0.0) + ((f & QPB_LK_R) ? 1.0 : 0.0) + c->h_bcx[p];
            cy[p] = ((f & QPB_LK_U) ? 1.0 : 0.0) + ((f & QPB_LK_D) ? 1.0 : 0.0) + c->h_bcy[p];
        }
        QPB_CUDA(cudaMemcpy(c->d_cx, cx.data(), sizeof(double) * c->ncd, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(c->d_cy, cy.data(), sizeof(double) * c->ncd, cudaMemcpyHostToDevice));
    }
    int rc;
    const bool no_pipe = getenv("QPB_NO_PIPE") && getenv("QPB_NO_PIPE")[0] == '1';
    if (s.mode == 0) {
        const bool px = !no_pipe && cf.nx % 16 == 0 && qpbp_chunk(cf.nx, 0) > 0 && cf.ne <= 2048;
        const bool py = !no_pipe && cf.nx % 2 == 0 && qpbp_chunk(cf.ny, 1) > 0 && cf.ne <= 2048;
        if ((rc = setup_dir(c, s, s.fx, 0, false, true, px)) < 0) return rc;
        if (rc > 0) return QPB_OK;
        if ((rc = setup_dir(c, s, s.fy, 1, false, false, py)) < 0) return rc;
        if (rc > 0) return QPB_OK;
    } else if (s.mode == 1) {
        if ((rc = setup_dir(c, s, s.fx, 0, true, true)) != 0) return rc < 0 ? rc : QPB_OK;
    } else {
        if ((rc = setup_dir(c, s, s.fy, 1, true, false)) != 0) return rc < 0 ? rc : QPB_OK;
    }
    if (s.mode != 2) setup_tma(c, s);
    s.fast = true;
    if ((rc = qpbp_plan(c, s, s.pipe)) != QPB_OK) return rc;
    // tables chunked for the pipelined kernels cannot be walked by the older kernels: generic sweeps instead
    if (s.mode == 0 && !s.pipe.x_ok && s.fx.d_tabg && (s.fx.S != pick_S(cf.nx, 0))) s.fast = false;
    if (s.mode == 0 && !s.pipe.y_ok && s.fy.d_tabg && (s.fy.S != pick_S(cf.ny, 1) || s.fy.Q > 16)) s.fast = false;
    if (long_lines && !(s.pipe.x_ok && s.pipe.y_ok)) s.fast = false;
    return QPB_OK;
}

template <int QP, int MODE>
static int launch_x_tma(qpb_ctx *c, const SweepArgs &A, const TmaMaps &maps) {
    constexpr int LINES = 256 / QP;
    const int tiles = (A.ny + LINES - 1) / LINES;
    const size_t ub = MODE == 0 ? ((size_t)(LINES + 2) * A.Q * 128 + 1023) / 1024 * 1024 : 0;
    const size_t bb = ((size_t)LINES * A.Q * 128 + 1023) / 1024 * 1024;
    const size_t smem = ub + bb + 64;
    auto kern = k_sweep_x_tma<QP, MODE>;
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    kern<<<A.ne * tiles, 256, smem, c->stream>>>(A, maps);
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

template <int MODE>
static int dispatch_x_tma(qpb_ctx *c, const SweepArgs &A, const TmaMaps &maps) {
    switch (next_pow2(A.Q)) {
        case 1: return launch_x_tma<1, MODE>(c, A, maps);
        case 2: return launch_x_tma<2, MODE>(c, A, maps);
        case 4: return launch_x_tma<4, MODE>(c, A, maps);
        case 8: return launch_x_tma<8, MODE>(c, A, maps);
        case 16: return launch_x_tma<16, MODE>(c, A, maps);
        default: return launch_x_tma<32, MODE>(c, A, maps);
    }
}

template <int S, int QP, int MODE>
static int launch_x(qpb_ctx *c, const SweepArgs &A) {
    constexpr int LINES = 256 / QP;
    const int tiles = (A.ny + LINES - 1) / LINES;
    const size_t smem = sizeof(double) * LINES * QP * S;
    auto kern = k_sweep_x<S, QP, MODE>;
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        configured = true;
    }
    kern<<<A.ne * tiles, 256, smem, c->stream>>>(A);
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

template <int S, int MODE>
static int dispatch_x(qpb_ctx *c, const SweepArgs &A) {
    const int qp = next_pow2(A.Q);
    switch (qp) {
        case 1: return launch_x<S, 1, MODE>(c, A);
        case 2: return launch_x<S, 2, MODE>(c, A);
        case 4: return launch_x<S, 4, MODE>(c, A);
        case 8: return launch_x<S, 8, MODE>(c, A);
        case 16: return launch_x<S, 16, MODE>(c, A);
        default: return launch_x<S, 32, MODE>(c, A);
    }
}

int qpbk_sweep_fast(qpb_ctx *c, DiffSlot &s, int dir, int iter, int mode, bool check) {
    if (mode == 0 && dir == 0 && s.pipe.x_ok) return qpbp_sweep(c, s, 0, iter, check);
    if (mode == 1 && dir == 1 && s.pipe.y_ok) return qpbp_sweep(c, s, 1, iter, check);
    const auto &cf = c->cfg;
    const DiffSlot::FastDir &fd = dir == 0 ? s.fx : s.fy;
    const int nlines = dir == 0 ? cf.ny : cf.nx;
    if (cf.flags & QPB_F_VARIABLE_D) {
        SweepArgs A{};
        A.ne = cf.ne; A.ny = cf.ny; A.nx = cf.nx; A.iter = iter; A.jmax = s.jmax; A.tol = s.d_tol;
        A.S = c->d_S; A.B = c->d_B; A.T1 = c->d_T1; A.flags = c->d_flags;
        A.shift = s.d_shift; A.jlen = s.d_jlen; A.tab = fd.d_tab; A.nclass = fd.nclass; A.npad = fd.npad; A.Q = fd.Q;
        A.res = c->d_res; A.unorm = c->d_unorm; A.done = c->d_done; A.iters_out = c->d_done + cf.ne;
        A.ex = s.d_ex; A.ey = s.d_ey; A.gbx = s.d_gbx; A.gby = s.d_gby;
        ScopedTimer tm(c, dir == 0 ? 0 : 1);
        c->diag.kernel_launches++;
        if (dir == 0) {
            const int qp = next_pow2(fd.Q);
            const int lines = 256 / qp, tiles = (cf.ny + lines - 1) / lines;
            const size_t smem = sizeof(double) * 256 * 16;
            switch (qp) {
                case 1: k_sweep_x_vard<16, 1><<<cf.ne * tiles, 256, smem, c->stream>>>(A); break;
                case 2: k_sweep_x_vard<16, 2><<<cf.ne * tiles, 256, smem, c->stream>>>(A); break;
                case 4: k_sweep_x_vard<16, 4><<<cf.ne * tiles, 256, smem, c->stream>>>(A); break;
                case 8: k_sweep_x_vard<16, 8><<<cf.ne * tiles, 256, smem, c->stream>>>(A); break;
                case 16: k_sweep_x_vard<16, 16><<<cf.ne * tiles, 256, smem, c->stream>>>(A); break;
                default: k_sweep_x_vard<16, 32><<<cf.ne * tiles, 256, smem, c->stream>>>(A); break;
            }
        } else {
            const int QW = fd.Q, tiles = (cf.nx + 31) / 32;
            k_sweep_y_vard<16><<<cf.ne * tiles, 32 * QW, sizeof(double) * 2 * QW * 32, c->stream>>>(A, QW);
        }
        QPB_CHECK_LAUNCH();
        return QPB_OK;
    }
    SweepArgs A;
    A.ne = cf.ne; A.ny = cf.ny; A.nx = cf.nx; A.iter = iter; A.jmax = s.jmax; A.tol = s.d_tol;
    A.S = c->d_S; A.B = c->d_B; A.T1 = c->d_T1; A.flags = c->d_flags; A.bcx = c->d_bcx; A.bcy = c->d_bcy;
    A.a_bin = s.d_a; A.shift = s.d_shift; A.jlen = s.d_jlen;
    A.cls = fd.d_cls;
    A.lk = (const uint8_t *)((const char *)fd.d_cls + cls_bytes(nlines));
    A.tab = fd.d_tab; A.nclass = fd.nclass; A.npad = fd.npad; A.Q = fd.Q;
    A.res = c->d_res; A.unorm = c->d_unorm; A.done = c->d_done; A.iters_out = c->d_done + cf.ne;
    ScopedTimer tm(c, dir == 0 ? 0 : 1);
    c->diag.kernel_launches++;
    if (dir == 0) {
        if (fd.use_tma) {
            const TmaMaps &maps = *reinterpret_cast<const TmaMaps *>(fd.tma.data());
            return mode == 0 ? dispatch_x_tma<0>(c, A, maps) : dispatch_x_tma<2>(c, A, maps);
        }
        if (fd.S == 16) return mode == 0 ? dispatch_x<16, 0>(c, A) : dispatch_x<16, 2>(c, A);
        return mode == 0 ? dispatch_x<32, 0>(c, A) : dispatch_x<32, 2>(c, A);
    }
    const int QW = fd.Q;
    const int tiles = (cf.nx + 31) / 32;
    const size_t smem = sizeof(double) * 2 * QW * 32;
    if (fd.S == 16) {
        if (mode == 1) k_sweep_y<16, 1><<<cf.ne * tiles, 32 * QW, smem, c->stream>>>(A, QW);
        else k_sweep_y<16, 2><<<cf.ne * tiles, 32 * QW, smem, c->stream>>>(A, QW);
    } else {
        if (mode == 1) k_sweep_y<32, 1><<<cf.ne * tiles, 32 * QW, smem, c->stream>>>(A, QW);
        else k_sweep_y<32, 2><<<cf.ne * tiles, 32 * QW, smem, c->stream>>>(A, QW);
    }
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}
