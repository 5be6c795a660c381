// Bin-resident Crank-Nicolson solve: one thread-block cluster keeps an energy bin on chip for the whole iteration.
//
// Same linear algebra as the sweep kernels (qpsim/solver.py:1428-1452: rhs = B u + dt D s ; u' = A^-1 rhs, solved here
// by the Peaceman-Rachford iteration in residual-correction form
//     res = b - A u ;  (H + r) d* = res  (rows) ;  (V + r) d = d*  (columns) ;  u <- u + 2 r d
// with H = I/2 + a Gx, V = I/2 + a Gy, A = H + V), but the per-launch sweeps move 24 B per cell, bin and sweep through
// HBM, about 36 times per step on a mask.  Here a cluster of up to 8 CTAs owns one bin at a time:
//
//   * CTA k of the cluster holds rows [k RP, (k+1) RP) of the bin: u (with one halo row on either side) and the work
//     array d in shared memory, the right-hand side b in registers.  HBM sees u once on the way in and once on the way
//     out (+ b written once for the Krylov fall-back): 24 B per cell and bin for the whole solve.
//   * the right-hand side b = (I + a L) u + dt D s (k_build_rhs) is formed from the resident u, not by a launch of its own.
//   * row solves: a thread owns two rows x 16 cells (two independent recurrences in flight; the row between them is read
//     once), chunk carries are affine maps composed by a warp-shuffle scan - rows never leave the CTA.
//   * column solves: a thread owns one column x RP rows; the carries between the CTAs of a column are pushed into the
//     other CTAs' shared memory (distributed shared memory), only as far as a carry reaches (products of the
//     multipliers below 1e-18 are not propagated, the bound the factor tables already carry).
//   * the stop test is the componentwise one of the sweep kernels, evaluated on the residual the row solve needs
//     anyway; its verdict travels with the carries.  Bins are dealt to the clusters from a queue, costly bins first.
//
// The LU factors come from the tables of the pipelined sweeps (DiffSlot::fx / fy: m and g = e m per (bin, shift, line
// class, position)); nothing is divided here.
#include "qpb_internal.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

namespace {

constexpr int RNT = 512;     // threads per CTA
constexpr int RCS_MAX = 8;   // CTAs per cluster (portable limit)
constexpr int RNX_MAX = 256; // longest row
constexpr int RREACH = 2;    // CTAs a column carry may reach (slots for 2 sources above and 2 below)

struct ResArgs {
    int ne, ny, nx, ncd, jmax;
    int Q;          // chunks of 16 cells per row
    int CS;         // CTAs per cluster
    int reach;      // CTAs a column carry reaches
    int xdepth;     // chunks a row carry reaches
    int maxit;
    int check_all;  // 1: evaluate the stop test from the first iteration on
    double *S, *B;
    const uint8_t *code;     // [ncd] geometry code of every cell (0 outside the mask)
    const double *dgl;       // [256] code -> linked neighbours + boundary diagonals (grid units)
    const double *src;
    const double *a_bin, *shift, *srccoef, *tol;
    const int *jlen, *known;
    const int *clsx, *clsy;
    const double *mx, *gx;   // [ne][jmax][nclx][npadx], chunk-interleaved 16-byte units
    const double *my, *gy;   // [ne][jmax][ncly][npady]
    const double4 *pax;      // [ne][jmax][nclx][Q]   (forward product, backward product, multiplier into the chunk, -)
    const double4 *pay;      // [ne][jmax][ncly][npady / 16]
    int nclx, ncly, npadx, npady;
    int *done, *iters_out, *queue;
};

// Shared-memory layout of a row: 16-byte unit u of chunk q (cells 16 q + 2u, 16 q + 2u + 1) sits at unit u ^ (q % 8) of
// the chunk.  Lanes that own neighbouring chunks (row solve, 128-bit accesses) and lanes that own neighbouring columns
// (column solve, 64-bit accesses) both sweep all banks, without padding.
__device__ __forceinline__ int scol(int x) { return (x & ~15) | (((((x & 15) >> 1) ^ ((x >> 4) & 7)) << 1) | (x & 1)); }

// unit u of the chunk whose swizzled byte offset is kq = 8 (row RW + 16 q) ^ ((q % 8) << 4)
__device__ __forceinline__ double2 lds2(const double *base, unsigned kq, int u) {
    return *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(base) + (kq ^ (unsigned)(u << 4)));
}
__device__ __forceinline__ void sts2(double *base, unsigned kq, int u, double2 v) {
    *reinterpret_cast<double2 *>(reinterpret_cast<char *>(base) + (kq ^ (unsigned)(u << 4))) = v;
}

// ---- point-to-point hand-over between the CTAs of the cluster ---------------------------------------------------------
// A value is stored into another CTA's shared memory with st.async, which also signs off its bytes on an mbarrier of
// the receiving CTA; the receiver waits on its own barrier.  No cluster-wide barrier, no memory fence at GPU scope
// (cluster.sync() costs a MEMBAR.ALL.GPU and an L1 invalidation each time), and a CTA only waits for the CTAs it hears from.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t peer_addr(uint32_t addr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void send_f64x2(uint32_t raddr, double a, double b, uint32_t rbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "d"(a),
                 "d"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ void send_f64(uint32_t raddr, double a, uint32_t rbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(a), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void send_u32(uint32_t raddr, uint32_t a, uint32_t rbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(raddr), "r"(a), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ unsigned code_byte(const uint4 &f, int t) {
    const unsigned w = t < 4 ? f.x : t < 8 ? f.y : t < 12 ? f.z : f.w;
    return (w >> ((t & 3) * 8)) & 0xffu;
}

// ---- first-order recurrences of the scaled Thomas algorithm on one chunk ------------------------------------------
// forward   y_t = v_t + g_{t-1} y_{t-1}      (g_{-1} = gprev, the last multiplier of the previous chunk)
// backward  x_t = z_t + g_t x_{t+1},  z_t = m_t y_t
// With one D per bin the multiplier is g_t = a m_t wherever cell t has a neighbour t+1 in the mask, and where it has
// none the product a m_t only reaches a cell outside the mask (m = 0 there: its value is dropped and the chain starts
// anew).  So the kernel reads ONE table, m, per solve and forms g on the fly: half the table bytes through L1, and the
// pivots are in registers before the stencil is through (no load in the middle of the solve).
// The probes run a chunk with a zero carry; how a carry passes through the chunk (the product of its multipliers) is
// tabulated per (bin, shift, line class, chunk) at plan time (k_chunk_products).
template <int N>
__device__ __forceinline__ double fwd_probe(const double (&v)[N], const double (&m)[N], double a) {
    double y = v[0];
#pragma unroll
    for (int t = 1; t < N; ++t) y = fma(a * m[t - 1], y, v[t]);
    return y;
}
template <int N>
__device__ __forceinline__ void fwd_apply(double (&v)[N], const double (&m)[N], double a, double gprev, double cin) {
    double y = fma(gprev, cin, v[0]);
    v[0] = y;
#pragma unroll
    for (int t = 1; t < N; ++t) {
        y = fma(a * m[t - 1], y, v[t]);
        v[t] = y;
    }
}
// z = m y, then the backward recurrence with a zero carry
template <int N>
__device__ __forceinline__ double bwd_probe(double (&z)[N], const double (&m)[N], double a) {
#pragma unroll
    for (int t = 0; t < N; ++t) z[t] *= m[t];
    double x = z[N - 1];
#pragma unroll
    for (int t = N - 2; t >= 0; --t) x = fma(a * m[t], x, z[t]);
    return x;
}
template <int N>
__device__ __forceinline__ void bwd_apply(double (&z)[N], const double (&m)[N], double a, double xin) {
    double x = fma(a * m[N - 1], xin, z[N - 1]);
    z[N - 1] = x;
#pragma unroll
    for (int t = N - 2; t >= 0; --t) {
        x = fma(a * m[t], x, z[t]);
        z[t] = x;
    }
}

// carry entering chunk q of a line whose chunks sit in WIDTH adjacent lanes (exclusive scan of affine maps)
// A carry that crosses `depth` chunks is below one part in 1e18: the scan stops after the steps that cover them.
template <int WIDTH, bool REVERSE>
__device__ __forceinline__ double lane_carry(double Am, double Bm, int q, int depth) {
#pragma unroll
    for (int off = 1; off < WIDTH; off <<= 1) {
        if (off > depth) break;
        const double Ao = REVERSE ? __shfl_down_sync(0xffffffffu, Am, off, WIDTH) : __shfl_up_sync(0xffffffffu, Am, off, WIDTH);
        const double Bo = REVERSE ? __shfl_down_sync(0xffffffffu, Bm, off, WIDTH) : __shfl_up_sync(0xffffffffu, Bm, off, WIDTH);
        const bool has = REVERSE ? (q + off < WIDTH) : (q >= off);
        if (has) {
            Bm = fma(Am, Bo, Bm);
            Am = Am * Ao;
        }
    }
    const double prev = REVERSE ? __shfl_down_sync(0xffffffffu, Bm, 1, WIDTH) : __shfl_up_sync(0xffffffffu, Bm, 1, WIDTH);
    const bool first = REVERSE ? (q == WIDTH - 1) : (q == 0);
    return first ? 0.0 : prev;
}

// 16 values of one chunk of an x table (16-byte units of neighbouring chunks side by side)
__device__ __forceinline__ void load_xtab(const double *__restrict__ tab, int Q, int q, double (&o)[16]) {
    const double2 *t2 = reinterpret_cast<const double2 *>(tab);
#pragma unroll
    for (int un = 0; un < 8; ++un) {
        const double2 t = t2[un * Q + q];
        o[2 * un] = t.x;
        o[2 * un + 1] = t.y;
    }
}

// ---- the stencil of one row chunk -----------------------------------------------------------------------------------
// With s = sum of the four neighbours and c1 = 1 + a (linked neighbours + boundary diagonals) (from the per-bin table
// lut, indexed by the cell's geometry code):   A u = c1 u - a s.
// MODE 0: b = (2 - c1) u + a s + sc src  (= u + a L u + dt D s) into Bs and v;  MODE 1: v = b - A u;  MODE 2: the same and the
// stop test: sgn stays negative as long as |b - A u| < tol (|b| + c1 |u| + a sum |neighbours|) in every cell of the mask.
// ku / kd: swizzled byte offsets of the chunk in U (halo row first) and in D / Bs; w0 / e15: the cells left and right of it.
template <int MODE>
__device__ __forceinline__ void row_stencil(const double *__restrict__ U, double *__restrict__ Bs, const double *__restrict__ lut,
                                            unsigned ku, unsigned kd, unsigned rwb, double w0, double e15, const uint4 &cd,
                                            double a, double sc, double tol, const double *__restrict__ src, int c0,
                                            double (&v)[16], int &sgn) {
    double w = w0;
    double2 c = lds2(U, ku, 0);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        double2 cn = make_double2(e15, 0.0);
        if (u < 7) cn = lds2(U, ku, u + 1);
        const double2 n = lds2(U, ku - rwb, u), so = lds2(U, ku + rwb, u);
        const unsigned k0 = code_byte(cd, 2 * u), k1 = code_byte(cd, 2 * u + 1);
        const double c10 = lut[k0], c11 = lut[k1];
        const double s0 = (w + c.y) + (n.x + so.x), s1 = (c.x + cn.x) + (n.y + so.y);
        if (MODE == 0) {
            const double b0 = k0 ? fma(a, s0, fma(2.0 - c10, c.x, sc * src[c0 + 2 * u])) : 0.0;
            const double b1 = k1 ? fma(a, s1, fma(2.0 - c11, c.y, sc * src[c0 + 2 * u + 1])) : 0.0;
            sts2(Bs, kd, u, make_double2(b0, b1));
            v[2 * u] = b0;
            v[2 * u + 1] = b1;
        } else {
            const double2 b = lds2(Bs, kd, u);
            const double r0 = fma(a, s0, fma(-c10, c.x, b.x)), r1 = fma(a, s1, fma(-c11, c.y, b.y));
            v[2 * u] = r0;
            v[2 * u + 1] = r1;
            if (MODE == 2) {
                const double q0 = (fabs(w) + fabs(c.y)) + (fabs(n.x) + fabs(so.x));
                const double q1 = (fabs(c.x) + fabs(cn.x)) + (fabs(n.y) + fabs(so.y));
                const double x0 = fma(-tol, fma(a, q0, fma(fabs(c10), fabs(c.x), fabs(b.x))), fabs(r0));
                const double x1 = fma(-tol, fma(a, q1, fma(fabs(c11), fabs(c.y), fabs(b.y))), fabs(r1));
                // excess <= -0.0 keeps the sign bit; a cell outside the mask (code 0) does not vote
                sgn &= k0 ? __double2hiint(x0) : -1;
                sgn &= k1 ? __double2hiint(x1) : -1;
            }
        }
        w = c.y;
        c = cn;
    }
}

// Products of the multipliers a carry meets on its way through a chunk: forward  g_{16q-1} g_{16q} ... g_{16q+14}
// (into the chunk and up to its last cell), backward  g_{16q} ... g_{16q+15}; with them g_{16q-1} itself, so that a solve
// reads nothing outside its own chunk of the tables.  One thread per (table row, chunk).
__global__ void k_chunk_products(long long nrows, int Qc, int interleaved, const double *__restrict__ g, double4 *__restrict__ out) {
    const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (id >= nrows * Qc) return;
    const long long row = id / Qc;
    const int q = (int)(id - row * Qc);
    const double *gr = g + row * (long long)Qc * 16;
    auto at = [&](int k) { return interleaved ? gr[(((k % 16) / 2) * Qc + k / 16) * 2 + (k & 1)] : gr[k]; };
    const double gin = q > 0 ? at(16 * q - 1) : 0.0;   // multiplier between the previous chunk's last cell and this chunk
    double f = gin, b = 1.0;
    for (int t = 0; t < 16; ++t) {
        const double gv = at(16 * q + t);
        if (t < 15) f *= gv;
        b *= gv;
    }
    out[id] = make_double4(f, b, gin, 0.0);
}

// RP rows per CTA (16 or 32), QP lanes per row in the row solve (power of two >= chunks per row)
// NXT: row length when it is known at compile time (every shared-memory address becomes an immediate), 0 otherwise
// FULL: the grid fills the cluster exactly (every thread owns a row chunk and a column chunk that exist, ny = CS * RP):
// the guards and the neutral values of absent chunks fall away at compile time
template <int RP, int QP, int NXT, bool FULL>
__global__ void __launch_bounds__(RNT, 1) k_pr_resident(const __grid_constant__ ResArgs A) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(128) double sm[];
    constexpr int NH = RP / 16;                       // column chunks of 16 rows per CTA
    const int RW = NXT ? NXT : A.nx;                  // doubles per shared-memory row (a multiple of 16)
    double *U = sm;                                   // [RP + 2][RW]  u with a halo row above and below
    double *D = U + (size_t)(RP + 2) * RW;            // [RP][RW]      d* / d
    double *Bs = D + (size_t)RP * RW;                 // [RP][RW]      right-hand side b
    double2 *slots = reinterpret_cast<double2 *>(Bs + (size_t)RP * RW);   // [2 RREACH][RNX_MAX] carry maps of other CTAs
    double2 *pairs = slots + 2 * RREACH * RNX_MAX;    // [2][RNX_MAX]  maps of the CTA's own other half (down / up)
    double *lut = reinterpret_cast<double *>(pairs + 2 * RNX_MAX);        // [256] 1 + a dg per geometry code
    double *shf = lut + 256;                          // [64] shifts of the bin
    int *ism = reinterpret_cast<int *>(shf + 64);
    int *conv = ism;           // [2][RCS_MAX] "some cell of CTA r is over its bound", by iteration parity
    int *binslot = ism + 16;   // bin handed out by CTA 0
    int *red = ism + 20;       // [16] per-warp verdicts
    // mbarriers: 0 column carries from above, 1 from below, 2 halo rows, 3 / 4 verdicts of even / odd iterations
    const uint32_t mb0 = smem_u32(ism + 40);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank(), CS = A.CS, reach = A.reach;
    const int nx = NXT ? NXT : A.nx, ny = A.ny, Q = NXT ? NXT / 16 : A.Q;
    const int y0 = rank * RP;
    // row solve: row g of the CTA, chunk q
    const int g = tid / QP, q = tid % QP;
    const bool xrow = FULL || g < RP;                       // whole warps
    const bool xact = FULL || (xrow && q < Q);
    const bool rowl = FULL || (xact && y0 + g < ny);        // the row exists
    // column solve: column yx, rows [16 h, 16 h + 16) of the CTA
    const int yx = tid & (RNX_MAX - 1), h = tid / RNX_MAX;
    const bool yact = FULL || (yx < nx && h < NH);
    const int ycol = scol(yact ? yx : 0);
    const int yr0 = y0 + 16 * h;                // first row of the column chunk in the grid

    for (int e = tid; e < (3 * RP + 2) * RW; e += RNT) sm[e] = 0.0;
    if (tid < 64) ism[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < 5; ++k) mbar_init(mb0 + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    unsigned ph = 0;   // phase parities of the five barriers, bit k = barrier k
    // what this CTA hears per iteration: carry maps of nA CTAs above and nB below (16 B per column), halo rows (8 B per column)
    const int nA = min(reach, rank), nB = min(reach, CS - 1 - rank), nC = (rank > 0) + (rank < CS - 1);
    const uint32_t slots_u32 = smem_u32(slots), conv_u32 = smem_u32(conv), U_u32 = smem_u32(U);

    // geometry of the thread's row chunk: the same for every bin
    uint4 cd = make_uint4(0, 0, 0, 0);
    const int qs = xact ? q : 0, gs = xact ? g : 0;
    const int c0 = (y0 + gs) * nx + 16 * qs;    // dense index of the chunk's first cell
    int clsr = 0;
    if (rowl) {
        cd = *reinterpret_cast<const uint4 *>(A.code + c0);
        clsr = A.clsx[y0 + g];
    }
    const int clsc = yact ? A.clsy[yx] : 0;
    const unsigned rwb = 8u * (unsigned)RW;
    const unsigned ku = (8u * (unsigned)((gs + 1) * RW + 16 * qs)) ^ ((unsigned)(qs & 7) << 4);   // chunk in U (halo row first)
    const unsigned kd = (8u * (unsigned)(gs * RW + 16 * qs)) ^ ((unsigned)(qs & 7) << 4);         // chunk in D and Bs
    // cells left and right of the chunk: last cell of chunk q-1, first cell of chunk q+1 (double index within U)
    // (at the ends of the row: a word of shared memory that stays zero)
    const int izero = (int)(reinterpret_cast<double *>(ism + 56) - U);
    const int iw = qs > 0 ? (gs + 1) * RW + 16 * (qs - 1) + (((7 ^ ((qs - 1) & 7)) << 1) | 1) : izero;
    const int ie = qs < Q - 1 ? (gs + 1) * RW + 16 * (qs + 1) + ((((qs + 1) & 7)) << 1) : izero;

    for (;;) {
        if (rank == 0 && tid == 0) {
            const int nb = atomicAdd(A.queue, 1);
            for (int r = 0; r < CS; ++r) *cluster.map_shared_rank(binslot, r) = nb;
        }
        cluster.sync();
        const int ord = *binslot;
        if (ord >= A.ne) break;
        const int bin = A.ne - 1 - ord;   // D(E) grows with E: the bins that iterate longest go first
        const double a = A.a_bin[bin], tol = A.tol[bin], sc = A.srccoef[bin];
        const int jl = A.jlen[bin];
        const int known = A.known[bin];
        const int check_from = (A.check_all || known <= 0) ? 0 : max(0, known - 2);

        // ---- the bin comes on chip ------------------------------------------------------------------------------
        if (tid < 256) lut[tid] = fma(a, A.dgl[tid], 1.0);
        if (tid < jl && tid < 64) shf[tid] = A.shift[(size_t)bin * A.jmax + tid];   // visible after the barrier below
        if (xact) {
            double2 w[8];
            if (rowl) {
                const double2 *s2 = reinterpret_cast<const double2 *>(A.S + (size_t)bin * A.ncd + c0);
#pragma unroll
                for (int un = 0; un < 8; ++un) w[un] = s2[un];
            } else {
#pragma unroll
                for (int un = 0; un < 8; ++un) w[un] = make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int un = 0; un < 8; ++un) sts2(U, ku, un, w[un]);
            // first / last row of the CTA: also the halo row of the neighbour above / below
            if (g == 0 && rank > 0) {
                double *r = cluster.map_shared_rank(U, rank - 1);
                const unsigned kh = (8u * (unsigned)((RP + 1) * RW + 16 * q)) ^ ((unsigned)(q & 7) << 4);
#pragma unroll
                for (int un = 0; un < 8; ++un) sts2(r, kh, un, w[un]);
            }
            if (g == RP - 1 && rank < CS - 1) {
                double *r = cluster.map_shared_rank(U, rank + 1);
                const unsigned kh = (8u * (unsigned)(16 * q)) ^ ((unsigned)(q & 7) << 4);
#pragma unroll
                for (int un = 0; un < 8; ++un) sts2(r, kh, un, w[un]);
            }
        }
        cluster.sync();

        if (xact) {
            double v[16];
            int sg = -1;
            row_stencil<0>(U, Bs, lut, ku, kd, rwb, U[iw], U[ie], cd, a, sc, tol, A.src, c0, v, sg);
            // b also goes to memory: the Krylov fall-back of a solve that stalls starts from it
            if (rowl) {
                double2 *o = reinterpret_cast<double2 *>(A.B + (size_t)bin * A.ncd + c0);
#pragma unroll
                for (int un = 0; un < 8; ++un) o[un] = make_double2(v[2 * un], v[2 * un + 1]);
            }
        }

        // ---- iterate ------------------------------------------------------------------------------------------------
        int it = 0;
        bool converged = false;
        int j = -1;
        for (; it < A.maxit; ++it) {
            j = j + 1 == jl ? 0 : j + 1;   // it % jl
            const double r = shf[j];
            const bool checking = it >= check_from;
            if (tid == 0) {
                if (nA) mbar_expect(mb0, (uint32_t)(nA * nx * 16));
                if (checking) mbar_expect(mb0 + 8 * (3 + (it & 1)), (uint32_t)(4 * CS));
            }
            // -- rows: d* = (H + r)^-1 (b - A u)
            int sgn = -1;
            if (xrow) {   // whole warps: lanes beyond the last chunk of a row run the scans on the neutral map
                double v[16], mt[16];
                double gp = 0.0;   // last multiplier of the chunk to the left
                const size_t tb = (((size_t)bin * A.jmax + j) * A.nclx + clsr) * A.npadx;
                double4 pa = make_double4(0.0, 0.0, 0.0, 0.0);
                if (rowl) {
                    load_xtab(A.mx + tb, Q, q, mt);
                    pa = A.pax[(((size_t)bin * A.jmax + j) * A.nclx + clsr) * Q + q];
                    gp = pa.z;
                    const double w0 = U[iw], e15 = U[ie];
                    if (checking)
                        row_stencil<2>(U, Bs, lut, ku, kd, rwb, w0, e15, cd, a, sc, tol, A.src, c0, v, sgn);
                    else
                        row_stencil<1>(U, Bs, lut, ku, kd, rwb, w0, e15, cd, a, sc, tol, A.src, c0, v, sgn);
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) mt[t] = v[t] = 0.0;
                }
                const double yin = lane_carry<QP, false>(pa.x, fwd_probe(v, mt, a), q, A.xdepth);
                fwd_apply(v, mt, a, gp, yin);
                const double xin = lane_carry<QP, true>(pa.y, bwd_probe(v, mt, a), q, A.xdepth);
                bwd_apply(v, mt, a, xin);
                if (xact) {
#pragma unroll
                    for (int un = 0; un < 8; ++un) sts2(D, kd, un, make_double2(v[2 * un], v[2 * un + 1]));
                }
            }
            if (checking) {
                sgn = __reduce_and_sync(0xffffffffu, sgn);
                if (lane == 0) red[warp] = sgn;
            }
            if (yact) {   // the factor tables of the column solve: into L1 behind the barrier
                const size_t tby = (((size_t)bin * A.jmax + j) * A.ncly + clsc) * A.npady + yr0;
                if (yr0 < A.npady) {
                    prefetch_l1(A.pay + (tby >> 4));
                    prefetch_l1(A.my + tby);
                }
            }
            __syncthreads();   // d* complete, per-warp verdicts visible
            if (checking && tid < CS) {
                int all = -1;
#pragma unroll
                for (int w = 0; w < RNT / 32; ++w) all &= red[w];
                // sign bit lost: some cell of mine is over its bound; every CTA (this one too) gets the verdict
                send_u32(peer_addr(conv_u32 + 4 * ((it & 1) * RCS_MAX + rank), tid), all >= 0 ? 1u : 0u,
                         peer_addr(mb0 + 8 * (3 + (it & 1)), tid));
            }
            // -- columns: d = (V + r)^-1 d*, forward elimination
            double v[16], mt[16], gp = 0.0;
            const bool ytab = FULL || (yact && yr0 < A.npady);   // the chunk lies inside the tables (they end on a multiple of 16)
            double Af = 0.0, Bf = 0.0;
            double4 pa = make_double4(0.0, 0.0, 0.0, 0.0);
            if (yact) {
                const size_t tb = (((size_t)bin * A.jmax + j) * A.ncly + clsc) * A.npady;
                if (ytab) {
                    const double2 *m2 = reinterpret_cast<const double2 *>(A.my + tb + yr0);
#pragma unroll
                    for (int un = 0; un < 8; ++un) {
                        const double2 t2 = m2[un];
                        mt[2 * un] = t2.x;
                        mt[2 * un + 1] = t2.y;
                    }
                    pa = A.pay[(tb + yr0) >> 4];
                    gp = pa.z;
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) mt[t] = 0.0;
                }
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = D[(16 * h + t) * RW + ycol];
                Af = pa.x;
                Bf = fwd_probe(v, mt, a);
                if (NH == 2 && h == 0) pairs[yx] = make_double2(Af, Bf);
            }
            if (NH == 2) __syncthreads();
            if (yact && h == NH - 1) {
                double Ac = Af, Bc = Bf;
                if (NH == 2) {   // the CTA's map: upper half first
                    const double2 up = pairs[yx];
                    Bc = fma(Af, up.y, Bf);
                    Ac = Af * up.x;
                }
                // my map is source "k CTAs above" for the CTA k below me: slot RREACH - k there
                for (int k = 1; k <= reach && rank + k < CS; ++k)
                    send_f64x2(peer_addr(slots_u32 + 16 * ((RREACH - k) * RNX_MAX + yx), rank + k), Ac, Bc,
                               peer_addr(mb0, rank + k));
            }
            if (nA) {
                mbar_wait(mb0, ph & 1u);
                ph ^= 1u;
            }
            if (checking) {
                const unsigned bit = 8u << (it & 1);
                mbar_wait(mb0 + 8 * (3 + (it & 1)), (ph & bit) ? 1u : 0u);
                ph ^= bit;
                int over = 0;
                for (int rr = 0; rr < CS; ++rr) over |= conv[(it & 1) * RCS_MAX + rr];
                if (!over) {
                    converged = true;
                    break;     // the input of this iteration satisfies the system in every cell: u is the answer
                }
            }
            if (tid == 0) {   // the iteration goes on: carries from below and the new halo rows are on their way
                if (nB) mbar_expect(mb0 + 8, (uint32_t)(nB * nx * 16));
                if (nC) mbar_expect(mb0 + 16, (uint32_t)(nC * nx * 8));
            }
            double Ab = 0.0, Bb = 0.0;
            if (yact) {
                double cin = 0.0;
                for (int k = min(reach, rank); k >= 1; --k) {   // farthest source first
                    const double2 m = slots[(RREACH - k) * RNX_MAX + yx];
                    cin = fma(m.x, cin, m.y);
                }
                if (NH == 2 && h == 1) {
                    const double2 up = pairs[yx];
                    cin = fma(up.x, cin, up.y);
                }
                fwd_apply(v, mt, a, gp, cin);
                Ab = pa.y;
                Bb = bwd_probe(v, mt, a);
                if (NH == 2 && h == 1) pairs[RNX_MAX + yx] = make_double2(Ab, Bb);
            }
            if (NH == 2) __syncthreads();
            if (yact && h == 0) {
                double Ac = Ab, Bc = Bb;
                if (NH == 2) {   // the CTA's map: lower half first
                    const double2 dn = pairs[RNX_MAX + yx];
                    Bc = fma(Ab, dn.y, Bb);
                    Ac = Ab * dn.x;
                }
                // source "k CTAs below" for the CTA k above me: slot RREACH + k - 1 there
                for (int k = 1; k <= reach && rank - k >= 0; ++k)
                    send_f64x2(peer_addr(slots_u32 + 16 * ((RREACH + k - 1) * RNX_MAX + yx), rank - k), Ac, Bc,
                               peer_addr(mb0 + 8, rank - k));
            }
            if (nB) {
                mbar_wait(mb0 + 8, (ph >> 1) & 1u);
                ph ^= 2u;
            }
            if (yact) {
                double xin = 0.0;
                for (int k = min(reach, CS - 1 - rank); k >= 1; --k) {
                    const double2 m = slots[(RREACH + k - 1) * RNX_MAX + yx];
                    xin = fma(m.x, xin, m.y);
                }
                if (NH == 2 && h == 0) {
                    const double2 dn = pairs[RNX_MAX + yx];
                    xin = fma(dn.x, xin, dn.y);
                }
                bwd_apply(v, mt, a, xin);
                const double r2 = 2.0 * r;
                double *uc = U + (16 * h + 1) * RW + ycol;
                double first = 0.0, last = 0.0;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const double un = fma(r2, v[t], uc[t * RW]);
                    uc[t * RW] = un;
                    if (t == 0) first = un;
                    if (t == 15) last = un;
                }
                if (h == 0 && rank > 0)
                    send_f64(peer_addr(U_u32 + 8 * ((RP + 1) * RW + ycol), rank - 1), first, peer_addr(mb0 + 16, rank - 1));
                if (h == NH - 1 && rank < CS - 1)
                    send_f64(peer_addr(U_u32 + 8 * ycol, rank + 1), last, peer_addr(mb0 + 16, rank + 1));
            }
            // the factor tables of the next row solve: into L1 while the halo rows travel
            if (rowl && it + 1 < A.maxit) {
                const size_t tbn = (((size_t)bin * A.jmax + (j + 1 == jl ? 0 : j + 1)) * A.nclx + clsr) * A.npadx;
                const size_t o = ((size_t)(q & 7) * Q + q) * 2;
                prefetch_l1(A.mx + tbn + o);
                prefetch_l1(A.pax + (tbn >> 4) + q);
            }
            __syncthreads();   // u is current in this CTA; everybody is done with the carry slots
            if (nC) {
                mbar_wait(mb0 + 16, (ph >> 2) & 1u);   // ... and its halo rows have arrived
                ph ^= 4u;
            }
        }

        // ---- the bin goes back --------------------------------------------------------------------------------------
        if (rowl) {
            double2 *o = reinterpret_cast<double2 *>(A.S + (size_t)bin * A.ncd + c0);
#pragma unroll
            for (int un = 0; un < 8; ++un) o[un] = lds2(U, ku, un);
        }
        if (rank == 0 && tid == 0) {
            A.done[bin] = converged ? 1 : 0;
            A.iters_out[bin] = it;
        }
        // the next round's first cluster barrier separates these reads of U (and of lut) from the next bin's stores
    }
}

template <int RP, int QP, int NXT, bool FULL = false>
int launch_resident(qpb_ctx *c, const DiffSlot &s, const ResArgs &A, bool query, int *nclusters) {
    auto kern = k_pr_resident<RP, QP, NXT, FULL>;
    // per device and per launch: the attribute is device state, and grids of different row lengths share an instance
    QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(std::max(1, s.res.nclusters) * s.res.CS));
    cfg.blockDim = dim3(RNT);
    cfg.dynamicSmemBytes = s.res.smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)s.res.CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (query) {
        int n = 0;
        cfg.gridDim = dim3((unsigned)(64 * s.res.CS));
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        *nclusters = n;
        return QPB_OK;
    }
    QPB_CUDA(cudaLaunchKernelEx(&cfg, kern, A));
    return QPB_OK;
}

int dispatch_resident(qpb_ctx *c, const DiffSlot &s, const ResArgs &A, bool query, int *nclusters) {
    const bool n256 = c->cfg.nx == 256;
    if (s.res.RP == 32) {
        // 256 x 256 (BASELINE configs[1]): RNT = 32 rows x 16 chunks = 256 columns x 2 halves, nothing is absent
        if (n256 && c->cfg.ny == 256 && s.fy.npad == 256) return launch_resident<32, 16, 256, true>(c, s, A, query, nclusters);
        if (n256) return launch_resident<32, 16, 256>(c, s, A, query, nclusters);
        if (s.res.QP == 16) return launch_resident<32, 16, 0>(c, s, A, query, nclusters);
        return launch_resident<32, 8, 0>(c, s, A, query, nclusters);
    }
    if (n256) return launch_resident<16, 16, 256>(c, s, A, query, nclusters);
    if (s.res.QP == 16) return launch_resident<16, 16, 0>(c, s, A, query, nclusters);
    return launch_resident<16, 8, 0>(c, s, A, query, nclusters);
}

}  // namespace

// Decide whether the prepared solve of this slot runs bin-resident (after the tables of the pipelined sweeps exist).
int qpbr_plan(qpb_ctx *c, DiffSlot &s) {
    s.res = DiffSlot::Resident();
    if (const char *e = getenv("QPB_NO_RESIDENT"))
        if (e[0] == '1') return QPB_OK;
    const auto &cf = c->cfg;
    if (cf.flags & QPB_F_VARIABLE_D) return QPB_OK;
    if (s.mode != 0 || !s.fast || s.spectral || s.krylov || s.jmax > 64) return QPB_OK;
    if (!s.fx.d_tab || !s.fx.d_tabg || !s.fy.d_tab || !s.fy.d_tabg || s.fx.S != 16 || s.fy.S != 16) return QPB_OK;
    if (cf.nx % 16 != 0 || cf.nx < 96 || cf.nx > RNX_MAX || cf.ny < 32 || cf.ny > 32 * RCS_MAX) return QPB_OK;
    DiffSlot::Resident r;
    const int Q = cf.nx / 16;
    r.QP = Q > 8 ? 16 : 8;
    r.pitch = cf.nx;
    const int rows = s.fy.carry_depth >= s.fy.Q ? cf.ny : s.fy.carry_depth * 16;   // rows a column carry reaches
    // 16 rows per CTA spread a small grid over more SMs; 32 when the grid is taller than 8 x 16 rows or when a carry
    // would cross more than RREACH CTAs of 16 rows
    for (r.RP = cf.ny > 16 * RCS_MAX ? 32 : 16; r.RP <= 32; r.RP += 16) {
        r.CS = (cf.ny + r.RP - 1) / r.RP;
        r.reach = r.CS > 1 ? std::max(1, std::min(r.CS - 1, (rows + r.RP - 1) / r.RP)) : 0;
        if (r.reach <= RREACH) break;
    }
    if (r.RP > 32) return QPB_OK;   // carries that cross more than 2 CTAs of 32 rows: the launched sweeps take the solve
    r.smem = sizeof(double) * (size_t)(3 * r.RP + 2) * r.pitch + sizeof(double2) * (2 * RREACH + 2) * RNX_MAX +
             sizeof(double) * (256 + 64) + 512;
    // geometry codes: one byte per cell into the table of distinct diagonals (linked neighbours + boundary terms)
    std::vector<uint8_t> code(c->ncd, 0);
    std::vector<double> dgl(256, 0.0);
    int ncode = 1;   // code 0: outside the mask
    for (int p = 0; p < c->ncd; ++p) {
        const unsigned f = c->h_flags[p];
        if (!(f & QPB_IN)) continue;
        const double dg = (double)(((f & QPB_LK_L) ? 1 : 0) + ((f & QPB_LK_R) ? 1 : 0) + ((f & QPB_LK_U) ? 1 : 0) +
                                   ((f & QPB_LK_D) ? 1 : 0)) + (c->h_bcx[p] + c->h_bcy[p]);
        int k = 1;
        while (k < ncode && dgl[k] != dg) ++k;
        if (k == ncode) {
            if (ncode == 256) return QPB_OK;   // more distinct wall terms than a byte can name: launched sweeps
            dgl[ncode++] = dg;
        }
        code[p] = (uint8_t)k;
    }
    s.res = r;
    ResArgs A{};
    int n = 0;
    const int rc = dispatch_resident(c, s, A, true, &n);
    if (rc != QPB_OK || n < 1) {
        s.res = DiffSlot::Resident();
        return QPB_OK;
    }
    s.res.nclusters = std::min(n, cf.ne);
    if (const char *e = getenv("QPB_DEBUG_RES"))
        if (e[0] == '1')
            fprintf(stderr, "[qpb] resident solve: %d rows per CTA, clusters of %d, %d clusters resident, carry reach %d CTAs, "
                            "%zu B shared memory\n", r.RP, r.CS, n, r.reach, r.smem);
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_resq, sizeof(int)));
    {
        const long long rx = (long long)cf.ne * s.jmax * s.fx.nclass, ry = (long long)cf.ne * s.jmax * s.fy.nclass;
        const int Qx = s.fx.npad / 16, Qy = s.fy.npad / 16;
        QPB_CUDA(qpb_dev_malloc((void **)&s.d_respax, sizeof(double4) * (size_t)(rx * Qx)));
        QPB_CUDA(qpb_dev_malloc((void **)&s.d_respay, sizeof(double4) * (size_t)(ry * Qy)));
        k_chunk_products<<<(unsigned)ceil_div64(rx * Qx, 128), 128, 0, c->stream>>>(rx, Qx, 1, s.fx.d_tabg, (double4 *)s.d_respax);
        k_chunk_products<<<(unsigned)ceil_div64(ry * Qy, 128), 128, 0, c->stream>>>(ry, Qy, 0, s.fy.d_tabg, (double4 *)s.d_respay);
        QPB_CHECK_LAUNCH();
        c->diag.kernel_launches += 2;
    }
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_rescode, (size_t)c->ncd));
    QPB_CUDA(qpb_dev_malloc((void **)&s.d_reslut, sizeof(double) * 256));
    QPB_CUDA(cudaMemcpy(s.d_rescode, code.data(), (size_t)c->ncd, cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(s.d_reslut, dgl.data(), sizeof(double) * 256, cudaMemcpyHostToDevice));
    s.res.ok = true;
    return QPB_OK;
}

// A u = rhs for every bin of the slot, state in c->d_S in place (the right-hand side is formed on chip).  Returns
// QPB_E_NOCONV - with b in c->d_B - when a bin did not meet the stop test within the iteration cap.
int qpbr_solve(qpb_ctx *c, DiffSlot &s, std::vector<int> &h_done) {
    const auto &cf = c->cfg;
    const int ne = cf.ne;
    ResArgs A{};
    A.ne = ne; A.ny = cf.ny; A.nx = cf.nx; A.ncd = c->ncd; A.jmax = s.jmax;
    A.Q = cf.nx / 16; A.CS = s.res.CS; A.reach = s.res.reach; A.maxit = c->maxit;
    A.xdepth = std::max(1, s.fx.carry_depth);
    A.pax = (const double4 *)s.d_respax; A.pay = (const double4 *)s.d_respay;
    A.check_all = !(s.known_iters > 0 && (s.solves % 16) != 0);
    A.S = c->d_S; A.B = c->d_B; A.code = s.d_rescode; A.dgl = s.d_reslut; A.src = c->d_srcgeom;
    A.a_bin = s.d_a; A.shift = s.d_shift; A.srccoef = s.d_src; A.tol = s.d_tol; A.jlen = s.d_jlen; A.known = s.d_known;
    A.clsx = s.fx.d_cls; A.clsy = s.fy.d_cls;
    A.mx = s.fx.d_tab; A.gx = s.fx.d_tabg; A.my = s.fy.d_tab; A.gy = s.fy.d_tabg;
    A.nclx = s.fx.nclass; A.ncly = s.fy.nclass; A.npadx = s.fx.npad; A.npady = s.fy.npad;
    A.done = c->d_done; A.iters_out = c->d_done + ne; A.queue = s.d_resq;
    QPB_CUDA(cudaMemsetAsync(c->d_done, 0, sizeof(int) * 2 * (size_t)ne, c->stream));
    QPB_CUDA(cudaMemsetAsync(s.d_resq, 0, sizeof(int), c->stream));
    {
        ScopedTimer tm(c, 0);
        int rc = dispatch_resident(c, s, A, false, nullptr);
        if (rc != QPB_OK) return rc;
        c->diag.kernel_launches++;
    }
    h_done.resize(2 * (size_t)ne);
    QPB_CUDA(cudaMemcpyAsync(h_done.data(), c->d_done, sizeof(int) * 2 * (size_t)ne, cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    for (int b = 0; b < ne; ++b)
        if (!h_done[b]) return QPB_E_NOCONV;
    return QPB_OK;
}
