// Persistent, TMA-pipelined tridiagonal sweeps (the hot diffusion kernels for uniform D, lines up to 512 cells).
//
// Same linear algebra as qpb_sweep_fast.cu / qpb_diffusion.cu (one Peaceman-Rachford half step per launch:
// solver.py:1428-1452 restated as batched x / y tridiagonal solves), reorganised so that the kernel is bound by
// HBM and not by instruction issue or load latency:
//
//  * one persistent CTA per SM (8 warps, the whole register file) walks a static list of tiles (bin, block of
//    lines).  A ring of NS input stages is kept full with TMA (cp.async.bulk.tensor + mbarrier complete_tx): at the
//    top of iteration k one elected thread issues the loads of tile k+NS-1 into the stage that iteration k-1
//    released; results leave through a double-buffered shared-memory tile and a TMA store.  Loads of tiles
//    k+1..k+NS-1, the solve of tile k and the store of tile k-1 overlap.
//  * the sweeps contain no selects, bit tests or divisions: the geometry enters as two per-cell doubles
//    (cx, cy = number of linked neighbours + boundary diagonal, 0 outside the mask; cells outside the mask hold
//    u = 0, so  sum over linked neighbours of (u0 - u_nb) = c*u0 - u_left - u_right ) and the LU factors come from
//    tables  m_t = 1/pivot  (0 outside the mask) and  g_t = e_{t+1} m_t  factored once per prepared step.
//  * scaled Thomas recurrences   y~_t = d_t + g_{t-1} y~_{t-1},   x_t = m_t y~_t + g_t x_{t+1}:  2 FP64 operations
//    per cell and direction; a thread owns a 16-cell chunk, chunk carries are affine maps composed by a
//    warp-shuffle scan (x sweep) or through shared memory (y sweep), then added back as a geometric correction.
//
// Traffic per cell*bin: x sweep reads u (+2 halo rows per 16) and b, writes u*; y sweep reads u* and u, writes u.
#include "qpb_internal.h"

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// Programmatic dependent launch: the next sweep kernel of the stream may start its prologue (barrier init, geometry
// of its block of lines) on SMs that this grid has already left; it reads nothing a predecessor wrote before the wait.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Inclusive Kogge-Stone scan of affine maps over WIDTH adjacent lanes; returns the carry entering this lane's chunk.
// `reach`: only offsets below it are composed (the maps contract so fast that farther chunks cannot be felt).
template <int WIDTH, bool REVERSE>
__device__ __forceinline__ double warp_carry(double A, double B, int q, int reach) {
#pragma unroll
    for (int off = 1; off < WIDTH; off <<= 1) {
        if (off >= reach) break;
        const double Ao = REVERSE ? __shfl_down_sync(0xffffffffu, A, off, WIDTH) : __shfl_up_sync(0xffffffffu, A, off, WIDTH);
        const double Bo = REVERSE ? __shfl_down_sync(0xffffffffu, B, off, WIDTH) : __shfl_up_sync(0xffffffffu, B, off, WIDTH);
        const bool has = REVERSE ? (q + off < WIDTH) : (q >= off);
        if (has) {
            B = fma(A, Bo, B);
            A = A * Ao;
        }
    }
    const double prev = REVERSE ? __shfl_down_sync(0xffffffffu, B, 1, WIDTH) : __shfl_up_sync(0xffffffffu, B, 1, WIDTH);
    const bool first = REVERSE ? (q == WIDTH - 1) : (q == 0);
    return first ? 0.0 : prev;
}

struct PipeArgs {
    int ne, ny, nx, iter, jmax;
    const double *tol;   // [ne]
    const double *cx, *cy;       // dense geometry doubles
    const uint8_t *flags;        // dense flag bytes (QPB_IN = 16)
    const double *a_bin, *shift;
    const int *jlen;
    const int *cls;              // class of every line
    const double *tabm, *tabg;   // [ne][jmax][nclass][npad]
    int nclass, npad, Q;         // Q = chunks per line (table layout), npad = Q * S
    // Lines longer than 32 chunks are cut into segments of `qi` chunks that are solved with `halo` extra chunks on
    // either side and zero carries at the cut: the carry reach (FastDir::carry_depth, products of g below 1e-18)
    // bounds the influence of everything farther away, so the interior of a segment is exact to that level.
    int qs, qi, halo, nseg;      // chunks per tile (halo included), interior chunks, halo chunks, segments per line
    // 1: the x sweep stores u* - u and the y sweep solves for it directly (both sweeps pipelined).  The y sweep updates
    // u in place, so its halo rows may already hold a neighbouring tile's new values: with the difference coming
    // from the x sweep, the old u is only read where this tile alone writes.
    int delta;
    // 1: the result of a tile is written over one of its own input tiles and stored from there (no separate output
    // buffers; the stage is refilled in the middle of the next iteration, after its store has read it);
    // 0: double-buffered output tiles, the stage is refilled at the top of the next iteration.
    int inplace;
    int tiles_per_bin, ntiles;
    int depth;                   // carry reach in chunks (see FastDir::carry_depth)
    int check_all;               // 1: every bin measures / tests the residual in this iteration
    int rev;                     // 1: tiles are walked from the last bin to the first (see qpbp_sweep: L2 reuse between sweeps)
    int prefetch;                // 1: the factor-table lines of the next tile are prefetched into L1 while this one is solved
    const int *known;            // [ne] iterations the previous solve of each bin needed (<= 0: unknown)
    unsigned long long *res, *unorm;
    int *done, *iters_out;
};

struct XMaps {
    CUtensorMap u, b, out;
};
struct YMaps {
    CUtensorMap u, w, out;
};

// MODE 0 (x sweep): a bin is active until it is marked done.  MODE 1 (y sweep): additionally the residual that the
// x sweep of this iteration measured decides; the first tile of a bin records the decision.
// A bin measures (x sweep) and tests (y sweep) the residual only from two iterations before the count its previous
// solve needed: consecutive time steps need about as many, and the unchecked sweeps skip a third of the arithmetic.
__device__ __forceinline__ bool bin_checks(const PipeArgs &A, int bin) {
    if (A.check_all) return true;
    const int kn = A.known[bin];
    return kn <= 0 || A.iter >= kn - 2;
}

template <int MODE>
__device__ __forceinline__ bool tile_active(const PipeArgs &A, int bin, bool leader) {
    if (A.done[bin]) return false;
    if (MODE == 1 && bin_checks(A, bin)) {
        const double r = __longlong_as_double((long long)A.res[(long long)A.iter * A.ne + bin]);
        const double un = __longlong_as_double((long long)A.unorm[(long long)A.iter * A.ne + bin]);
        (void)un;
        if (r <= 0.0) {   // no cell exceeded its componentwise bound (recorded by the x sweep)
            if (leader) {
                A.iters_out[bin] = A.iter;
                __threadfence();
                A.done[bin] = 1;
            }
            return false;
        }
    }
    return true;
}

// scaled Thomas solve of one S-cell chunk held in v[]; carries are resolved by the caller
template <int S>
struct ChunkSolve {
    double v[S];
    double g[S];
    double m[S];

    // forward, zero carry in.  Returns (A, B): carry_out = A*carry_in + B.
    __device__ __forceinline__ void forward(double &A, double &B) {
        double y = v[0], P = g[0];
#pragma unroll
        for (int t = 1; t < S; ++t) {
            y = fma(g[t - 1], y, v[t]);
            v[t] = y;
            P *= g[t];
        }
        A = P;
        B = g[S - 1] * y;
    }
    __device__ __forceinline__ void forward_fix(double carry) {
        double c = carry;
        v[0] += c;
#pragma unroll
        for (int t = 1; t < S; ++t) {
            c *= g[t - 1];
            v[t] += c;
        }
    }
    // backward, zero carry in.  Returns B (A is the same product as in forward).
    __device__ __forceinline__ double backward() {
        double x = m[S - 1] * v[S - 1];
        v[S - 1] = x;
#pragma unroll
        for (int t = S - 2; t >= 0; --t) {
            x = fma(g[t], x, m[t] * v[t]);
            v[t] = x;
        }
        return x;
    }
    __device__ __forceinline__ void backward_fix(double carry) {
        double c = carry;
#pragma unroll
        for (int t = S - 1; t >= 0; --t) {
            c *= g[t];
            v[t] += c;
        }
    }
};

template <int NT>
__device__ __forceinline__ void cta_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }

// per-bin parameters of this launch in shared memory (the tile loops touch no global scalars)
template <int MODE, int NT>
__device__ __forceinline__ void load_bin_params(const PipeArgs &A, double *s_a, double *s_rho, int *s_j) {
    for (int b = threadIdx.x; b < A.ne; b += NT) {
        const bool act = tile_active<MODE>(A, b, blockIdx.x == 0);
        const int j = A.iter % A.jlen[b];
        s_j[b] = act ? (j | (bin_checks(A, b) ? 0x10000 : 0)) : -1;   // bit 16: this bin measures the residual
        s_a[b] = A.a_bin[b];
        s_rho[b] = A.shift[(long long)b * A.jmax + j];
    }
}

// =========================================================================================================
// x sweep
// =========================================================================================================
// Tile = (bin, R = NT/QP consecutive rows), NT = 4096/S threads; smem stage = u tile with one halo row above and
// below + b tile, both as 128-byte-swizzled boxes (16 doubles | nx/16 | rows | 1 bin).  Thread (g, q) owns the
// S-cell chunk q of tile row g; the chunks of a row sit in QP adjacent lanes.
// SEG = false: whole lines per tile, separate output tiles (segment count, halo and output mode are compile-time
// constants: the short-line kernels carry none of the segment arithmetic).
// FULL: the tile fits the grid exactly (every lane owns a chunk of the row, every tile row exists): the validity
// predicates are compile-time true and none of the masking selects are generated.
template <int S, int QP, int NS, int NT, bool SEG, bool FULL = false>
__global__ void __launch_bounds__(NT, NT <= 128 ? 2 : 1)
k_sweep_x_pipe(PipeArgs Ain, const __grid_constant__ XMaps maps) {
    static_assert(!(SEG && FULL), "segmented tiles are never exact fits");
    const PipeArgs &A = Ain;
    const bool inplace = SEG && A.inplace != 0, delta = SEG && A.delta != 0;
    constexpr int R = NT / QP;
    constexpr int UPC = S / 2;           // 16-byte units per chunk
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int Q = A.Q;                   // chunks per row
    const int QS = A.qs, QI = SEG ? A.qi : A.qs, H = SEG ? A.halo : 0, nseg = SEG ? A.nseg : 1;
    const int Q16 = QS * S / 16;         // 128-byte units per tile row
    const int QI16 = QI * S / 16;        // ... of the interior (what is stored)
    const int u_bytes = ((R + 2) * Q16 * 128 + 1023) / 1024 * 1024;
    const int b_bytes = (R * Q16 * 128 + 1023) / 1024 * 1024;
    const int stage_bytes = u_bytes + b_bytes;
    // in-place mode: the result of a tile is written over its own b tile (dead once the right-hand side is formed)
    unsigned char *out_base = smraw + (size_t)NS * stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_base + (inplace ? 0 : 2 * (size_t)b_bytes));
    const int tid = threadIdx.x;
    const uint32_t full0 = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double *s_a = reinterpret_cast<double *>(bars + 8);
    double *s_rho = s_a + A.ne;
    int *s_j = reinterpret_cast<int *>(s_rho + A.ne);
    const int tpb = A.tiles_per_bin;
    const int g = tid / QP, q = tid - g * QP;
    const int nx = A.nx;
    // geometry of this thread's chunk (constant while the CTA stays on one block of lines)
    int cur_key = -1;
    bool qok = false, inter = false;
    int qc = 0;
    double cx[S], cy[S];
    unsigned flw[S / 4];
#pragma unroll
    for (int i = 0; i < S / 4; ++i) flw[i] = 0;
    int cls = 0;
    bool rowok = false;
    auto load_geometry = [&](int rem) {
        const int rblk = rem / nseg, sg = rem - rblk * nseg;
        const int y = rblk * R + g;
        if (y * nseg + sg == cur_key) return;
        cur_key = y * nseg + sg;
        const int qa = sg * QI - H + q;            // chunk of the row this thread solves
        qok = FULL || (q < QS && qa >= 0 && qa < Q);
        inter = FULL || (qok && q >= H && q < H + QI);       // ... and stores / measures
        qc = FULL ? q : min(max(qa, 0), Q - 1);
        rowok = FULL || (y < A.ny && qok);
        const int yc = min(y, A.ny - 1);
        cls = A.cls[yc];
        const size_t o = (size_t)yc * nx + qc * S;
        const unsigned *pf = reinterpret_cast<const unsigned *>(A.flags + o);
#pragma unroll
        for (int i = 0; i < S / 4; ++i) flw[i] = (FULL || (rowok && inter)) ? pf[i] : 0u;
        const double2 *px = reinterpret_cast<const double2 *>(A.cx + o);
        const double2 *py = reinterpret_cast<const double2 *>(A.cy + o);
#pragma unroll
        for (int un = 0; un < UPC; ++un) {
            const double2 a2 = px[un], b2 = py[un];
            cx[2 * un] = (FULL || rowok) ? a2.x : 0.0;
            cx[2 * un + 1] = (FULL || rowok) ? a2.y : 0.0;
            cy[2 * un] = (FULL || rowok) ? b2.x : 0.0;
            cy[2 * un + 1] = (FULL || rowok) ? b2.y : 0.0;
        }
    };
    // Static data first (the geometry of the CTA's first tile: with a grid that is a multiple of the tiles per bin it
    // is the geometry of all its tiles), then wait for the preceding grid of the stream: everything below reads what
    // it wrote (convergence flags, residuals, the state).
    if (!SEG) {   // segmented instances are launched in plain stream order (launch_pdl): they measured slower with it
        pdl_launch_dependents();
        if ((int)blockIdx.x < A.ntiles) load_geometry((int)blockIdx.x % tpb);
        pdl_wait();
    }
    load_bin_params<0, NT>(A, s_a, s_rho, s_j);
    __syncthreads();
    // producer cursor (thread 0 only): next tile index to load and number of loads issued
    int pt = blockIdx.x, pk = 0;
    auto produce = [&]() {
        while (pt < A.ntiles) {
            const int bin = pt / tpb;
            if (s_j[bin] >= 0) {
                const int rem = pt - bin * tpb;
                const int rblk = rem / nseg, sg = rem - rblk * nseg;
                const int y0 = rblk * R;
                const int c1 = (sg * QI - H) * S / 16;   // first 128-byte unit of the tile (negative: zero filled)
                const int s = pk % NS;
                const uint32_t dst = smem_u32(smraw + (size_t)s * stage_bytes);
                mbar_expect(full0 + 8 * s, (uint32_t)((R + 2) * Q16 * 128 + R * Q16 * 128));
                tma_load_4d(dst, &maps.u, full0 + 8 * s, 0, c1, y0 - 1, bin);
                tma_load_4d(dst + u_bytes, &maps.b, full0 + 8 * s, 0, c1, y0, bin);
                ++pk;
                pt += gridDim.x;
                return;
            }
            pt += gridDim.x;
        }
    };
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < NS - 1; ++i) produce();
    }
    const int ql = min(q, QS - 1);             // chunk slot inside the tile
    const int reach = A.depth + 1;           // an inclusive scan over offsets < reach covers `depth` earlier chunks
    const int r16 = (ql * S) >> 4;             // 128-byte unit of this chunk within its tile row
    const int ubase = ((ql * S) & 15) >> 1;    // first 16-byte unit of the chunk inside that 128-byte unit
    const int r16o = r16 - ((H * S) >> 4);     // ... within the stored interior
    int k = 0;
    for (int t = blockIdx.x; t < A.ntiles; t += gridDim.x) {
        const int bin = t / tpb;
        const int jraw = s_j[bin];
        if (jraw < 0) continue;
        const int jidx = jraw & 0xffff;
        const bool chk = (jraw & 0x10000) != 0;
        const int rem = t - bin * tpb;
        const int rblk = rem / nseg, sg = rem - rblk * nseg;
        const int y0 = rblk * R;
        // separate output tiles: every thread finished reading the stage of tile k-1 before the last named barrier of
        // that iteration, so it is refilled right away
        if (tid == 0 && !inplace) produce();
        load_geometry(rem);   // no-op while the CTA stays on its block of lines
        const double a = s_a[bin];
        const double rho = s_rho[bin];
        const bool qok_c = FULL ? true : qok, rowok_c = FULL ? true : rowok, inter_c = FULL ? true : inter;
        ChunkSolve<S> ch;
        double2 mR[UPC], gR[UPC];
        {   // LU factors of this chunk (L2 / L1 resident table): issued before the wait on the tile, first touched (masked)
            // after the right-hand side has been formed
            const size_t base = (((size_t)bin * A.jmax + jidx) * A.nclass + cls) * A.npad;
            const double2 *pm = reinterpret_cast<const double2 *>(A.tabm + base);
            const double2 *pg = reinterpret_cast<const double2 *>(A.tabg + base);
#pragma unroll
            for (int un = 0; un < UPC; ++un) {
                mR[un] = pm[un * Q + qc];
                gR[un] = pg[un * Q + qc];
            }
            if (SEG) {   // segmented tiles: masked at once (measured: the late touch costs 3 % there)
#pragma unroll
                for (int un = 0; un < UPC; ++un) {
                    ch.m[2 * un] = rowok_c ? mR[un].x : 0.0;
                    ch.m[2 * un + 1] = rowok_c ? mR[un].y : 0.0;
                    ch.g[2 * un] = rowok_c ? gR[un].x : 0.0;
                    ch.g[2 * un + 1] = rowok_c ? gR[un].y : 0.0;
                }
            }
        }
        if (A.prefetch) {
            // The tables are L2 resident, but an L2 round trip per tile sat exposed in front of the solve (ncu: the first
            // use of the factors).  The CTA stays on its block of lines, so the next tile differs only in the bin: its
            // rows of the tables are requested into L1 a whole tile ahead (2 lines per thread, no registers held).
            int tn = t + gridDim.x;
            while (tn < A.ntiles && s_j[tn / tpb] < 0) tn += gridDim.x;
            if (tn < A.ntiles) {
                const int bn = tn / tpb;
                const size_t basen = (((size_t)bn * A.jmax + (s_j[bn] & 0xffff)) * A.nclass + cls) * A.npad;
                // lane q of a row covers unit q / 2 ... : UPC units x Q lanes x 16 bytes per row and table = UPC*Q/8 lines
                const int nline = UPC * Q / 8;
                for (int l = q; l < nline; l += QP) {
                    prefetch_l1(A.tabm + basen + (size_t)l * 16);
                    prefetch_l1(A.tabg + basen + (size_t)l * 16);
                }
            }
        }
        const int s = k % NS;
        mbar_wait(full0 + 8 * s, (k / NS) & 1);
        const double *su = reinterpret_cast<const double *>(smraw + (size_t)s * stage_bytes);
        const double *sb = reinterpret_cast<const double *>(smraw + (size_t)s * stage_bytes + u_bytes);
        int rhi = 0, uhi = 0;
        {
            const int rb = g * Q16 + r16;            // 128-byte row index in the b tile
            const int ru = (g + 1) * Q16 + r16;      // in the u tile (one halo row on top)
            const double2 *pb = reinterpret_cast<const double2 *>(sb + (size_t)rb * 16);
            const double2 *pc = reinterpret_cast<const double2 *>(su + (size_t)ru * 16);
            const double2 *pu = reinterpret_cast<const double2 *>(su + (size_t)(ru - Q16) * 16);
            const double2 *pd = reinterpret_cast<const double2 *>(su + (size_t)(ru + Q16) * 16);
            const int swb = rb & 7, swc = ru & 7, swu = (ru - Q16) & 7, swd = (ru + Q16) & 7;
            double uc[S];
#pragma unroll
            for (int un = 0; un < UPC; ++un) {
                const double2 tq = pc[(ubase + un) ^ swc];
                uc[2 * un] = qok_c ? tq.x : 0.0;
                uc[2 * un + 1] = qok_c ? tq.y : 0.0;
            }
            double ul = __shfl_up_sync(0xffffffffu, uc[S - 1], 1, QP);
            double ur = __shfl_down_sync(0xffffffffu, uc[0], 1, QP);
            if (q == 0) ul = 0.0;
            if (q == QP - 1) ur = 0.0;
            const double rm = rho - 0.5, rp = rho + 0.5;
            const double tolb = chk ? A.tol[bin] : 0.0;
#pragma unroll
            for (int un = 0; un < UPC; ++un) {
                const double2 tu = pu[(ubase + un) ^ swu], td = pd[(ubase + un) ^ swd], tb = pb[(ubase + un) ^ swb];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int tt = 2 * un + h;
                    const double u0 = uc[tt];
                    const double uu = h == 0 ? tu.x : tu.y, ud = h == 0 ? td.x : td.y, bv = h == 0 ? tb.x : tb.y;
                    const double left = tt == 0 ? ul : uc[tt - 1], right = tt == S - 1 ? ur : uc[tt + 1];
                    const double cross = fma(cy[tt], u0, -uu) - ud;
                    const double along = fma(cx[tt], u0, -left) - right;
                    const double d = fma(-a, cross, fma(rm, u0, bv));      // b - (V - rho) u
                    ch.v[tt] = d;
                    if (chk) {
                        // componentwise stop test (Oettli-Prager): |b - A u|_i <= tol (|A||u| + |b|)_i in every cell; what
                        // is recorded is the largest excess over that bound, 0 when the bin has converged
                        const double rres = fma(-a, along, fma(-rp, u0, d));   // b - A u
                        const double nb = (fabs(left) + fabs(right)) + (fabs(uu) + fabs(ud));
                        const double wgt = fabs(bv) + fma(a, nb, fma(a, cx[tt] + cy[tt], 1.0) * fabs(u0));
                        const double ex = fma(-tolb, wgt, fabs(rres));
                        const int rh = ex > 0.0 ? __double2hiint(ex) : 0;
                        if (flw[tt >> 2] & (16u << (8 * (tt & 3)))) rhi = max(rhi, rh);
                        uhi = max(uhi, __double2hiint(u0) & 0x7fffffff);
                    }
                }
            }
        }
        if (chk) {
            rhi = __reduce_max_sync(0xffffffffu, rhi);
            uhi = __reduce_max_sync(0xffffffffu, uhi);
            if ((tid & 31) == 0) {
                // high words only: the residual norm is rounded up, the solution norm down (both conservative)
                atomicMax(&A.res[(long long)A.iter * A.ne + bin], ((unsigned long long)(unsigned)(rhi + (rhi ? 1 : 0))) << 32);
                atomicMax(&A.unorm[(long long)A.iter * A.ne + bin], ((unsigned long long)(unsigned)uhi) << 32);
            }
        }
        // Refill the stage of tile k-1 with tile k+NS-1: every thread finished with it before the last named barrier of
        // the previous iteration, and by now (a right-hand side later) its store has read the shared memory.
        if (tid == 0 && inplace) {
            bulk_wait_read<0>();
            produce();
        }
        if (!SEG) {
#pragma unroll
            for (int un = 0; un < UPC; ++un) {
                ch.m[2 * un] = rowok_c ? mR[un].x : 0.0;
                ch.m[2 * un + 1] = rowok_c ? mR[un].y : 0.0;
                ch.g[2 * un] = rowok_c ? gR[un].x : 0.0;
                ch.g[2 * un + 1] = rowok_c ? gR[un].y : 0.0;
            }
        }
        double Am, Bm;
        ch.forward(Am, Bm);
        const double yin = warp_carry<QP, false>(Am, Bm, q, reach);
        ch.forward_fix(yin);
        Bm = ch.backward();
        const double xin = warp_carry<QP, true>(Am, Bm, q, reach);
        ch.backward_fix(xin);
        // ---- out tile ----
        double *so = inplace ? const_cast<double *>(sb)
                               : reinterpret_cast<double *>(out_base + (size_t)(k & 1) * b_bytes);
        // in place: everybody has formed its right-hand side; else: thread 0 has seen the store of tile k-2 finish reading
        cta_bar<NT>(1);
        if (inter_c) {
            const int rb = g * QI16 + r16o;
            double2 *dst = reinterpret_cast<double2 *>(so + (size_t)rb * 16);
            const int swb = rb & 7;
            if (delta) {
                const int ru = (g + 1) * Q16 + r16;
                const double2 *pc = reinterpret_cast<const double2 *>(su + (size_t)ru * 16);
                const int swc = ru & 7;
#pragma unroll
                for (int un = 0; un < UPC; ++un) {
                    const double2 u2 = pc[(ubase + un) ^ swc];
                    dst[(ubase + un) ^ swb] = make_double2(ch.v[2 * un] - u2.x, ch.v[2 * un + 1] - u2.y);
                }
            } else {
#pragma unroll
                for (int un = 0; un < UPC; ++un) dst[(ubase + un) ^ swb] = make_double2(ch.v[2 * un], ch.v[2 * un + 1]);
            }
        }
        fence_async_smem();
        cta_bar<NT>(2);
        if (tid == 0) {
            tma_store_4d(&maps.out, smem_u32(so), 0, sg * QI * S / 16, y0, bin);
            bulk_commit();
            if (!inplace) bulk_wait_read<1>();
        }
        ++k;
    }
    if (tid == 0) bulk_wait_read<0>();
}

// =========================================================================================================
// y sweep
// =========================================================================================================
// Tile = (bin, strip of CW columns, all rows), NT = 4096/S threads.  smem stage = u strip + u* strip as
// [npad rows][CW] (no swizzle: the CW lanes of a row read one contiguous segment).  Thread (q, c): chunk q (S rows) of
// column c.  
template <int S, int CW, int NS, int NT, bool SEG, bool FULL = false>
__global__ void __launch_bounds__(NT, NT <= 128 ? 2 : 1)
k_sweep_y_pipe(PipeArgs Ain, const __grid_constant__ YMaps maps) {
    static_assert(!(SEG && FULL), "segmented tiles are never exact fits");
    const PipeArgs &A = Ain;
    const bool inplace = SEG && A.inplace != 0, delta = SEG && A.delta != 0;
    constexpr int NCH = NT / CW;      // chunk slots per column
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int Q = A.Q;                // chunks per column
    const int QS = A.qs, QI = SEG ? A.qi : A.qs, H = SEG ? A.halo : 0, nseg = SEG ? A.nseg : 1;
    const int npad = A.npad;          // rows of a factor table
    const int trows = QS * S;         // rows of a tile (halo included)
    const int strip_bytes = (trows * CW * 8 + 127) / 128 * 128;
    const int stage_bytes = 2 * strip_bytes;
    // in-place mode: the new u of a tile is written over its own u* strip (every thread has both strips of its chunk in
    // registers by then)
    unsigned char *out_base = smraw + (size_t)NS * stage_bytes;
    // chunk carries: (A, forward B, backward B) x [NCH][CW], one set per tile parity: a tile publishes into the set the
    // tile before the previous one used, so publishing needs no barrier of its own (three CTA barriers per tile)
    double *carry = reinterpret_cast<double *>(out_base + (inplace ? 0 : 2 * (size_t)strip_bytes));   // [2][3][NCH][CW]
    uint64_t *bars = reinterpret_cast<uint64_t *>(carry + 6 * NCH * CW);
    const int tid = threadIdx.x;
    const uint32_t full0 = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double *s_a = reinterpret_cast<double *>(bars + 8);
    double *s_rho = s_a + A.ne;
    int *s_j = reinterpret_cast<int *>(s_rho + A.ne);
    if (!SEG) {
        pdl_launch_dependents();
        pdl_wait();   // everything below reads what the preceding grid wrote
    }
    load_bin_params<1, NT>(A, s_a, s_rho, s_j);
    __syncthreads();
    const int tpb = A.tiles_per_bin;
    const int nbox = (trows + 255) / 256;         // TMA boxes per strip (box rows <= 256)
    const int box_rows = trows / nbox;            // host guarantees divisibility
    const int orows = QI * S;                     // stored rows
    const int nbox_o = (orows + 255) / 256;
    const int box_rows_o = orows / nbox_o;
    // tile of a position in the walking order (rev: from the last bin down, what the preceding sweep wrote last)
    auto tile_of = [&](int pos) { return (!SEG && A.rev) ? A.ntiles - 1 - pos : pos; };
    int pt = blockIdx.x, pk = 0;
    auto produce = [&]() {
        while (pt < A.ntiles) {
            const int ptile = tile_of(pt);
            const int bin = ptile / tpb;
            if (s_j[bin] >= 0) {
                const int rem = ptile - bin * tpb;
                const int strip = rem / nseg, sg = rem - strip * nseg;
                const int x0 = strip * CW;
                const int row0 = (sg * QI - H) * S;     // negative / beyond the grid: zero filled
                const int s = pk % NS;
                const uint32_t dst = smem_u32(smraw + (size_t)s * stage_bytes);
                mbar_expect(full0 + 8 * s, (uint32_t)(2 * trows * CW * 8));
                for (int bx = 0; bx < nbox; ++bx) {
                    tma_load_3d(dst + bx * box_rows * CW * 8, &maps.u, full0 + 8 * s, x0, row0 + bx * box_rows, bin);
                    tma_load_3d(dst + strip_bytes + bx * box_rows * CW * 8, &maps.w, full0 + 8 * s, x0, row0 + bx * box_rows, bin);
                }
                ++pk;
                pt += gridDim.x;
                return;
            }
            pt += gridDim.x;
        }
    };
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < NS - 1; ++i) produce();
    }
    const int q = tid / CW, c = tid - q * CW;
    const int r0 = (q < QS ? q : 0) * S;            // first row of this thread's chunk inside the tile
    bool qok = false, inter = false;
    auto next_tile = [&](int t) {
        while (t < A.ntiles && s_j[tile_of(t) / tpb] < 0) t += gridDim.x;
        return t;
    };
    double2 mR[S / 2], gR[S / 2];   // raw factor loads of the current tile (masked after the wait on the tile)
    int cur_strip = -1, cls = 0;
    auto factor_base = [&](int pos) {   // table offset of this thread's chunk for the tile at `pos`; sets qok / inter
        const int t = tile_of(pos);
        const int bin = t / tpb;
        const int rem = t - bin * tpb;
        const int strip = rem / nseg, sg = rem - strip * nseg;
        const int qa = sg * QI - H + q;               // chunk of the column this thread solves
        qok = q < QS && qa >= 0 && qa < Q;
        inter = qok && q >= H && q < H + QI;
        if (SEG || strip != cur_strip) {   // the CTA normally stays on one strip: the class lookup is not on the tile's path
            cur_strip = strip;
            cls = A.cls[min(strip * CW + c, A.nx - 1)];
        }
        return (((size_t)bin * A.jmax + (s_j[bin] & 0xffff)) * A.nclass + cls) * npad + (size_t)min(max(qa, 0), Q - 1) * S;
    };
    auto load_factors = [&](int t) {   // issue the loads of tile t's factors
        const size_t base = factor_base(t);
        const double2 *pm = reinterpret_cast<const double2 *>(A.tabm + base);
        const double2 *pg = reinterpret_cast<const double2 *>(A.tabg + base);
#pragma unroll
        for (int un = 0; un < S / 2; ++un) {
            mR[un] = pm[un];
            gR[un] = pg[un];
        }
    };
    int k = 0;
    int t = next_tile(blockIdx.x);
    while (t < A.ntiles) {
        const int tile = tile_of(t);
        const int bin = tile / tpb;
        const int rem = tile - bin * tpb;
        const int strip = rem / nseg, sg = rem - strip * nseg;
        const int x0 = strip * CW;
        const int tn = next_tile(t + gridDim.x);
        if (tid == 0 && !inplace) produce();
        const double rho2 = 2.0 * s_rho[bin];
        ChunkSolve<S> ch;
        load_factors(t);   // issued before the wait on the tile (a register prefetch of the next tile measured slower)
        const bool qok_t = FULL ? true : qok, inter_t = FULL ? true : inter;
        if (SEG) {   // segmented tiles: masked at once (see the x sweep)
#pragma unroll
            for (int un = 0; un < S / 2; ++un) {
                ch.m[2 * un] = qok_t ? mR[un].x : 0.0;
                ch.m[2 * un + 1] = qok_t ? mR[un].y : 0.0;
                ch.g[2 * un] = qok_t ? gR[un].x : 0.0;
                ch.g[2 * un + 1] = qok_t ? gR[un].y : 0.0;
            }
        }
        if (A.prefetch && tn < A.ntiles) {   // next tile's table lines into L1 (see the x sweep); one line per table
            const size_t basen = factor_base(tn);
            prefetch_l1(A.tabm + basen);
            prefetch_l1(A.tabg + basen);
        }
        const int s = k % NS;
        mbar_wait(full0 + 8 * s, (k / NS) & 1);
        const double *su = reinterpret_cast<const double *>(smraw + (size_t)s * stage_bytes) + (size_t)r0 * CW + c;
        const double *sw = reinterpret_cast<const double *>(smraw + (size_t)s * stage_bytes + strip_bytes) + (size_t)r0 * CW + c;
        double uold[S];
#pragma unroll
        for (int tt = 0; tt < S; ++tt) {
            uold[tt] = su[tt * CW];
            ch.v[tt] = delta ? sw[tt * CW] : sw[tt * CW] - uold[tt];
        }
        // the factors are first touched here, behind the wait and the tile reads: their L2 latency is off the critical path
        if (!SEG) {
#pragma unroll
            for (int un = 0; un < S / 2; ++un) {
                ch.m[2 * un] = qok_t ? mR[un].x : 0.0;
                ch.m[2 * un + 1] = qok_t ? mR[un].y : 0.0;
                ch.g[2 * un] = qok_t ? gR[un].x : 0.0;
                ch.g[2 * un + 1] = qok_t ? gR[un].y : 0.0;
            }
        }
        // refill the stage of tile k-1 with tile k+NS-1 (its store, issued an iteration ago, has read the shared memory)
        if (tid == 0 && inplace) {
            bulk_wait_read<0>();
            produce();
        }
        double Am, Bm;
        ch.forward(Am, Bm);
        double *cA = carry + (size_t)(SEG ? 0 : (k & 1)) * 3 * NCH * CW, *cB = cA + NCH * CW, *cR = cB + NCH * CW;
        if (SEG) cta_bar<NT>(1);   // segmented tiles keep the five-barrier schedule (measured faster there)
        cA[q * CW + c] = Am;
        cB[q * CW + c] = Bm;
        cta_bar<NT>(2);
        double cin = 0.0;
        for (int kk = max(0, q - A.depth); kk < q; ++kk) cin = fma(cA[kk * CW + c], cin, cB[kk * CW + c]);
        ch.forward_fix(cin);
        Bm = ch.backward();
        if (SEG) cta_bar<NT>(3);
        cR[q * CW + c] = Bm;
        cta_bar<NT>(4);
        cin = 0.0;
        for (int kk = min(NCH - 1, q + A.depth); kk > q; --kk) cin = fma(cA[kk * CW + c], cin, cR[kk * CW + c]);
        ch.backward_fix(cin);
        unsigned char *obuf = inplace ? smraw + (size_t)s * stage_bytes + strip_bytes
                                        : out_base + (size_t)(k & 1) * strip_bytes;
        double *so = reinterpret_cast<double *>(obuf) + (size_t)r0 * CW + c;   // in place: this thread's own u* chunk
        if (inter_t) {
#pragma unroll
            for (int tt = 0; tt < S; ++tt) so[tt * CW] = fma(rho2, ch.v[tt], uold[tt]);
        }
        fence_async_smem();
        cta_bar<NT>(5);
        if (tid == 0) {
            const uint32_t src = smem_u32(obuf) + H * S * CW * 8;
            for (int bx = 0; bx < nbox_o; ++bx)
                tma_store_3d(&maps.out, src + bx * box_rows_o * CW * 8, x0, sg * orows + bx * box_rows_o, bin);
            bulk_commit();
            if (!inplace) bulk_wait_read<1>();
        }
        ++k;
        t = tn;
    }
    if (tid == 0) bulk_wait_read<0>();
}

// ---- tensor maps -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// x: 4-D view (16 | nx/16 | ny | ne) of a dense [ne][ny][nx] array, box (16 | Q | rows | 1), 128-byte swizzle
bool make_xmap(CUtensorMap *m, double *base, int ne, int ny, int nx, int Q, int rows) {
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

// y: 3-D view (nx | ny | ne), box (CW | rows | 1), no swizzle
bool make_ymap(CUtensorMap *m, double *base, int ne, int ny, int nx, int cw, int rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)ne};
    const cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)ny * nx * 8};
    const cuuint32_t box[3] = {(cuuint32_t)cw, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

constexpr int SMEM_CAP = 227 * 1024;

size_t param_smem(int ne) { return (size_t)ne * 20 + 16; }

// x: S cells per chunk, QP lanes per row, nx16 = nx/16
size_t x_smem(int nt, int QP, int nx16, int ns, bool inplace) {
    const int R = nt / QP;
    const size_t ub = ((size_t)(R + 2) * nx16 * 128 + 1023) / 1024 * 1024;
    const size_t bb = ((size_t)R * nx16 * 128 + 1023) / 1024 * 1024;
    return ns * (ub + bb) + (inplace ? 0 : 2 * bb) + 64;
}

// Launch with programmatic stream serialization (see pdl_wait): consecutive sweep kernels overlap the tail of one
// with the prologue of the next.  QPB_PIPE_PDL=0: plain stream order.
template <class Kern, class Maps>
int launch_pdl(Kern kern, int grid, int nt, size_t smem, qpb_ctx *c, const PipeArgs &A, const Maps &maps, bool pdl) {
    const char *e = getenv("QPB_PIPE_PDL");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)nt);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (!pdl || (e && e[0] == '0')) ? 0 : 1;
    QPB_CUDA(cudaLaunchKernelEx(&cfg, kern, A, maps));
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

template <int S, int QP, int NS, int NT, bool SEG, bool FULL = false>
int launch_x_seg(qpb_ctx *c, const PipeArgs &A, const XMaps &maps, int grid) {
    auto kern = k_sweep_x_pipe<S, QP, NS, NT, SEG, FULL>;
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP));
        configured = true;
    }
    return launch_pdl(kern, grid, NT, x_smem(NT, QP, A.qs * S / 16, NS, A.inplace != 0) + param_smem(A.ne), c, A, maps, !SEG);
}

template <int S, int QP, int NS, int NT>
int launch_x(qpb_ctx *c, const PipeArgs &A, const XMaps &maps, int grid) {
    const bool seg = A.nseg > 1 || A.inplace || A.delta;
    if (S == 16 && seg) return launch_x_seg<S, QP, NS, NT, (S == 16)>(c, A, maps, grid);
    if (seg) {
        qpb_set_error("segmented x sweep planned for an unsupported tile shape");
        return QPB_E_INVALID;
    }
    // exact fit: every lane owns a chunk (QP == chunks per row == chunks per tile) and every tile row exists
    if (S == 16 && A.qs == QP && A.Q == QP && A.ny % (NT / QP) == 0 && !getenv("QPB_PIPE_NOFULL"))
        return launch_x_seg<S, QP, NS, NT, false, (S == 16)>(c, A, maps, grid);
    return launch_x_seg<S, QP, NS, NT, false>(c, A, maps, grid);
}

size_t y_smem(int nt, int cw, int trows, int ns, bool inplace) {
    const size_t sb = ((size_t)trows * cw * 8 + 127) / 128 * 128;
    return ns * 2 * sb + (inplace ? 0 : 2 * sb) + sizeof(double) * 6 * nt + 64;
}

template <int S, int CW, int NS, int NT, bool SEG, bool FULL = false>
int launch_y_seg(qpb_ctx *c, const PipeArgs &A, const YMaps &maps, int grid) {
    auto kern = k_sweep_y_pipe<S, CW, NS, NT, SEG, FULL>;
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP));
        configured = true;
    }
    return launch_pdl(kern, grid, NT, y_smem(NT, CW, A.qs * S, NS, A.inplace != 0) + param_smem(A.ne), c, A, maps, !SEG);
}

template <int S, int CW, int NS, int NT>
int launch_y(qpb_ctx *c, const PipeArgs &A, const YMaps &maps, int grid) {
    const bool seg = A.nseg > 1 || A.inplace || A.delta;
    if (S == 16 && CW <= 8 && seg) return launch_y_seg<S, CW, NS, NT, (S == 16 && CW <= 8)>(c, A, maps, grid);
    if (seg) {
        qpb_set_error("segmented y sweep planned for an unsupported tile shape");
        return QPB_E_INVALID;
    }
    // exact fit: the chunk slots of the CTA are the chunks of a column and the strips tile the grid
    if (S == 16 && A.qs == NT / CW && A.Q == NT / CW && A.nx % CW == 0 && !getenv("QPB_PIPE_NOFULL"))
        return launch_y_seg<S, CW, NS, NT, false, (S == 16)>(c, A, maps, grid);
    return launch_y_seg<S, CW, NS, NT, false>(c, A, maps, grid);
}

int pick_grid(int ntiles, int tpb, int nsm) {
    if (ntiles <= nsm) return ntiles;
    // a multiple of the tiles per bin keeps every CTA on one block of lines (its geometry stays in registers)
    if (tpb <= nsm) return (nsm / tpb) * tpb;
    return nsm;
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------------
// chunk length the pipelined kernels want for a line of n cells (0: not supported)
int qpbp_chunk(int n, int dir) {
    (void)dir;
    // S = 16 (256 threads, ~250 registers each) measured faster than S = 8 (512 threads) on C2: 43/40 us vs 59/58 us
    // per x/y sweep; QPB_PIPE_S=8 selects the short chunks for experiments
    const char *e = getenv("QPB_PIPE_S");
    if (e && e[0] == '8' && n <= 256) return 8;
    return 16;   // lines longer than 512 cells are solved in overlapping segments of 32 chunks
}

// segments of a line of Q chunks whose carries reach `depth` chunks: false when the halo would eat the tile
static bool plan_segments(int Q, int depth, int &qs, int &qi, int &halo, int &nseg) {
    if (Q <= 32) {
        qs = qi = Q;
        halo = 0;
        nseg = 1;
        return true;
    }
    halo = std::max(1, depth);
    if (const char *e = getenv("QPB_HALO_EXTRA")) halo += atoi(e);
    if (getenv("QPB_DEBUG_RES")) fprintf(stderr, "[qpb] segments: Q %d depth %d halo %d\n", Q, depth, halo);
    const int qi_max = 32 - 2 * halo;      // a tile has at most 32 chunks (one warp scans a row)
    if (qi_max < 8) return false;
    nseg = (Q + qi_max - 1) / qi_max;
    qi = (Q + nseg - 1) / nseg;             // equal segments: the last tile is not mostly padding
    qs = qi + 2 * halo;
    return true;
}

int qpbp_plan(qpb_ctx *c, DiffSlot &s, PipePlan &p) {
    p = PipePlan();
    const auto &cf = c->cfg;
    if (getenv("QPB_NO_PIPE") && getenv("QPB_NO_PIPE")[0] == '1') return QPB_OK;
    if (s.mode != 0 || !s.fast || !encode_fn() || cf.ne > 2048) return QPB_OK;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    p.nsm = nsm;
    // threads per CTA: 4096 / S, one CTA per SM (two CTAs of 128 threads per SM measured no faster: the kernels are
    // bound by the bytes they move, not by the latency of one tile's recurrences)
    p.nt = 4096 / std::max(qpbp_chunk(cf.nx, 0), 8);
    const size_t smem_cap = (size_t)SMEM_CAP;
    // measured (profiles/): short lines run fastest with separate output tiles and the deepest ring that fits,
    // segmented long lines with in-place output and two stages
    int max_ns = 4;
    if (const char *e = getenv("QPB_PIPE_NS")) max_ns = std::max(2, std::min(4, atoi(e)));
    int force_inplace = -1;
    if (const char *e = getenv("QPB_PIPE_INPLACE")) force_inplace = atoi(e) != 0;
    // x sweep: rows cut into S-cell chunks owned by adjacent lanes (<= 32 chunks), TMA boxes of whole 128-byte units
    int qs = 0, qi = 0, halo = 0, nseg = 1;
    if (cf.nx % 16 == 0 && s.fx.d_tabg && s.fx.S == qpbp_chunk(cf.nx, 0) && s.fx.S > 0 &&
        (s.fx.Q <= 32 || s.fx.S == 16) && plan_segments(s.fx.Q, s.fx.carry_depth, qs, qi, halo, nseg)) {
        const int S = s.fx.S, QP = std::max(4, next_pow2(qs)), R = p.nt / QP;
        const int u16 = qs * S / 16, o16 = qi * S / 16;   // 128-byte units per tile row: loaded, stored
        const bool inplace = force_inplace >= 0 ? force_inplace != 0 : nseg > 1;
        int ns = 0;
        for (int cand = (inplace && force_inplace < 0 && !getenv("QPB_PIPE_NS")) ? 2 : max_ns; cand >= 2 && !ns; --cand)
            if (x_smem(p.nt, QP, u16, cand, inplace) + param_smem(cf.ne) <= smem_cap) ns = cand;
        p.x_inplace = inplace;
        XMaps maps;
        if (QP <= 32 && ns && make_xmap(&maps.u, c->d_S, cf.ne, cf.ny, cf.nx, u16, R + 2) &&
            make_xmap(&maps.b, c->d_B, cf.ne, cf.ny, cf.nx, u16, R) &&
            make_xmap(&maps.out, c->d_T1, cf.ne, cf.ny, cf.nx, o16, R)) {
            p.x_ok = true;
            p.x_qp = QP;
            p.x_ns = ns;
            p.x_qs = qs; p.x_qi = qi; p.x_halo = halo; p.x_nseg = nseg;
            p.x_tpb = ((cf.ny + R - 1) / R) * nseg;
            p.xmaps.resize(sizeof(XMaps));
            memcpy(p.xmaps.data(), &maps, sizeof(XMaps));
        }
    }
    // y sweep: strips of CW columns, columns cut into S-row chunks (<= NT/CW chunks)
    if (cf.nx % 2 == 0 && s.fy.d_tabg && s.fy.S == qpbp_chunk(cf.ny, 1) && s.fy.S > 0 &&
        plan_segments(s.fy.Q, s.fy.carry_depth, qs, qi, halo, nseg)) {
        const int S = s.fy.S, NT = p.nt;
        const int trows = qs * S, orows = qi * S;     // rows of a tile: loaded, stored
        int cw = std::min(32, NT / next_pow2(qs));
        while (cw > 4 && cw / 2 >= cf.nx) cw /= 2;    // narrow grids: do not load columns that do not exist
        const int nbox = (trows + 255) / 256, nbox_o = (orows + 255) / 256;
        if (cw >= 4 && trows % nbox == 0 && orows % nbox_o == 0 && ((trows / nbox) * cw) % 16 == 0 &&
            ((orows / nbox_o) * cw) % 16 == 0 && (halo * S * cw * 8) % 128 == 0) {
            const bool inplace = force_inplace >= 0 ? force_inplace != 0 : nseg > 1;
            int ns = 0;
            for (int cand = (inplace && force_inplace < 0 && !getenv("QPB_PIPE_NS")) ? 2 : max_ns; cand >= 2 && !ns; --cand)
                if (y_smem(NT, cw, trows, cand, inplace) + param_smem(cf.ne) <= smem_cap) ns = cand;
            p.y_inplace = inplace;
            YMaps maps;
            const int rows = trows / nbox;
            if (ns && make_ymap(&maps.u, c->d_S, cf.ne, cf.ny, cf.nx, cw, rows) &&
                make_ymap(&maps.w, c->d_T1, cf.ne, cf.ny, cf.nx, cw, rows) &&
                make_ymap(&maps.out, c->d_S, cf.ne, cf.ny, cf.nx, cw, orows / nbox_o)) {
                p.y_ok = true;
                p.y_cw = cw;
                p.y_ns = ns;
                p.y_qs = qs; p.y_qi = qi; p.y_halo = halo; p.y_nseg = nseg;
                p.y_tpb = ((cf.nx + cw - 1) / cw) * nseg;
                p.ymaps.resize(sizeof(YMaps));
                memcpy(p.ymaps.data(), &maps, sizeof(YMaps));
            }
        }
    }
    return QPB_OK;
}

template <int S, int NT>
static int dispatch_x(qpb_ctx *c, const PipeArgs &A, const XMaps &maps, int grid, int qp, int ns) {
#define QPB_X(QP)                                                      \
    case QP:                                                           \
        return ns >= 4 ? launch_x<S, QP, 4, NT>(c, A, maps, grid)                                    \
               : ns == 3 ? launch_x<S, QP, 3, NT>(c, A, maps, grid) : launch_x<S, QP, 2, NT>(c, A, maps, grid);
    switch (qp) {
        QPB_X(4) QPB_X(8) QPB_X(16)
        default:
            return ns >= 4 ? launch_x<S, 32, 4, NT>(c, A, maps, grid)
                   : ns == 3 ? launch_x<S, 32, 3, NT>(c, A, maps, grid) : launch_x<S, 32, 2, NT>(c, A, maps, grid);
    }
#undef QPB_X
}

template <int S, int NT>
static int dispatch_y(qpb_ctx *c, const PipeArgs &A, const YMaps &maps, int grid, int cw, int ns) {
#define QPB_Y(CW)                                                      \
    case CW:                                                           \
        return ns >= 4 ? launch_y<S, CW, 4, NT>(c, A, maps, grid)                                    \
               : ns == 3 ? launch_y<S, CW, 3, NT>(c, A, maps, grid) : launch_y<S, CW, 2, NT>(c, A, maps, grid);
    switch (cw) {
        QPB_Y(4) QPB_Y(8) QPB_Y(16)
        default:
            return ns >= 4 ? launch_y<S, 32, 4, NT>(c, A, maps, grid)
                   : ns == 3 ? launch_y<S, 32, 3, NT>(c, A, maps, grid) : launch_y<S, 32, 2, NT>(c, A, maps, grid);
    }
#undef QPB_Y
}

int qpbp_sweep(qpb_ctx *c, DiffSlot &s, int dir, int iter, bool check) {
    const auto &cf = c->cfg;
    const PipePlan &p = s.pipe;
    const DiffSlot::FastDir &fd = dir == 0 ? s.fx : s.fy;
    PipeArgs A;
    A.ne = cf.ne; A.ny = cf.ny; A.nx = cf.nx; A.iter = iter; A.jmax = s.jmax; A.tol = s.d_tol;
    A.cx = c->d_cx; A.cy = c->d_cy; A.flags = c->d_flags;
    A.a_bin = s.d_a; A.shift = s.d_shift; A.jlen = s.d_jlen;
    A.cls = fd.d_cls;
    A.tabm = fd.d_tab; A.tabg = fd.d_tabg; A.nclass = fd.nclass; A.npad = fd.npad; A.Q = fd.Q;
    A.res = c->d_res; A.unorm = c->d_unorm; A.done = c->d_done; A.iters_out = c->d_done + cf.ne;
    A.depth = std::max(1, fd.carry_depth);
    A.delta = (p.x_ok && p.y_ok && p.y_nseg > 1) ? 1 : 0;   // only a segmented y sweep reads rows it does not own
    A.check_all = check ? 1 : 0;
    {   // L2 reuse between consecutive sweeps: the x sweep walks the bins upwards, the y sweep downwards, so each starts
        // with what its predecessor wrote last (126 MB of L2 against 3 x 64 MiB of state and work arrays at C2:
        // y sweep 42.1 -> 40.5 us).  An evict_last policy on the right-hand side tiles measured no gain.
        const char *e = getenv("QPB_PIPE_REV");
        A.rev = (dir == 1 && !(e && e[0] == '0')) ? 1 : 0;
    }
    {   // measured at C2: y sweep 47.2 -> 42.7 us per launch, x sweep unchanged; QPB_PIPE_PREFETCH=0 switches it off
        const char *e = getenv("QPB_PIPE_PREFETCH");
        A.prefetch = !(e && e[0] == '0');
    }
    A.known = s.d_known;
    ScopedTimer tm(c, dir == 0 ? 0 : 1);
    c->diag.kernel_launches++;
    A.qs = dir == 0 ? p.x_qs : p.y_qs;
    A.qi = dir == 0 ? p.x_qi : p.y_qi;
    A.halo = dir == 0 ? p.x_halo : p.y_halo;
    A.nseg = dir == 0 ? p.x_nseg : p.y_nseg;
    A.inplace = (dir == 0 ? p.x_inplace : p.y_inplace) ? 1 : 0;
    if (dir == 0) {
        A.tiles_per_bin = p.x_tpb;
        A.ntiles = p.x_tpb * cf.ne;
        const int grid = pick_grid(A.ntiles, p.x_tpb, p.nsm);
        // the prefetch addresses assume that a CTA's next tile differs from this one only in the bin (whole lines, grid a
        // multiple of the tiles per bin); on segmented or wandering tilings it fetched the wrong rows and cost 4 % at
        // 2048^2
        if (A.nseg > 1 || grid % p.x_tpb != 0) A.prefetch = 0;
        const XMaps &maps = *reinterpret_cast<const XMaps *>(p.xmaps.data());
        if (fd.S == 8) return dispatch_x<8, 512>(c, A, maps, grid, p.x_qp, p.x_ns);
        return dispatch_x<16, 256>(c, A, maps, grid, p.x_qp, p.x_ns);
    }
    A.tiles_per_bin = p.y_tpb;
    A.ntiles = p.y_tpb * cf.ne;
    const int grid = pick_grid(A.ntiles, p.y_tpb, p.nsm);
    if (A.nseg > 1 || grid % p.y_tpb != 0) A.prefetch = 0;   // see above: the class lookup of a new strip would sit in the tile's path
    const YMaps &maps = *reinterpret_cast<const YMaps *>(p.ymaps.data());
    if (fd.S == 8) return dispatch_y<8, 512>(c, A, maps, grid, p.y_cw, p.y_ns);
    return dispatch_y<16, 256>(c, A, maps, grid, p.y_cw, p.y_ns);
}
