// Micro-benchmark of the collision kernel's rows pass (qpb_collide_struct.cuh, pass 1) on shared-memory-resident
// operands: what keeps its FP64 pipe from the rate a pure DFMA loop reaches?  Same CTA shape as the real kernel (512
// threads, 32 cells, one CTA per SM, columns [idx][32] in shared memory), the real qp_tile code for variant 0 and
// re-orderings of its instruction stream for the others.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {
__device__ __forceinline__ double relax_update(double n, double g, double l, double dt) { return n + g - l * dt; }
__device__ __forceinline__ double affine_growth(double y, double a, double b, double dt) { return y + a - b * dt; }
#include "../quasiparticle-physics-simulation_b200/csrc/qpb_collide_struct.cuh"

constexpr int CC = 32, NT = 512, NEP = 128;

// variant 1: column-major walk of the tile (s outer, r inner): consecutive DFMAs share n_j / p_j in one operand slot
template <int SIDE>
__device__ __forceinline__ void tile_v1(const double2 *__restrict__ kt, const double *__restrict__ cn,
                                        const double *__restrict__ cp, const double *__restrict__ cnd,
                                        const double *__restrict__ cns, int i0, int j0, double (&L)[TI], double (&G)[TI]) {
    double nsw[TI + TJ - 1], ndw[TI + TJ - 1];
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) nsw[t] = cns[(i0 + j0 + t) * CC];
    const int kb = i0 - j0;
    const int base = SIDE == 0 ? kb - (TJ - 1) : -kb - (TI - 1);
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) ndw[t] = cnd[(base + t) * CC];
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
        const double nj = cn[(j0 + s) * CC], pj = cp[(j0 + s) * CC];
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            const double2 kv = kt[r * TJ + s];
            const double e = kv.x * ndw[SIDE == 0 ? r - s + TJ - 1 : s - r + TI - 1];
            if (SIDE == 0) {
                L[r] = fma(e + kv.x, pj, L[r]);
                G[r] = fma(e, nj, G[r]);
            } else {
                L[r] = fma(e, pj, L[r]);
                G[r] = fma(e + kv.x, nj, G[r]);
            }
            const double g = kv.y * nsw[r + s];
            L[r] = fma(g + kv.y, nj, L[r]);
            G[r] = fma(g, pj, G[r]);
        }
    }
}

// variant 2: like the real tile, but the "+1" of the spontaneous term is an FMA of its own (no e -> e + kx chain)
template <int SIDE>
__device__ __forceinline__ void tile_v2(const double2 *__restrict__ kt, const double *__restrict__ cn,
                                        const double *__restrict__ cp, const double *__restrict__ cnd,
                                        const double *__restrict__ cns, int i0, int j0, double (&L)[TI], double (&G)[TI]) {
    double nj[TJ], pj[TJ];
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
        nj[s] = cn[(j0 + s) * CC];
        pj[s] = cp[(j0 + s) * CC];
    }
    double nsw[TI + TJ - 1], ndw[TI + TJ - 1];
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) nsw[t] = cns[(i0 + j0 + t) * CC];
    const int kb = i0 - j0;
    const int base = SIDE == 0 ? kb - (TJ - 1) : -kb - (TI - 1);
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) ndw[t] = cnd[(base + t) * CC];
#pragma unroll
    for (int r = 0; r < TI; ++r) {
#pragma unroll
        for (int s = 0; s < TJ; ++s) {
            const double2 kv = kt[r * TJ + s];
            const double nd = ndw[SIDE == 0 ? r - s + TJ - 1 : s - r + TI - 1];
            const double e = kv.x * nd, e1 = fma(kv.x, nd, kv.x);
            L[r] = fma(SIDE == 0 ? e1 : e, pj[s], L[r]);
            G[r] = fma(SIDE == 0 ? e : e1, nj[s], G[r]);
            const double ns = nsw[r + s];
            const double g = kv.y * ns, g1 = fma(kv.y, ns, kv.y);
            L[r] = fma(g1, nj[s], L[r]);
            G[r] = fma(g, pj[s], G[r]);
        }
    }
}

// variant 3: products first (all e, g of a row), then the accumulations: two clean instruction groups per row
template <int SIDE>
__device__ __forceinline__ void tile_v3(const double2 *__restrict__ kt, const double *__restrict__ cn,
                                        const double *__restrict__ cp, const double *__restrict__ cnd,
                                        const double *__restrict__ cns, int i0, int j0, double (&L)[TI], double (&G)[TI]) {
    double nj[TJ], pj[TJ];
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
        nj[s] = cn[(j0 + s) * CC];
        pj[s] = cp[(j0 + s) * CC];
    }
    double nsw[TI + TJ - 1], ndw[TI + TJ - 1];
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) nsw[t] = cns[(i0 + j0 + t) * CC];
    const int kb = i0 - j0;
    const int base = SIDE == 0 ? kb - (TJ - 1) : -kb - (TI - 1);
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) ndw[t] = cnd[(base + t) * CC];
#pragma unroll
    for (int r = 0; r < TI; ++r) {
        double e[TJ], e1[TJ], g[TJ], g1[TJ];
#pragma unroll
        for (int s = 0; s < TJ; ++s) {
            const double2 kv = kt[r * TJ + s];
            e[s] = kv.x * ndw[SIDE == 0 ? r - s + TJ - 1 : s - r + TI - 1];
            e1[s] = e[s] + kv.x;
            g[s] = kv.y * nsw[r + s];
            g1[s] = g[s] + kv.y;
        }
        double l = L[r], gg = G[r];
#pragma unroll
        for (int s = 0; s < TJ; ++s) {
            l = fma(SIDE == 0 ? e1[s] : e[s], pj[s], l);
            gg = fma(SIDE == 0 ? e[s] : e1[s], nj[s], gg);
            l = fma(g1[s], nj[s], l);
            gg = fma(g[s], pj[s], gg);
        }
        L[r] = l;
        G[r] = gg;
    }
}

// variant 4: half a column at a time (4 rows): the 16 products first, then the 16 accumulations grouped by the
// multiplicand they share (p_j, p_j, n_j, n_j), so that consecutive DFMAs can take it from the operand reuse cache
template <int SIDE>
__device__ __forceinline__ void tile_v4(const double2 *__restrict__ kt, const double *__restrict__ cn,
                                        const double *__restrict__ cp, const double *__restrict__ cnd,
                                        const double *__restrict__ cns, int i0, int j0, double (&L)[TI], double (&G)[TI]) {
    double nsw[TI + TJ - 1], ndw[TI + TJ - 1];
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) nsw[t] = cns[(i0 + j0 + t) * CC];
    const int kb = i0 - j0;
    const int base = SIDE == 0 ? kb - (TJ - 1) : -kb - (TI - 1);
#pragma unroll
    for (int t = 0; t < TI + TJ - 1; ++t) ndw[t] = cnd[(base + t) * CC];
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
        const double nj = cn[(j0 + s) * CC], pj = cp[(j0 + s) * CC];
#pragma unroll
        for (int h = 0; h < TI; h += 4) {
            double e[4], e1[4], g[4], g1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = h + q;
                const double2 kv = kt[r * TJ + s];
                e[q] = kv.x * ndw[SIDE == 0 ? r - s + TJ - 1 : s - r + TI - 1];
                e1[q] = e[q] + kv.x;
                g[q] = kv.y * nsw[r + s];
                g1[q] = g[q] + kv.y;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) L[h + q] = fma(SIDE == 0 ? e1[q] : e[q], pj, L[h + q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) G[h + q] = fma(g[q], pj, G[h + q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) G[h + q] = fma(SIDE == 0 ? e[q] : e1[q], nj, G[h + q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) L[h + q] = fma(g1[q], nj, L[h + q]);
        }
    }
}

// variant 5: 256 DFMAs per tile on register operands only (the pipe's own ceiling in this CTA shape)
__device__ __forceinline__ void tile_peak(double x, double y, double (&L)[TI], double (&G)[TI]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            L[r] = fma(L[r], x, y);
            G[r] = fma(G[r], y, x);
        }
    }
}

// variant 6: 256 FP64 instructions per tile in the real mix (64 DMUL, 64 DADD, 128 DFMA) on register operands
__device__ __forceinline__ void tile_mix(double x, double y, double (&L)[TI], double (&G)[TI]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            const double e = L[r] * x, e1 = e + y;
            L[r] = fma(e1, x, G[r]);
            G[r] = fma(e, y, L[r]);
        }
    }
}

// variants 7 / 8: 256 DFMAs per tile, accumulate form.  7: every DFMA reads three registers no neighbour shares
// (X[r], Y[r], acc[r]); 8: one multiplicand is the same register for all (operand reuse possible)
__device__ __forceinline__ void tile_distinct(const double (&X)[TI], const double (&Y)[TI], double (&L)[TI], double (&G)[TI]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            L[r] = fma(X[r], Y[(r + k) & 7], L[r]);
            G[r] = fma(Y[r], X[(r + k + 3) & 7], G[r]);
        }
    }
}
__device__ __forceinline__ void tile_shared(const double (&X)[TI], const double (&Y)[TI], double (&L)[TI], double (&G)[TI]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            L[r] = fma(X[k & 7], Y[r], L[r]);
            G[r] = fma(X[k & 7], Y[(r + 1) & 7], G[r]);
        }
    }
}

template <int VAR, bool ASYNC>
__global__ void __launch_bounds__(NT, 1) k_rows(const double2 *__restrict__ K2, double *__restrict__ out, int reps) {
    extern __shared__ __align__(16) double sm[];
    constexpr int nep = NEP, ncol = nep + PADF + PADB;
    constexpr int STAGE_BYTES = TI * TJ * 16;
    double *sn = sm, *sp = sn + ncol * CC, *snd = sp + ncol * CC, *sns = snd + nep * CC;
    char *ring_all = reinterpret_cast<char *>(sns + 2 * nep * CC);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < (2 * ncol + 3 * nep) * CC; e += NT) sm[e] = 1e-3 * ((e * 2654435761u) >> 20) / 4096.0;
    char *ring = ring_all + (size_t)warp * NSTAGE * STAGE_BYTES;
    for (int e = lane; e < NSTAGE * STAGE_BYTES / 8; e += 32) reinterpret_cast<double *>(ring)[e] = 1e-3 * (e + 1);
    __syncthreads();
    const double *cn = sn + PADF * CC + lane, *cp = sp + PADF * CC + lane, *cnd = snd + lane, *cns = sns + lane;
    const int i0 = warp * TI;
    const int ntile = nep / TJ;
    double L[TI], G[TI];
#pragma unroll
    for (int r = 0; r < TI; ++r) L[r] = G[r] = 1e-3 * r;
    double X[TI], Y[TI];
#pragma unroll
    for (int r = 0; r < TI; ++r) {
        X[r] = cn[r * CC] * 1e-3;
        Y[r] = cp[r * CC] * 1e-3;
    }
    const char *gk = reinterpret_cast<const char *>(K2 + (size_t)i0 * nep);
    const size_t rstride = (size_t)nep * 16;
    for (int rep = 0; rep < reps; ++rep) {
        if (ASYNC) {
            __syncwarp();
#pragma unroll
            for (int t = 0; t < NSTAGE - 1; ++t) {
                ring_prefetch<CC, TJ>(ring + t * STAGE_BYTES, gk + (size_t)t * TJ * 16, rstride, lane);
                cp_async_commit();
            }
        }
        for (int t = 0; t < ntile; ++t) {
            if (ASYNC) {
                __syncwarp();
                const int tn = t + NSTAGE - 1;
                if (tn < ntile) ring_prefetch<CC, TJ>(ring + (tn % NSTAGE) * STAGE_BYTES, gk + (size_t)tn * TJ * 16, rstride, lane);
                cp_async_commit();
                cp_async_wait<NSTAGE - 1>();
                __syncwarp();
            }
            const double2 *kt = reinterpret_cast<const double2 *>(ring + (t % NSTAGE) * STAGE_BYTES);
            const int j0 = t * TJ;
            const int kb = i0 - j0;
            if (VAR == 5) { tile_peak(1.0000001, 1e-9, L, G); continue; }
            if (VAR == 6) { tile_mix(1.0000001, 1e-9, L, G); continue; }
            if (VAR == 7) { tile_distinct(X, Y, L, G); continue; }
            if (VAR == 8) { tile_shared(X, Y, L, G); continue; }
            if (kb >= TJ) {
                if (VAR == 0) qp_tile<CC, true, true, 0>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 1) tile_v1<0>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 2) tile_v2<0>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 3) tile_v3<0>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 4) tile_v4<0>(kt, cn, cp, cnd, cns, i0, j0, L, G);
            } else if (kb <= -TI) {
                if (VAR == 0) qp_tile<CC, true, true, 1>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 1) tile_v1<1>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 2) tile_v2<1>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 3) tile_v3<1>(kt, cn, cp, cnd, cns, i0, j0, L, G);
                if (VAR == 4) tile_v4<1>(kt, cn, cp, cnd, cns, i0, j0, L, G);
            } else if (kb == 0) qp_tile_diag<CC, true, true, 0>(kt, cn, cp, cnd, cns, i0, j0, L, G);
            else if (kb == -TJ) qp_tile_diag<CC, true, true, -TJ>(kt, cn, cp, cnd, cns, i0, j0, L, G);
            else qp_tile<CC, true, true, 2>(kt, cn, cp, cnd, cns, i0, j0, L, G);
        }
        if (ASYNC) cp_async_wait<0>();
    }
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < TI; ++r) acc += L[r] + G[r];
    out[(size_t)blockIdx.x * NT + tid] = acc;
}

template <int VAR, bool ASYNC>
void run(const char *name, const double2 *K2, double *out, int nsm, int reps, double ghz) {
    const size_t smem = sizeof(double) * CC * (2 * (NEP + PADF + PADB) + 3 * NEP) + (size_t)(NT / 32) * NSTAGE * TI * TJ * 16;
    cudaFuncSetAttribute(k_rows<VAR, ASYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_rows<VAR, ASYNC><<<nsm, NT, smem>>>(K2, out, 2);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        k_rows<VAR, ASYNC><<<nsm, NT, smem>>>(K2, out, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    // per SM: 16 warps x 32 tiles x 256 FP64 warp instructions per repetition; the pipe takes 2 per clock and SM
    const double inst = 16.0 * 32 * 256 * reps;
    const double clk = best * 1e-3 * ghz * 1e9;
    printf("%-34s %8.3f ms  %6.1f clk per warp-tile  FP64 pipe %5.1f %%  err=%s\n", name, best, clk / (32.0 * reps) / 16.0 * 16.0 / 16.0,
           100.0 * inst / (2.0 * clk), cudaGetErrorString(cudaGetLastError()));
}
}  // namespace

int main() {
    int dev = 0, nsm = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double ghz = khz * 1e-6;
    double2 *K2;
    double *out;
    cudaMalloc(&K2, sizeof(double2) * NEP * NEP);
    cudaMalloc(&out, sizeof(double) * nsm * NT);
    std::vector<double2> h(NEP * NEP);
    for (int i = 0; i < NEP * NEP; ++i) h[i] = make_double2(1e-3 * (i % 97), 2e-3 * (i % 89));
    cudaMemcpy(K2, h.data(), sizeof(double2) * NEP * NEP, cudaMemcpyHostToDevice);
    const int reps = 400;
    printf("SMs %d, clock %.3f GHz (nominal); FP64 pipe %% assumes that clock\n", nsm, ghz);
    run<5, false>("pure DFMA, registers", K2, out, nsm, reps, ghz);
    run<6, false>("real mix MUL/ADD/FMA, registers", K2, out, nsm, reps, ghz);
    run<7, false>("DFMA acc += x[r]*y[r'] (distinct)", K2, out, nsm, reps, ghz);
    run<8, false>("DFMA acc += x*y[r] (shared operand)", K2, out, nsm, reps, ghz);
    run<0, false>("v0 real tile, K tile resident", K2, out, nsm, reps, ghz);
    run<0, true>("v0 real tile, cp.async ring", K2, out, nsm, reps, ghz);
    run<1, false>("v1 column-major walk", K2, out, nsm, reps, ghz);
    run<1, true>("v1 column-major walk, ring", K2, out, nsm, reps, ghz);
    run<2, false>("v2 fma for the +1", K2, out, nsm, reps, ghz);
    run<2, true>("v2 fma for the +1, ring", K2, out, nsm, reps, ghz);
    run<3, false>("v3 products first", K2, out, nsm, reps, ghz);
    run<3, true>("v3 products first, ring", K2, out, nsm, reps, ghz);
    run<4, false>("v4 grouped by shared operand", K2, out, nsm, reps, ghz);
    run<4, true>("v4 grouped by shared operand, ring", K2, out, nsm, reps, ghz);
    return 0;
}
