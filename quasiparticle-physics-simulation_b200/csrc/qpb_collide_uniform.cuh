// Collision kernel for a FROZEN, CELL-INDEPENDENT phonon state (included by qpb_collision.cu).
//
// When freeze_phonon_dynamics is set and every cell holds the same phonon occupations (the reference's validation
// configuration, qpsim/validation.py:23-33, and BASELINE config 4a), the effective kernels of solver.py:726-743
//     Ke = Ks o Np,   Ka = Kr o nS,   Kb = Kr o (1 + nS)
// no longer depend on the cell, so the update is the four matrix products of SURVEY.md section 8 (box C, "GEMM
// form"):   loss = dE Ke p + 2dE Kb n,   gain = p o (dE Ke^T n + 2dE Ka p),   followed by the relaxation update.
// B200's FP64 tensor rate equals its FP64 vector rate, so the products run as a fused register-tiled GEMV: lanes of
// a warp are 32 cells, a thread owns 8 rows, the 8 x 4 tiles of the packed matrix (Ke_ij, Ke_ji, 2Kb_ij, 2Ka_ij
// premultiplied by dE) stream from L2 through the same cp.async ring as the structured kernel; 4 FMA per matrix
// element and cell, half the structured kernel's row pass, and no phonon passes at all.
#pragma once

struct UniformArgs {
    int ne, nep, ncell, ncd;
    double *S;
    const int32_t *c2d;
    const double4 *K4;   // [nep][nep]  (dE Ke_ij, dE Ke_ji, 2dE Kb_ij, 2dE Ka_ij)
    const double *rho;   // [nep] zero padded
    double dt;
    int mode;            // epilogue, see GemmArgs::mode
    const double *gth;   // [nep] thermal generation (mode 2)
};

template <int CC, int NT>
__global__ void __launch_bounds__(NT, 1) k_collide_uniform(UniformArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int nep = A.nep;
    constexpr int SUBS = 32 / CC;
    constexpr int NWARP = NT / 32;
    constexpr int UST = TI * TJ * 32;                  // bytes of one tile
    double *sn = sm;                                   // [nep][CC]
    double *sp = sn + (size_t)nep * CC;                // [nep][CC]
    char *ring_all = reinterpret_cast<char *>(sp + (size_t)nep * CC);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cl = lane % CC, sub = lane / CC;
    constexpr int nslot = NWARP * SUBS;
    const int slot = warp * SUBS + sub;
    char *ring = ring_all + (size_t)slot * NSTAGE * UST;
    const int cell0 = blockIdx.x * CC;
    const int ncell = A.ncell;
    static_assert(NT % CC == 0, "threads per CTA must be a multiple of the cells per CTA");
    {
        constexpr int RPT = NT / CC;
        constexpr int UN = 8;
        const int c_me = tid % CC, row_me = tid / CC;
        const int q_me = cell0 + c_me;
        const bool live_me = q_me < ncell;
        const long long d_me = live_me ? A.c2d[q_me] : 0;
        for (int i0 = row_me; i0 < nep; i0 += RPT * UN) {
            double nv[UN], rv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * RPT;
                const bool ok = live_me && i < A.ne;
                nv[u] = ok ? A.S[(long long)i * A.ncd + d_me] : 0.0;
                rv[u] = ok ? A.rho[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * RPT;
                if (i < nep) {
                    sn[i * CC + c_me] = nv[u];
                    sp[i * CC + c_me] = rv[u] * fmax(1.0 - nv[u] / fmax(rv[u], 1e-30), 0.0);
                }
            }
        }
    }
    __syncthreads();
    const double *cn = sn + cl, *cp = sp + cl;
    const int q = cell0 + cl;
    const bool live = q < ncell;
    const int nib = nep / TI, ntile = nep / TJ;
    const int nround = (nib + nslot - 1) / nslot;
    for (int rd = 0; rd < nround; ++rd) {
        const int ib = slot + rd * nslot;
        const bool work = ib < nib;
        const int i0 = work ? ib * TI : 0;
        double L[TI], G[TI];
#pragma unroll
        for (int r = 0; r < TI; ++r) L[r] = G[r] = 0.0;
        const char *gk = reinterpret_cast<const char *>(A.K4 + (size_t)i0 * nep);
        const size_t rstride = (size_t)nep * 32;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NSTAGE - 1; ++t) {
            if (t < ntile) ring_prefetch<CC, 2 * TJ>(ring + t * UST, gk + (size_t)t * TJ * 32, rstride, cl);
            cp_async_commit();
        }
        for (int t = 0; t < ntile; ++t) {
            __syncwarp();
            const int tn = t + NSTAGE - 1;
            if (tn < ntile) ring_prefetch<CC, 2 * TJ>(ring + (tn % NSTAGE) * UST, gk + (size_t)tn * TJ * 32, rstride, cl);
            cp_async_commit();
            cp_async_wait<NSTAGE - 1>();
            __syncwarp();
            const double2 *kt = reinterpret_cast<const double2 *>(ring + (t % NSTAGE) * UST);
            const int j0 = t * TJ;
            double nj[TJ], pj[TJ];
#pragma unroll
            for (int s = 0; s < TJ; ++s) {
                nj[s] = cn[(j0 + s) * CC];
                pj[s] = cp[(j0 + s) * CC];
            }
#pragma unroll
            for (int r = 0; r < TI; ++r) {
#pragma unroll
                for (int s = 0; s < TJ; ++s) {
                    const double2 ke = kt[(r * TJ + s) * 2], kr = kt[(r * TJ + s) * 2 + 1];
                    L[r] = fma(ke.x, pj[s], L[r]);     // dE Ke_ij p_j
                    L[r] = fma(kr.x, nj[s], L[r]);     // 2dE Kb_ij n_j
                    G[r] = fma(ke.y, nj[s], G[r]);     // dE Ke_ji n_j
                    G[r] = fma(kr.y, pj[s], G[r]);     // 2dE Ka_ij p_j
                }
            }
        }
        cp_async_wait<0>();
        if (live && work) {
            const int d = A.c2d[q];
#pragma unroll
            for (int r = 0; r < TI; ++r) {
                const int i = i0 + r;
                if (i < A.ne)
                    A.S[(long long)i * A.ncd + d] = collision_epilogue(A.mode, cn[i * CC], cp[i * CC], G[r], L[r],
                                                                       A.mode == 2 ? A.gth[i] : 0.0, A.dt);
            }
        }
    }
}
