// Structured collision kernel (included by qpb_collision.cu inside its anonymous namespace).
//
// Uniform energy grid: the phonon index maps depend on |i-j| and i+j only, the kernels are symmetric, one gap
// table.  Work decomposition:
//   * lanes of a warp are different CELLS (CC cells per CTA; 32/CC "sub-slots" per warp when CC < 32), so every
//     per-cell vector n, p, n_ph(|i-j|), n_ph(i+j) is a shared-memory column read without bank conflicts and
//     every kernel-matrix element is the same for all lanes of a sub-slot;
//   * a thread owns TI = 8 rows (pass 1), diagonals (pass 2) or anti-diagonals (pass 3) and walks the other index
//     in register tiles of TJ = 4; its accumulators never leave registers, so there is no cross-thread reduction;
//   * the kernel-matrix tiles (8 x 4 elements) stream from L2 through a per-warp cp.async ring (3 stages), so the
//     FP64 pipe does not wait on global loads; operands of a tile are read as 128-bit broadcasts.  The tables are
//     stored tile by tile, a ring stage is one flat copy (RingFeed);
//   * the columns pair what is read together: (n, p) of an energy index, and the phonon occupations of two
//     consecutive indices of a family, each one 128-bit shared-memory read - 61 % of the kernel's warp instructions
//     are DFMA / DMUL / DADD (DESIGN.md section 4.2 has the step-by-step record).
//
//   pass 1 (rows, quasiparticles)      L_i = sum_j [dE Ks (nd + [i>j]) p_j + 2dE Kr (1 + ns) n_j]
//                                      G_i = sum_j [dE Ks (nd + [j>i]) n_j + 2dE Kr ns p_j],  gain_i = p_i G_i
//   pass 2 (diagonals k = i-j > 0)     A_k = sum_j n_{j+k} dE Ks[j+k,j] p_j,  C_k = sum_j n_j dE Ks[j,j+k] p_{j+k}
//   pass 3 (anti-diagonals m = i+j)    R_m = sum n_i dE Kr n_j,  B_m = sum p_i dE Kr p_j   (symmetric halves folded)
#pragma once

constexpr int TI = 8;     // rows / diagonals / anti-diagonals per thread
constexpr int TJ = 4;     // columns per register tile
constexpr int PADF = 8;   // zero padding in front of the n,p columns in shared memory
constexpr int PADB = 16;  // and behind
constexpr int NSTAGE = 3; // cp.async ring depth
constexpr int NEPMAX = 512; // largest padded energy grid of the structured kernel (index tables ride in the parameters)
constexpr int QPB_MAX_RANKS = 8;

struct StructArgs {
    int ne, nep, nw, ncell, ncd;
    double *S;
    double *P;
    const int32_t *c2d;
    const double2 *K2;   // [nep][nep]  (dE*Ks, 2dE*Kr)                                  all three stored tile by tile
    const double *KsD;   // [nep][nep]  dE*Ks[j+k][j]
    const double *KrA;   // [2nep][nep] dE*Kr[m-j][j] * (2 if j<m-j, 1 if j==m-j, else 0)
    const double *rho;   // [nep] zero padded
    double dt;
    // Index tables as kernel parameters (constant bank: a lookup is not a memory round trip).  -1 = none.
    //   dmap[k]  phonon bin of the diagonal k = |i-j|          mofk[k]  >= 0 when that bin is also fed by an anti-diagonal
    //   smap[m]  phonon bin of the anti-diagonal m = i+j       kofm[m]  diagonal index k sharing that bin, or -1
    int16_t dmap[NEPMAX], mofk[NEPMAX], smap[2 * NEPMAX], kofm[2 * NEPMAX];
    // Layout exchange fused into the kernel (multi-GPU).  The quasiparticle state of a run lives in two
    // layouts: cell sharded here, bin sharded (dense grid) in the diffusion contexts of all ranks, whose state arrays are
    // mapped into this process (peer memory over NVLink).  xmode 1: the updated n(E) of my cells is stored straight
    // into the owning rank's diffusion state instead of my own; xmode 2: the staging loads read n(E) from there.
    //   route[i] = owner rank << 10 | row of bin i in the owner's state;  xdense[q] = dense grid index of my cell q
    // Several gap tables (non-uniform gap): the cells are regrouped so that a CTA's CC cells share one table.
    //   cperm[group * CC + lane] = cell index or -1 (padding);  ggid[group] = table of the group;  tables of gap g
    //   start at K2 + g*nep^2, KsD + g*nep^2, KrA + g*2*nep^2, rho + g*nep.  cperm == nullptr: cells in order, one table.
    const int32_t *cperm;
    const int32_t *ggid;
    int xmode;
    long long xncd;
    double *xpeer[QPB_MAX_RANKS];
    const int32_t *xdense;
    int16_t route[NEPMAX];
};

__device__ __forceinline__ double *exchange_slot(const StructArgs &A, int i, long long dense) {
    const int r = A.route[i];
    return A.xpeer[r >> 10] + (long long)(r & 1023) * A.xncd + dense;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Copy one TI-row tile (UPR 16-byte units per row) from global to this sub-slot's ring stage (k_collide_uniform).
template <int CC, int UPR>
__device__ __forceinline__ void ring_prefetch(char *stage, const char *gsrc, size_t row_stride, int cl) {
    constexpr int U = TI * UPR;
#pragma unroll
    for (int e0 = 0; e0 < U; e0 += CC) {
        const int e = e0 + cl;
        if (U % CC == 0 || e < U) cp_async16(stage + e * 16, gsrc + (size_t)(e / UPR) * row_stride + (e % UPR) * 16);
    }
}

__device__ __forceinline__ void cp_async16_s(uint32_t sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}

// Feeds a sub-slot's ring with the TI x TJ tiles of a table that follow each other along a row block.  The tables are
// stored tile by tile (struct_tile_index: a tile is one contiguous run of UNITS 16-byte units, the tiles of a row block
// follow each other), so a tile is a flat copy: whole 128-byte lines from L2 and few shared-memory write wavefronts
// (row-major tables: every 32- or 64-byte row piece of a tile arrived as a write of its own, 8-16 wavefronts per copy
// instruction where 4 are needed).  Everything that does not change from tile to tile is formed once per walk.
template <int CC, int UNITS>
struct RingFeed {
    static constexpr int N = (UNITS + CC - 1) / CC;    // units per lane
    static constexpr int STAGE = TI * TJ * 16;         // bytes per stage (the largest tile)
    const char *src;      // this lane's first unit of the next tile to load
    uint32_t dst0;        // shared address of the lane's first unit in stage 0
    uint32_t load_stage;  // stage the next load goes to
    bool mine;            // lane takes part (tiles smaller than the sub-slot)
    __device__ __forceinline__ void start(char *ring, const char *tile0, int cl) {
        src = tile0 + cl * 16;
        dst0 = (uint32_t)__cvta_generic_to_shared(ring) + cl * 16;
        load_stage = 0;
        mine = UNITS % CC == 0 || cl < UNITS;
    }
    // load the next tile (when there is one) and step to the tile after it
    __device__ __forceinline__ void load(bool there) {
        if (there && mine) {
            const uint32_t d = dst0 + load_stage * STAGE;
#pragma unroll
            for (int n = 0; n < N; ++n) cp_async16_s(d + n * CC * 16, src + n * CC * 16);
        }
        src += UNITS * 16;
        load_stage = load_stage + 1 == NSTAGE ? 0 : load_stage + 1;
    }
};

// Element (row, col) of a table stored tile by tile; ncols = columns of the table (a multiple of TJ), rows in blocks of TI.
__host__ __device__ __forceinline__ size_t struct_tile_index(int row, int col, int ncols) {
    return ((size_t)(row / TI) * (ncols / TJ) + col / TJ) * (TI * TJ) + (row % TI) * TJ + col % TJ;
}

// quasiparticle tile, one side of the diagonal (SIDE 0: i > j everywhere, 1: i < j everywhere, 2: mixed)
// cnd2 / cns2: the phonon occupations at |i-j| and i+j, two consecutive indices per 128-bit slot (ph_pair below)
template <int CC, bool SC, bool RC, int SIDE>
__device__ __forceinline__ void qp_tile(const double2 *__restrict__ kt, const double2 *__restrict__ cnp,
                                        const double2 *__restrict__ cnd2,
                                        const double2 *__restrict__ cns2, int i0, int j0, double (&L)[TI],
                                        double (&G)[TI]) {
    double nj[TJ], pj[TJ];
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
        const double2 v = cnp[(j0 + s) * CC];   // (n_j, p_j): one 128-bit read
        nj[s] = v.x;
        pj[s] = v.y;
    }
    // windows of TI + TJ - 1 = 11 consecutive indices as six 128-bit reads: i0 + j0 is even, the first index of the
    // |i-j| window is odd (kb is a multiple of TJ), so that one starts a slot earlier (element 0 unused)
    static_assert((TI + TJ - 1) <= 11 && TI % 2 == 0 && TJ % 2 == 0, "window slots");
    double nsw[12], ndw[12];
    if (RC) {
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const double2 v = cns2[(((i0 + j0) >> 1) + t) * CC];
            nsw[2 * t] = v.x;
            nsw[2 * t + 1] = v.y;
        }
    }
    const int kb = i0 - j0;
    if (SC && SIDE != 2) {
        const int base = SIDE == 0 ? kb - (TJ - 1) : -kb - (TI - 1);   // odd
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const double2 v = cnd2[(((base - 1) >> 1) + t) * CC];
            ndw[2 * t] = v.x;
            ndw[2 * t + 1] = v.y;
        }
    }
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            const double2 kv = kt[r * TJ + s];
            if (SC) {
                if (SIDE == 0) {
                    const double e = kv.x * ndw[r - s + TJ];
                    L[r] = fma(kv.x, pj[s], L[r]);       // spontaneous emission out of i
                    L[r] = fma(e, pj[s], L[r]);          // stimulated
                    G[r] = fma(e, nj[s], G[r]);
                } else if (SIDE == 1) {
                    const double e = kv.x * ndw[s - r + TI];
                    L[r] = fma(e, pj[s], L[r]);
                    G[r] = fma(e, nj[s], G[r]);
                    G[r] = fma(kv.x, nj[s], G[r]);       // spontaneous emission into i
                } else {
                    const int k = kb + r - s;
                    if (k != 0) {
                        const int ka = k > 0 ? k : -k;
                        const double2 v = cnd2[(ka >> 1) * CC];
                        const double e = kv.x * ((ka & 1) ? v.y : v.x);
                        L[r] = fma(k > 0 ? e + kv.x : e, pj[s], L[r]);
                        G[r] = fma(k > 0 ? e : e + kv.x, nj[s], G[r]);
                    }
                }
            }
            if (RC) {
                const double g = kv.y * nsw[r + s];
                L[r] = fma(kv.y, nj[s], L[r]);
                L[r] = fma(g, nj[s], L[r]);
                G[r] = fma(g, pj[s], G[r]);
            }
        }
    }
}

// quasiparticle tile that contains the diagonal: i0 is a multiple of TI and j0 of TJ, so kb = i0 - j0 is 0 or -TJ and
// every element's side of the diagonal is known at compile time (the diagonal itself carries Ks = 0)
template <int CC, bool SC, bool RC, int KB>
__device__ __forceinline__ void qp_tile_diag(const double2 *__restrict__ kt, const double2 *__restrict__ cnp,
                                             const double2 *__restrict__ cnd2,
                                             const double2 *__restrict__ cns2, int i0, int j0, double (&L)[TI],
                                             double (&G)[TI]) {
    double nj[TJ], pj[TJ];
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
        const double2 v = cnp[(j0 + s) * CC];   // (n_j, p_j): one 128-bit read
        nj[s] = v.x;
        pj[s] = v.y;
    }
    double nsw[12], nda[TI];   // nda[m] = n_ph at |i-j| = m, m < TI covers both values of KB
    if (RC) {
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const double2 v = cns2[(((i0 + j0) >> 1) + t) * CC];
            nsw[2 * t] = v.x;
            nsw[2 * t + 1] = v.y;
        }
    }
    if (SC) {
#pragma unroll
        for (int m = 0; m < TI; m += 2) {
            const double2 v = cnd2[(m >> 1) * CC];
            nda[m] = v.x;
            nda[m + 1] = v.y;
        }
    }
#pragma unroll
    for (int s = 0; s < TJ; ++s) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
            const double2 kv = kt[r * TJ + s];
            const int k = KB + r - s;   // compile-time after unrolling
            if (SC && k != 0) {
                const double e = kv.x * nda[k > 0 ? k : -k];
                if (k > 0) {
                    L[r] = fma(kv.x, pj[s], L[r]);
                    L[r] = fma(e, pj[s], L[r]);
                    G[r] = fma(e, nj[s], G[r]);
                } else {
                    L[r] = fma(e, pj[s], L[r]);
                    G[r] = fma(e, nj[s], G[r]);
                    G[r] = fma(kv.x, nj[s], G[r]);
                }
            }
            if (RC) {
                const double g = kv.y * nsw[r + s];
                L[r] = fma(kv.y, nj[s], L[r]);
                L[r] = fma(g, nj[s], L[r]);
                G[r] = fma(g, pj[s], G[r]);
            }
        }
    }
}

// CC = cells per CTA; NT = threads per CTA.
template <int CC, int NT, bool SC, bool RC, bool PH>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) k_collide_struct(const __grid_constant__ StructArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int nep = A.nep;
    const int ncol = nep + PADF + PADB;
    constexpr int SUBS = 32 / CC;
    constexpr int NWARP = NT / 32;
    constexpr int STAGE_BYTES = TI * TJ * 16;               // one K2 tile (the largest)
    double2 *snp = reinterpret_cast<double2 *>(sm);    // [ncol][CC]  (n, p) interleaved: a window element is one 128-bit read
    double *snd = sm + (size_t)2 * ncol * CC;          // [nep][CC]   n_ph at |i-j| ; later: stash a (diag family)
    double *sns = snd + (size_t)nep * CC;              // [2nep][CC]  n_ph at i+j   ; later: stash b (diag family)
    char *ring_all = reinterpret_cast<char *>(sns + (size_t)2 * nep * CC);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cl = lane % CC, sub = lane / CC;
    constexpr int nslot = NWARP * SUBS;
    const int slot = warp * SUBS + sub;
    char *ring = ring_all + (size_t)(warp * SUBS + sub) * NSTAGE * STAGE_BYTES;
    const int cell0 = blockIdx.x * CC;
    const int ncell = A.ncell;
    const int gid = A.cperm ? A.ggid[blockIdx.x] : 0;
    const double2 *tK2 = A.K2 + (size_t)gid * nep * nep;
    const double *tKsD = A.KsD + (size_t)gid * nep * nep;
    const double *tKrA = A.KrA + (size_t)gid * 2 * nep * nep;
    const double *trho = A.rho + (size_t)gid * nep;
    // cell of a lane / staging thread (regrouped by gap table when there are several)
    auto cell_of = [&](int lane_cell) { return A.cperm ? A.cperm[cell0 + lane_cell] : cell0 + lane_cell; };

    // ---- stage the per-cell columns -------------------------------------------------------------------
    // NT is a multiple of CC, so a thread always serves the same cell: its dense index is read once and every
    // column load below is independent of the others (8 in flight per thread).
    static_assert(NT % CC == 0, "threads per CTA must be a multiple of the cells per CTA");
    {
        // Every element goes global -> shared with an 8-byte cp.async (no register staging): all loads of the thread
        // are in flight together, one memory round trip for the whole staging instead of one per batch.
        constexpr int RPT = NT / CC;   // columns covered by one pass of the CTA
        const int c_me = tid % CC, row_me = tid / CC;
        const int q_me = cell_of(c_me);
        const bool live_me = q_me >= 0 && q_me < ncell;
        const long long d_me = live_me ? (A.xmode == 2 ? A.xdense[q_me] : A.c2d[q_me]) : 0;
        for (int col = row_me; col < ncol; col += RPT) {
            const int i = col - PADF;
            if (live_me && i >= 0 && i < A.ne)
                cp_async8(&snp[col * CC + c_me].x,
                          A.xmode == 2 ? exchange_slot(A, i, d_me) : &A.S[(long long)i * A.ncd + d_me]);
            else snp[col * CC + c_me].x = 0.0;
        }
        // phonon occupations of the two index families (snd and sns are contiguous: 3*nep rows)
        for (int idx = row_me; idx < 3 * nep; idx += RPT) {
            int om = -1;
            if (live_me) {
                if (idx < nep) {
                    if (idx < A.ne) om = A.dmap[idx];
                } else if (idx - nep < 2 * A.ne - 1) {
                    om = A.smap[idx - nep];
                }
            }
            // two consecutive indices of a family share a 128-bit slot: [idx / 2][cell][idx % 2] (nep is even, so the
            // formula runs through both families)
            double *slot_ph = &snd[((size_t)(idx >> 1) * CC + c_me) * 2 + (idx & 1)];
            if (om >= 0) cp_async8(slot_ph, &A.P[(long long)om * ncell + q_me]);
            else *slot_ph = 0.0;
        }
        cp_async_commit();
        cp_async_wait<0>();
        // p = rho * max(1 - n / max(rho, 1e-30), 0) of the thread's own elements (solver.py:719-721, 738)
        for (int col = row_me; col < ncol; col += RPT) {
            const int i = col - PADF;
            const double rv = (i >= 0 && i < A.ne) ? trho[i] : 0.0;
            const double nv = snp[col * CC + c_me].x;
            snp[col * CC + c_me].y = rv * fmax(1.0 - nv / fmax(rv, 1e-30), 0.0);
        }
    }
    __syncthreads();
    const double2 *cnp = snp + (size_t)PADF * CC + cl;   // cnp[idx*CC] = (n[idx], p[idx]) of this lane's cell
    const double2 *cnd2 = reinterpret_cast<const double2 *>(snd) + cl;   // cnd2[(k / 2) * CC] = n_ph at |i-j| = k, k + 1 (k even)
    const double2 *cns2 = reinterpret_cast<const double2 *>(sns) + cl;   // cns2[(m / 2) * CC] = n_ph at i+j = m, m + 1 (m even)
    const int q = cell_of(cl);
    const bool live = q >= 0 && q < ncell;

    // ---- pass 1: rows ------------------------------------------------------------------------------------
    {
        const int nib = nep / TI;
        const int ntile = nep / TJ;
        // all sub-slots of a warp iterate the same number of times so that __syncwarp() is legal
        const int nround = (nib + nslot - 1) / nslot;
        for (int rd = 0; rd < nround; ++rd) {
            const int ib = slot + rd * nslot;
            const bool work = ib < nib;
            const int i0 = work ? ib * TI : 0;
            double L[TI], G[TI];
#pragma unroll
            for (int r = 0; r < TI; ++r) L[r] = G[r] = 0.0;
            const char *gk = reinterpret_cast<const char *>(tK2 + struct_tile_index(i0, 0, nep));
            __syncwarp();  // the previous round's last tile is no longer being read
            RingFeed<CC, TI * TJ> feed;
            feed.start(ring, gk, cl);
#pragma unroll
            for (int t = 0; t < NSTAGE - 1; ++t) {
                feed.load(t < ntile);
                cp_async_commit();
            }
            const char *kt_stage = ring;
            for (int t = 0; t < ntile; ++t) {
                // the lane's own pieces of tile t have landed; after the warp barrier everybody's have, and everybody is
                // done with tile t - 1, whose stage the next load overwrites
                cp_async_wait<NSTAGE - 2>();
                __syncwarp();
                feed.load(t + NSTAGE - 1 < ntile);
                cp_async_commit();
                const double2 *kt = reinterpret_cast<const double2 *>(kt_stage);
                kt_stage = kt_stage + STAGE_BYTES == ring + NSTAGE * STAGE_BYTES ? ring : kt_stage + STAGE_BYTES;
                const int j0 = t * TJ;
                const int kb = i0 - j0;
                if (kb >= TJ) qp_tile<CC, SC, RC, 0>(kt, cnp, cnd2, cns2, i0, j0, L, G);
                else if (kb <= -TI) qp_tile<CC, SC, RC, 1>(kt, cnp, cnd2, cns2, i0, j0, L, G);
                else if (kb == 0) qp_tile_diag<CC, SC, RC, 0>(kt, cnp, cnd2, cns2, i0, j0, L, G);
                else if (kb == -TJ) qp_tile_diag<CC, SC, RC, -TJ>(kt, cnp, cnd2, cns2, i0, j0, L, G);
                else qp_tile<CC, SC, RC, 2>(kt, cnp, cnd2, cns2, i0, j0, L, G);
            }
            cp_async_wait<0>();
            if (live && work) {
                const long long d = A.xmode == 1 ? A.xdense[q] : A.c2d[q];
#pragma unroll
                for (int r = 0; r < TI; ++r) {
                    const int i = i0 + r;
                    if (i < A.ne) {
                        double *dst = A.xmode == 1 ? exchange_slot(A, i, d) : &A.S[(long long)i * A.ncd + d];
                        const double2 v = cnp[i * CC];
                        *dst = relax_update(v.x, v.y * G[r], L[r], A.dt);
                    }
                }
            }
        }
    }
    if (!PH) return;
    __syncthreads();  // everyone is done with n_ph(|i-j|), n_ph(i+j): the region becomes the diagonal-family stash
    double *sta = snd;                         // a of the diagonal family  [nep][CC]
    double *stb = snd + (size_t)nep * CC;      // b of the diagonal family  [nep][CC]

    // ---- pass 2: diagonals k = i-j > 0 (scattering phonon source) ----------------------------------------
    // Blocks of 8 diagonals, a long one paired with a short one; when there are fewer pairs than warp slots every
    // pair is cut along j into `parts` pieces whose partial sums meet in shared memory.
    if (SC) {
        const int nkb = nep / TI;
        const int npair = (nkb + 1) / 2;
        int parts = 1;
        while (parts * 2 * npair <= nslot && parts < 8) parts *= 2;
        if (parts > 1) {
            for (int e = tid; e < 2 * nep * CC; e += NT) sta[e] = 0.0;
            __syncthreads();
        }
        const int nunit = npair * parts;
        const int nround = (nunit + nslot - 1) / nslot;
        for (int rd = 0; rd < nround; ++rd) {
            const int unit = slot + rd * nslot;
            const int it = unit / parts, part = unit - it * parts;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int kbk = half == 0 ? it : nkb - 1 - it;   // pair a long block with a short one
                const bool work = unit < nunit && !(half == 1 && kbk == it);
                const int k0 = work ? kbk * TI : 0;
                double Aem[TI], Cab[TI];
#pragma unroll
                for (int r = 0; r < TI; ++r) Aem[r] = Cab[r] = 0.0;
                // this piece's tile range; the longest range among the lanes of the warp decides the trip count
                const int nt_all = work ? (nep - k0) / TJ : 0;
                const int t_lo = (nt_all * part) / parts, t_hi = (nt_all * (part + 1)) / parts;
                const int mytiles = t_hi - t_lo;
                int ntile = mytiles;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ntile = max(ntile, __shfl_xor_sync(0xffffffffu, ntile, o));
                const char *gk = reinterpret_cast<const char *>(tKsD + struct_tile_index(k0, t_lo * TJ, nep));
                __syncwarp();
                RingFeed<CC, TI * TJ / 2> feed;
                feed.start(ring, gk, cl);
#pragma unroll
                for (int t = 0; t < NSTAGE - 1; ++t) {
                    feed.load(t < mytiles);
                    cp_async_commit();
                }
                const char *kt_stage = ring;
                for (int t = 0; t < ntile; ++t) {
                    cp_async_wait<NSTAGE - 2>();
                    __syncwarp();
                    feed.load(t + NSTAGE - 1 < mytiles);
                    cp_async_commit();
                    if (t >= mytiles) continue;
                    const double *kt = reinterpret_cast<const double *>(kt_stage);
                    kt_stage = kt_stage + STAGE_BYTES == ring + NSTAGE * STAGE_BYTES ? ring : kt_stage + STAGE_BYTES;
                    const int j0 = (t_lo + t) * TJ;
                    double nj[TJ], pj[TJ], nwn[TI + TJ - 1], pwn[TI + TJ - 1];
#pragma unroll
                    for (int s = 0; s < TJ; ++s) {
                        const double2 v = cnp[(j0 + s) * CC];
                        nj[s] = v.x;
                        pj[s] = v.y;
                    }
#pragma unroll
                    for (int t2 = 0; t2 < TI + TJ - 1; ++t2) {
                        const double2 v = cnp[(j0 + k0 + t2) * CC];
                        nwn[t2] = v.x;
                        pwn[t2] = v.y;
                    }
#pragma unroll
                    for (int r = 0; r < TI; ++r) {
#pragma unroll
                        for (int s = 0; s < TJ; ++s) {
                            const double kv = kt[r * TJ + s];
                            Aem[r] = fma(nwn[r + s], kv * pj[s], Aem[r]);   // n_{j+k} Ks p_j   (emission, i = j+k)
                            Cab[r] = fma(pwn[r + s], kv * nj[s], Cab[r]);   // n_j Ks p_{j+k}   (absorption, i = j)
                        }
                    }
                }
                cp_async_wait<0>();
                if (live && work) {
#pragma unroll
                    for (int r = 0; r < TI; ++r) {
                        const int k = k0 + r;
                        if (k >= A.ne) continue;
                        const double a = Aem[r], b = Aem[r] - Cab[r];
                        if (parts > 1) {
                            atomicAdd(&sta[k * CC + cl], a);
                            atomicAdd(&stb[k * CC + cl], b);
                        } else {
                            const int om = A.dmap[k];
                            if (RC && A.mofk[k] >= 0) {
                                sta[k * CC + cl] = a;
                                stb[k * CC + cl] = b;
                            } else {
                                const long long o = (long long)om * ncell + q;
                                A.P[o] = affine_growth(A.P[o], a, b, A.dt);
                            }
                        }
                    }
                }
            }
        }
        if (parts > 1) {
            // the pieces have met: every thread finishes its share of (diagonal, cell) pairs
            __syncthreads();
            constexpr int RPT = NT / CC;
            const int c_me = tid % CC, q_me = cell_of(c_me);
            if (q_me >= 0 && q_me < ncell) {
                for (int k = tid / CC; k < A.ne; k += RPT) {
                    const int om = A.dmap[k];
                    if (RC && A.mofk[k] >= 0) continue;   // also fed by an anti-diagonal: stays in the stash for pass 3
                    const long long o = (long long)om * ncell + q_me;
                    A.P[o] = affine_growth(A.P[o], sta[k * CC + c_me], stb[k * CC + c_me], A.dt);
                }
            }
        }
    }
    if (!RC) return;
    __syncthreads();

    // ---- pass 3: anti-diagonals m = i+j (recombination / pair breaking phonon source) ------------------------
    {
        const int nmb = 2 * nep / TI;       // blocks of anti-diagonals (the last one is partly padding)
        const int hb = nmb / 2;
        const int nround = (hb + nslot - 1) / nslot;
        for (int rd = 0; rd < nround; ++rd) {
            const int it = slot + rd * nslot;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const bool work = it < hb;
                const int mb = work ? it + half * hb : 0;   // work(mb) + work(mb + hb) is constant
                const int m0 = mb * TI;
                double R[TI], Bp[TI];
#pragma unroll
                for (int r = 0; r < TI; ++r) R[r] = Bp[r] = 0.0;
                int jlo = m0 - (nep - 1);
                jlo = jlo < 0 ? 0 : (jlo / TJ) * TJ;
                int jhi = (m0 + TI - 1) / 2;          // largest j with j <= m-j for some m of the block
                if (jhi > nep - 1) jhi = nep - 1;
                const int mytiles = work ? (jhi - jlo) / TJ + 1 : 0;
                int ntile = mytiles;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ntile = max(ntile, __shfl_xor_sync(0xffffffffu, ntile, o));
                const char *gk = reinterpret_cast<const char *>(tKrA + struct_tile_index(m0, jlo, nep));
                __syncwarp();
                RingFeed<CC, TI * TJ / 2> feed;
                feed.start(ring, gk, cl);
#pragma unroll
                for (int t = 0; t < NSTAGE - 1; ++t) {
                    feed.load(t < mytiles);
                    cp_async_commit();
                }
                const char *kt_stage = ring;
                for (int t = 0; t < ntile; ++t) {
                    cp_async_wait<NSTAGE - 2>();
                    __syncwarp();
                    feed.load(t + NSTAGE - 1 < mytiles);
                    cp_async_commit();
                    if (t >= mytiles) continue;
                    const double *kt = reinterpret_cast<const double *>(kt_stage);
                    kt_stage = kt_stage + STAGE_BYTES == ring + NSTAGE * STAGE_BYTES ? ring : kt_stage + STAGE_BYTES;
                    const int j0 = jlo + t * TJ;
                    double nj[TJ], pj[TJ], nwn[TI + TJ - 1], pwn[TI + TJ - 1];
#pragma unroll
                    for (int s = 0; s < TJ; ++s) {
                        const double2 v = cnp[(j0 + s) * CC];
                        nj[s] = v.x;
                        pj[s] = v.y;
                    }
                    const int base = m0 - j0 - (TJ - 1);     // index m-j = base + (r - s + TJ-1)
#pragma unroll
                    for (int t2 = 0; t2 < TI + TJ - 1; ++t2) {
                        const double2 v = cnp[(base + t2) * CC];
                        nwn[t2] = v.x;
                        pwn[t2] = v.y;
                    }
#pragma unroll
                    for (int r = 0; r < TI; ++r) {
#pragma unroll
                        for (int s = 0; s < TJ; ++s) {
                            const double kv = kt[r * TJ + s];
                            R[r] = fma(nwn[r - s + TJ - 1], kv * nj[s], R[r]);
                            Bp[r] = fma(pwn[r - s + TJ - 1], kv * pj[s], Bp[r]);
                        }
                    }
                }
                cp_async_wait<0>();
                if (live && work) {
#pragma unroll
                    for (int r = 0; r < TI; ++r) {
                        const int m = m0 + r;
                        if (m >= 2 * A.ne - 1) continue;
                        const int om = A.smap[m];
                        double a = R[r], b = R[r];
                        const int k = SC ? A.kofm[m] : -1;
                        if (k >= 0) {   // same phonon bin also fed by the diagonal family
                            a = sta[k * CC + cl] + R[r];
                            b = stb[k * CC + cl] + R[r];
                        }
                        b -= Bp[r];
                        const long long o = (long long)om * ncell + q;
                        A.P[o] = affine_growth(A.P[o], a, b, A.dt);
                    }
                }
            }
        }
    }
}
