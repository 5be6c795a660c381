// FP64 tensor-core (DMMA) form of the collision step for FROZEN, CELL-INDEPENDENT phonons (included by
// qpb_collision.cu).  BASELINE config 4a: with every cell sharing one frozen phonon state the effective kernels of
// solver.py:726-743 are cell independent, and the update of all cells is (box C, "GEMM form")
//
//     [ L ]   [ dE Ke      2dE Kb   ]   [ P ]        L = loss,   gain = P o G
//     [ G ] = [ 2dE Ka     dE Ke^T  ] x [ S ]        S = n[NE][N],  P = rho o max(1 - S/rho, 0)
//
// one (2NE x 2NE) x (2NE x N) product followed by the element-wise relaxation update (solver.py:655-665).
//
//  * k_gemm_pack writes the right-hand operands once per call, compact and zero padded: Xn = S, Xp = P  [nep][npadc]
//    (one division per cell*bin instead of one per cell*bin and row block);
//  * k_collide_gemm: CTA tile = 64 energy rows (128 accumulator rows: L and G) x 128 cells, K chunks of 16 bins
//    through a 3-stage cp.async ring; 8 warps as 4 (row groups of 16) x 2 (cell halves of 64); a warp keeps
//    2 x 8 DMMA tiles of L and of G in registers (64 doubles per lane) and issues 64 mma.m8n8k4.f64 per 4 bins.
//    Shared-memory pitches (20 doubles for the matrices, 132 for the vectors) make every fragment load conflict
//    free.  The epilogue reads n, p of its own rows, applies the relaxation update and scatters into the dense state;
//    nothing reads the dense state during the product, so the update is in place.
#pragma once

struct GemmArgs {
    int ne, nep, ncell, npadc, ncd;
    double *S;
    const int32_t *c2d;
    const double *M;     // [4][nep][nep] row major: dE Ke_ij | 2dE Kb_ij | 2dE Ka_ij | dE Ke_ji
    double *Xn, *Xp;     // [nep][npadc]
    const double *rho;   // [nep] zero padded
    double dt;
    // epilogue: 0 relaxation update of the coupled solver (solver.py:655-665);
    //           1 forward-Euler scattering  n + dt (p G - n L)   (solver.py:551-581, M = [dE Ks | 0 | 0 | dE Ks^T]);
    //           2 forward-Euler recombination  n + dt (gth - n L)   (solver.py:584-605, M = [0 | 2dE Kr | 0 | 0])
    int mode;
    const double *gth;   // [nep] thermal generation (mode 2)
};

__device__ __forceinline__ double collision_epilogue(int mode, double n, double p, double G, double L, double gth,
                                                     double dt) {
    if (mode == 0) return relax_update(n, p * G, L, dt);
    if (mode == 1) return fmax(n + dt * (p * G - n * L), 0.0);
    return fmax(n + dt * (gth - n * L), 0.0);
}

constexpr int GM_BM = 64, GM_BN = 128, GM_BK = 16, GM_ST = 3;
constexpr int GM_AP = GM_BK + 4;     // matrix tile pitch (doubles)
constexpr int GM_BP = GM_BN + 4;     // vector tile pitch
constexpr int GM_A_BYTES = 4 * GM_BM * GM_AP * 8;
constexpr int GM_B_BYTES = 2 * GM_BK * GM_BP * 8;
constexpr int GM_STAGE_BYTES = GM_A_BYTES + GM_B_BYTES;

__global__ void k_gemm_pack(GemmArgs A) {
    const long long total = (long long)A.ne * A.ncell;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(g / A.ncell);
        const int q = (int)(g - (long long)i * A.ncell);
        const double n = A.S[(long long)i * A.ncd + A.c2d[q]];
        const double r = A.rho[i];
        A.Xn[(long long)i * A.npadc + q] = n;
        A.Xp[(long long)i * A.npadc + q] = r * fmax(1.0 - n / fmax(r, 1e-30), 0.0);   // solver.py:719-721, 738
    }
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) k_collide_gemm(GemmArgs A) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    const int i0 = blockIdx.y * GM_BM, c0 = blockIdx.x * GM_BN;
    const int nep = A.nep;
    const size_t npadc = (size_t)A.npadc;
    const int nchunk = nep / GM_BK;

    auto load_stage = [&](int kc, int s) {
        unsigned char *sa = gsm + (size_t)s * GM_STAGE_BYTES;
        unsigned char *sb = sa + GM_A_BYTES;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * 256;                 // 4 matrices x 64 rows x 8 sixteen-byte units
            const int m = idx >> 9, r = (idx >> 3) & 63, ch = idx & 7;
            const double *src = A.M + ((size_t)m * nep + i0 + r) * nep + kc * GM_BK + ch * 2;
            cp_async16(sa + ((size_t)(m * GM_BM + r) * GM_AP + ch * 2) * 8, src);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * 256;                 // 2 vectors x 16 rows x 64 units
            const int a = idx >> 10, r = (idx >> 6) & 15, ch = idx & 63;
            const double *src = (a == 0 ? A.Xp : A.Xn) + (size_t)(kc * GM_BK + r) * npadc + c0 + ch * 2;
            cp_async16(sb + ((size_t)(a * GM_BK + r) * GM_BP + ch * 2) * 8, src);
        }
    };

    double L[2][8][2], G[2][8][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) L[rt][ct][0] = L[rt][ct][1] = G[rt][ct][0] = G[rt][ct][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GM_ST - 1; ++s) {
        if (s < nchunk) load_stage(s, s);
        cp_async_commit();
    }
    const int arow = wr * 16 + (lane >> 2), ak = lane & 3;     // A fragment: row lane/4, k lane%4
    const int bcol = wc * 64 + (lane >> 2), bk = lane & 3;     // B fragment: k lane%4, column lane/4
    for (int kc = 0; kc < nchunk; ++kc) {
        cp_async_wait<GM_ST - 2>();
        __syncthreads();   // stage kc landed for everybody; the stage refilled below was consumed in iteration kc-1
        if (kc + GM_ST - 1 < nchunk) load_stage(kc + GM_ST - 1, (kc + GM_ST - 1) % GM_ST);
        cp_async_commit();
        const double *sa = reinterpret_cast<const double *>(gsm + (size_t)(kc % GM_ST) * GM_STAGE_BYTES);
        const double *sb = sa + GM_A_BYTES / 8;
#pragma unroll
        for (int ks = 0; ks < GM_BK / 4; ++ks) {
            double a[4][2], bp[8], bn[8];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) a[m][rt] = sa[(m * GM_BM + arow + rt * 8) * GM_AP + ks * 4 + ak];
#pragma unroll
            for (int ct = 0; ct < 8; ++ct) {
                bp[ct] = sb[(ks * 4 + bk) * GM_BP + bcol + ct * 8];
                bn[ct] = sb[(GM_BK + ks * 4 + bk) * GM_BP + bcol + ct * 8];
            }
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int ct = 0; ct < 8; ++ct) {
                    dmma884(L[rt][ct][0], L[rt][ct][1], a[0][rt], bp[ct]);   // dE Ke p
                    dmma884(L[rt][ct][0], L[rt][ct][1], a[1][rt], bn[ct]);   // 2dE Kb n
                    dmma884(G[rt][ct][0], G[rt][ct][1], a[2][rt], bp[ct]);   // 2dE Ka p
                    dmma884(G[rt][ct][0], G[rt][ct][1], a[3][rt], bn[ct]);   // dE Ke^T n
                }
        }
    }
    cp_async_wait<0>();
    // ---- epilogue: accumulator element (row lane/4, columns 2*(lane%4) + {0,1}) of every 8 x 8 tile ----
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int i = i0 + wr * 16 + rt * 8 + (lane >> 2);
        if (i >= A.ne) continue;
        const double gth = A.mode == 2 ? A.gth[i] : 0.0;
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) {
            const int q = c0 + wc * 64 + ct * 8 + 2 * (lane & 3);
            if (q >= A.ncell) continue;
            const double2 n2 = *reinterpret_cast<const double2 *>(A.Xn + (size_t)i * npadc + q);
            const double2 p2 = *reinterpret_cast<const double2 *>(A.Xp + (size_t)i * npadc + q);
            A.S[(long long)i * A.ncd + A.c2d[q]] =
                collision_epilogue(A.mode, n2.x, p2.x, G[rt][ct][0], L[rt][ct][0], gth, A.dt);
            if (q + 1 < A.ncell)
                A.S[(long long)i * A.ncd + A.c2d[q + 1]] =
                    collision_epilogue(A.mode, n2.y, p2.y, G[rt][ct][1], L[rt][ct][1], gth, A.dt);
        }
    }
}
