// Per-cell coupled quasiparticle + phonon collision update.
//
// Replaces qpsim/solver.py:794-875 (the Python loop over pixels) and :703-791, :640-665, :686-700 (the
// per-pixel arithmetic):
//   w = max(1 - n/max(rho,1e-30), 0), p = rho*w
//   scattering   gain_i += dE rho_i w_i sum_j Ks[j,i] Np[j,i] n_j      loss_i += dE sum_j Ks[i,j] Np[i,j] p_j
//                Np[i,j] = n_ph[idx_diff[i,j]] (+1 when E_i > E_j), zero diagonal
//   recombination loss_i += 2dE sum_j Kr[i,j] (1+nS[i,j]) n_j          gain_i += 2dE p_i sum_j Kr[i,j] nS[i,j] p_j
//   n'  = max(e^{-mu dt} n + (1-e^{-mu dt})/mu * max(gain + (mu-loss) n, 0), 0),  mu = max(loss, 0)
//   phonons (from the OLD n, p):  a,b[idx_diff] += / -= dE n_i Ks[i,j] p_j ;  a,b[idx_sum] += dE n_i Kr n_j ;
//                b[idx_sum] -= dE p_i Kr p_j ;  n_ph' = max(e^x n_ph + (e^x-1)/b * a, 0), x = clip(b dt, +-80)
//
// Two implementations:
//  * k_collide_struct — uniform energy grid (idx_diff = f(|i-j|), idx_sum = f(i+j)), symmetric kernels, one gap
//    table.  Lanes of a warp are different CELLS, every thread owns 8 rows (or 8 diagonals / anti-diagonals) and
//    sweeps the other index in register tiles, so the kernel matrices are warp-uniform loads and the per-cell
//    vectors are conflict-free shared-memory columns.  Three passes: rows (quasiparticles), diagonals (scattering
//    phonon source), anti-diagonals (recombination / pair-breaking phonon source).
//  * k_collide_generic — arbitrary index maps / per-cell gap tables; shared-memory atomics.
#include "qpb_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace {

// exp as the reference's libm evaluates it.  The reference forms (exp(x) - 1)/b and (1 - exp(-mu dt))/mu
// (solver.py:661, 697), which amplify a 1-ulp difference in exp(x) by 1/|x|.  glibc's exp is correctly rounded
// for all but near-tie inputs; for small |x| the correctly rounded value is fl(1 + expm1(x)), so this form agrees
// with it bit for bit where the amplification matters, while CUDA's 1-ulp exp() would not.
// One code path for every argument (lanes are different cells: a branch on |x| would run both sides): beyond the
// small-|x| range 1 + expm1(x) is exp(x) to an ulp, where nothing amplifies the difference.
__device__ __forceinline__ double exp_ref(double x) {
    return 1.0 + expm1(x);
}

__device__ __forceinline__ double relax_update(double n, double gain, double loss, double dt) {
    const double mu = fmax(loss, 0.0);                       // solver.py:655
    const double P = fmax(gain + (mu - loss) * n, 0.0);      // solver.py:656
    const double decay = exp_ref(-mu * dt);
    const double coeff = mu < 1e-14 ? dt : (1.0 - decay) / mu;
    return fmax(decay * n + coeff * P, 0.0);
}

__device__ __forceinline__ double affine_growth(double y, double a, double b, double dt) {
    const double x = fmin(fmax(b * dt, -80.0), 80.0);        // solver.py:693
    const double ex = exp_ref(x);
    const double coeff = fabs(b) < 1e-14 ? dt : (ex - 1.0) / b;
    return fmax(ex * y + coeff * a, 0.0);
}

// =========================================================================================================
// generic kernel
// =========================================================================================================
struct GenericArgs {
    int ne, nw, ncell, ncd, ngap;
    double *S;
    double *P;
    const int32_t *c2d;
    const double *Ks, *KsT, *Kr, *KrT, *rho;
    const int32_t *gapid;
    const int32_t *idxd, *idxdT, *idxs;
    const int8_t *sign, *signT;
    double dE, dt;
    int scat, rec, update_ph;
};

template <int CG>
__global__ void __launch_bounds__(128) k_collide_generic(GenericArgs A) {
    extern __shared__ double sm[];
    const int ne = A.ne, nw = A.nw;
    const int per = 3 * ne + 3 * nw;
    const int cell0 = blockIdx.x * CG;
    const int tid = threadIdx.x, nt = blockDim.x;
    int gid[CG];
    bool same = true;
#pragma unroll
    for (int cc = 0; cc < CG; ++cc) {
        const int q = min(cell0 + cc, A.ncell - 1);
        gid[cc] = A.gapid ? A.gapid[q] : 0;
        same = same && gid[cc] == gid[0];
    }
#pragma unroll
    for (int cc = 0; cc < CG; ++cc) {
        double *n = sm + cc * per, *w = n + ne, *p = w + ne, *ph = p + ne, *a = ph + nw, *b = a + nw;
        const int q = cell0 + cc;
        const bool live = q < A.ncell;
        const int d = live ? A.c2d[q] : 0;
        for (int i = tid; i < ne; i += nt) {
            const double nv = live ? A.S[(long long)i * A.ncd + d] : 0.0;
            const double r = A.rho[gid[cc] * ne + i];
            const double wv = fmax(1.0 - nv / fmax(r, 1e-30), 0.0);
            n[i] = nv;
            w[i] = wv;
            p[i] = r * wv;
        }
        for (int o = tid; o < nw; o += nt) {
            ph[o] = live ? A.P[(long long)o * A.ncell + q] : 0.0;
            a[o] = 0.0;
            b[o] = 0.0;
        }
    }
    __syncthreads();
    const size_t nn = (size_t)ne * ne;
    for (int i = tid; i < ne; i += nt) {
        double gs[CG], ls[CG], lr[CG], gp[CG];
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) gs[cc] = ls[cc] = lr[cc] = gp[cc] = 0.0;
        for (int j = 0; j < ne; ++j) {
            const size_t e = (size_t)j * ne + i;  // [j][i] of X == X[j,i];  [j][i] of XT == X[i,j]
            double ks_ij = 0.0, ks_ji = 0.0, kr_ij = 0.0;
            int id_ij = 0, id_ji = 0, sg_ij = 0, sg_ji = 0, is_ij = 0;
            if (A.scat) {
                id_ij = A.idxdT[e]; id_ji = A.idxd[e]; sg_ij = A.signT[e]; sg_ji = A.sign[e];
            }
            if (A.rec) is_ij = A.idxs[(size_t)i * ne + j];
#pragma unroll
            for (int cc = 0; cc < CG; ++cc) {
                double *n = sm + cc * per, *p = n + 2 * ne, *ph = p + ne, *a = ph + nw, *b = a + nw;
                if (cc == 0 || !same) {
                    const size_t g = (size_t)gid[cc] * nn;
                    if (A.scat) { ks_ij = A.KsT[g + e]; ks_ji = A.Ks[g + e]; }
                    if (A.rec) kr_ij = A.KrT[g + e];
                }
                const double nj = n[j], pj = p[j], ni = n[i], pi = p[i];
                if (A.scat) {
                    if (i != j) {
                        const double nd_ij = ph[id_ij], nd_ji = ph[id_ji];
                        const double np_ij = sg_ij > 0 ? 1.0 + nd_ij : nd_ij;
                        const double np_ji = sg_ji > 0 ? 1.0 + nd_ji : nd_ji;
                        gs[cc] += ks_ji * np_ji * nj;
                        ls[cc] += ks_ij * np_ij * pj;
                    }
                    if (A.update_ph && sg_ij != 0) {
                        const double Sv = A.dE * (ni * ks_ij * pj);
                        if (sg_ij > 0) { atomicAdd(&a[id_ij], Sv); atomicAdd(&b[id_ij], Sv); }
                        else atomicAdd(&b[id_ij], -Sv);
                    }
                }
                if (A.rec) {
                    const double nS = ph[is_ij];
                    lr[cc] += kr_ij * (1.0 + nS) * nj;
                    gp[cc] += kr_ij * nS * pj;
                    if (A.update_ph) {
                        const double R = A.dE * (ni * kr_ij * nj);
                        const double Bk = A.dE * (pi * kr_ij * pj);
                        atomicAdd(&a[is_ij], R);
                        atomicAdd(&b[is_ij], R - Bk);
                    }
                }
            }
        }
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) {
            const int q = cell0 + cc;
            if (q >= A.ncell) continue;
            double *n = sm + cc * per, *w = n + ne, *p = w + ne;
            const double r = A.rho[gid[cc] * ne + i];
            const double gain = A.dE * r * w[i] * gs[cc] + 2.0 * A.dE * p[i] * gp[cc];
            const double loss = A.dE * ls[cc] + 2.0 * A.dE * lr[cc];
            A.S[(long long)i * A.ncd + A.c2d[q]] = relax_update(n[i], gain, loss, A.dt);
        }
    }
    if (!A.update_ph || !(A.scat || A.rec)) return;
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < CG; ++cc) {
        const int q = cell0 + cc;
        if (q >= A.ncell) continue;
        double *ph = sm + cc * per + 3 * ne, *a = ph + nw, *b = a + nw;
        for (int o = tid; o < nw; o += nt)
            A.P[(long long)o * A.ncell + q] = affine_growth(ph[o], a[o], b[o], A.dt);
    }
}

#include "qpb_collide_struct.cuh"
#include "qpb_collide_gemm.cuh"
#include "qpb_collide_uniform.cuh"

struct StructTables {
    double2 *K2 = nullptr;
    double *KsD = nullptr, *KrA = nullptr, *rho = nullptr;
    int nep = 0;
};

}  // namespace

// The structured tables live in the scratch allocation of the context: [K2 | KsD | KrA | rho].
static StructTables carve_tables(qpb_ctx *c) {
    StructTables t;
    const int nep = ((c->cfg.ne + TI - 1) / TI) * TI;
    const size_t ng = (size_t)c->cfg.ngap;
    t.nep = nep;
    char *base = (char *)c->d_scratch;
    t.K2 = (double2 *)base;
    base += sizeof(double2) * ng * nep * nep;
    t.KsD = (double *)base;
    base += sizeof(double) * ng * nep * nep;
    t.KrA = (double *)base;
    base += sizeof(double) * ng * 2 * nep * nep;
    t.rho = (double *)base;
    return t;
}

// cells per CTA of the structured kernel for a padded energy grid (0: does not fit)
static int struct_cells_per_cta(int nep) {
    const size_t cap = 227 * 1024;
    auto need = [&](int cc, int nt) {
        return sizeof(double) * (size_t)cc * (2 * (nep + PADF + PADB) + 3 * nep) +
               (size_t)(nt / 32) * (32 / cc) * NSTAGE * (TI * TJ * 16);
    };
    if (need(32, 512) <= cap) return 32;
    if (need(16, 512) <= cap) return 16;
    if (need(8, 256) <= cap) return 8;
    if (need(4, 128) <= cap) return 4;
    return 0;
}

static void dev_free_groups(qpb_ctx *c) {
    if (c->d_cperm) qpb_dev_free(c->d_cperm);
    if (c->d_ggid) qpb_dev_free(c->d_ggid);
    c->d_cperm = c->d_ggid = nullptr;
    c->ngroups = 0;
    c->group_cc = 0;
}

int qpbk_collision_setup(qpb_ctx *c) {
    const auto &cf = c->cfg;
    const bool scat = cf.flags & QPB_F_SCATTERING, rec = cf.flags & QPB_F_RECOMBINATION;
    dev_free_groups(c);
    if (!(scat || rec)) return QPB_OK;
    const int ne = cf.ne, ng = cf.ngap;
    const int nep = ((ne + TI - 1) / TI) * TI;
    if (!c->structured || nep > NEPMAX || cf.nw > 32767) {
        c->structured = false;
        return QPB_OK;
    }
    const size_t per_gap = sizeof(double2) * (size_t)nep * nep + sizeof(double) * (size_t)3 * nep * nep +
                           sizeof(double) * (size_t)nep;
    const size_t bytes = per_gap * ng;
    // Several gap tables: the structured kernel needs the CC cells of a CTA to share one table, so the cells are
    // regrouped by gap id (padded to CC per group).  Not worth it when the padding more than doubles the work or the
    // tables outgrow 2 GB: the generic kernel (per-cell table lookups) takes those.
    std::vector<int32_t> cperm, ggid;
    if (ng > 1) {
        const int cc = struct_cells_per_cta(nep);
        if (cc == 0 || c->h_gapid.size() != (size_t)cf.ncell || bytes > ((size_t)2 << 30)) {
            c->structured = false;
            return QPB_OK;
        }
        std::vector<std::vector<int32_t>> members(ng);
        for (int q = 0; q < cf.ncell; ++q) members[c->h_gapid[q]].push_back(q);
        for (int g = 0; g < ng; ++g)
            for (size_t o = 0; o < members[g].size(); o += cc) {
                ggid.push_back(g);
                for (int l = 0; l < cc; ++l)
                    cperm.push_back(o + l < members[g].size() ? members[g][o + l] : -1);
            }
        if (cperm.size() > (size_t)2 * cf.ncell + (size_t)cc) {
            c->structured = false;
            return QPB_OK;
        }
        c->group_cc = cc;
    }
    if (c->d_scratch) qpb_dev_free(c->d_scratch);
    c->d_scratch = nullptr;
    QPB_CUDA(qpb_dev_malloc((void **)&c->d_scratch, bytes));
    c->scratch_bytes = bytes;
    // pull the uploaded matrices back (they are small) and build the padded / skewed copies, one set per gap table
    const size_t nn = (size_t)ne * ne;
    std::vector<double> Ks(nn * ng, 0.0), Kr(nn * ng, 0.0), rho((size_t)ne * ng);
    if (scat) QPB_CUDA(cudaMemcpy(Ks.data(), c->d_Ks, sizeof(double) * nn * ng, cudaMemcpyDeviceToHost));
    if (rec) QPB_CUDA(cudaMemcpy(Kr.data(), c->d_Kr, sizeof(double) * nn * ng, cudaMemcpyDeviceToHost));
    QPB_CUDA(cudaMemcpy(rho.data(), c->d_rho, sizeof(double) * ne * ng, cudaMemcpyDeviceToHost));
    const size_t np2 = (size_t)nep * nep;
    std::vector<double2> K2(np2 * ng, make_double2(0.0, 0.0));
    std::vector<double> KsD(np2 * ng, 0.0), KrA(2 * np2 * ng, 0.0), rhop((size_t)nep * ng, 0.0);
    const double dE = cf.dE;
    for (int g = 0; g < ng; ++g) {
        const double *ks = Ks.data() + nn * g, *kr = Kr.data() + nn * g;
        for (int i = 0; i < ne; ++i) {
            rhop[(size_t)g * nep + i] = rho[(size_t)g * ne + i];
            for (int j = 0; j < ne; ++j) {
                K2[np2 * g + struct_tile_index(i, j, nep)] = make_double2(dE * ks[(size_t)i * ne + j], 2.0 * dE * kr[(size_t)i * ne + j]);
                if (i >= j) KsD[np2 * g + struct_tile_index(i - j, j, nep)] = dE * ks[(size_t)i * ne + j];
                if (j <= i) {
                    const double wgt = j < i ? 2.0 : 1.0;
                    KrA[2 * np2 * g + struct_tile_index(i + j, j, nep)] = wgt * dE * kr[(size_t)i * ne + j];
                }
            }
        }
    }
    StructTables t = carve_tables(c);
    QPB_CUDA(cudaMemcpy(t.K2, K2.data(), sizeof(double2) * K2.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(t.KsD, KsD.data(), sizeof(double) * KsD.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(t.KrA, KrA.data(), sizeof(double) * KrA.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(t.rho, rhop.data(), sizeof(double) * rhop.size(), cudaMemcpyHostToDevice));
    if (ng > 1) {
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_cperm, sizeof(int32_t) * cperm.size()));
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_ggid, sizeof(int32_t) * ggid.size()));
        QPB_CUDA(cudaMemcpy(c->d_cperm, cperm.data(), sizeof(int32_t) * cperm.size(), cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(c->d_ggid, ggid.data(), sizeof(int32_t) * ggid.size(), cudaMemcpyHostToDevice));
        c->ngroups = (int)ggid.size();
    }
    return QPB_OK;
}

template <int CC>
struct StructCfg {
    // 16 cells per CTA (256 / 384 bins) run with 16 warps as well: two warp slots per row block keep every slot busy in
    // one round of each pass and give each scheduler four warps to hide the FP64 latency with (ncu at 256 bins with
    // 8 warps: 12.5 % warps active, top stall "wait")
    static constexpr int NT = CC >= 16 ? 512 : (CC >= 8 ? 256 : 128);
    static size_t smem(int nep, int nt = NT) {
        return sizeof(double) * (size_t)CC * (2 * (nep + PADF + PADB) + 3 * nep) +
               (size_t)(nt / 32) * (32 / CC) * NSTAGE * (TI * TJ * 16);
    }
};

// Energy grids of at most 64 (padded) bins have no more than 8 row blocks: a 512-thread CTA would leave half of its warp
// slots without a row block in every pass.  They run 256-thread CTAs of 32 cells, two per SM (2 x 105 KB of shared
// memory at 64 bins), so that one CTA stages its columns while the other one computes.  QPB_COLL_NT=512 restores the
// wide CTA (A/B runs).
constexpr int STRUCT_SMALL_NEP = 64;
constexpr int STRUCT_SMALL_NT = 256;
static bool struct_small_grid(int nep) {
    if (nep > STRUCT_SMALL_NEP) return false;
    const char *e = getenv("QPB_COLL_NT");
    return !(e && atoi(e) == 512);
}

template <int CC, int NT, bool SC, bool RC, bool PH>
static int launch_struct(qpb_ctx *c, const StructArgs &A, size_t smem) {
    auto kern = k_collide_struct<CC, NT, SC, RC, PH>;
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    const int blocks = A.cperm ? c->ngroups : (A.ncell + CC - 1) / CC;
    kern<<<blocks, NT, smem, c->stream>>>(A);
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

template <int CC, int NT>
static int dispatch_struct_nt(qpb_ctx *c, const StructArgs &A, size_t smem, bool sc, bool rc, bool ph) {
    if (sc && rc) return ph ? launch_struct<CC, NT, true, true, true>(c, A, smem) : launch_struct<CC, NT, true, true, false>(c, A, smem);
    if (sc) return ph ? launch_struct<CC, NT, true, false, true>(c, A, smem) : launch_struct<CC, NT, true, false, false>(c, A, smem);
    return ph ? launch_struct<CC, NT, false, true, true>(c, A, smem) : launch_struct<CC, NT, false, true, false>(c, A, smem);
}

template <int CC>
static int dispatch_struct(qpb_ctx *c, const StructArgs &A, size_t smem, bool sc, bool rc, bool ph) {
    if (CC == 32 && struct_small_grid(A.nep))
        return dispatch_struct_nt<CC, CC == 32 ? STRUCT_SMALL_NT : StructCfg<CC>::NT>(
            c, A, StructCfg<CC>::smem(A.nep, STRUCT_SMALL_NT), sc, rc, ph);
    return dispatch_struct_nt<CC, StructCfg<CC>::NT>(c, A, smem, sc, rc, ph);
}

template <int CG>
static int launch_generic(qpb_ctx *c, const GenericArgs &A, size_t smem) {
    auto kern = k_collide_generic<CG>;
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    const int blocks = (A.ncell + CG - 1) / CG;
    kern<<<blocks, 128, smem, c->stream>>>(A);
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

// ---- frozen, cell-independent phonons: packed effective kernels (solver.py:726-743 with a shared n_ph) -----------
int qpbk_uniform_setup(qpb_ctx *c, const double *n_ph, bool per_bin) {
    const auto &cf = c->cfg;
    const bool scat = cf.flags & QPB_F_SCATTERING, rec = cf.flags & QPB_F_RECOMBINATION;
    c->uniform_ph = false;
    c->gemm_ready = false;
    if (!(cf.flags & QPB_F_FREEZE_PHONONS) || !(scat || rec) || cf.ngap != 1 || !n_ph || !c->have_coll) return QPB_OK;
    if (getenv("QPB_NO_UNIFORM") && getenv("QPB_NO_UNIFORM")[0] == '1') return QPB_OK;
    const int ne = cf.ne, nw = cf.nw;
    const int nep = ((ne + TI - 1) / TI) * TI;
    std::vector<double> ph(nw);
    for (int o = 0; o < nw; ++o) {
        if (per_bin) {   // the caller passed one occupation per phonon bin
            ph[o] = n_ph[o];
            continue;
        }
        const double *row = n_ph + (size_t)o * cf.ncell;
        ph[o] = row[0];
        for (int q = 1; q < cf.ncell; ++q)
            if (row[q] != ph[o]) return QPB_OK;   // occupations differ between cells: general kernels
    }
    std::vector<double> Ks((size_t)ne * ne, 0.0), Kr((size_t)ne * ne, 0.0), rho(ne);
    std::vector<int32_t> idd((size_t)ne * ne), ids((size_t)ne * ne);
    std::vector<int8_t> sg((size_t)ne * ne);
    if (scat) QPB_CUDA(cudaMemcpy(Ks.data(), c->d_Ks, sizeof(double) * ne * ne, cudaMemcpyDeviceToHost));
    if (rec) QPB_CUDA(cudaMemcpy(Kr.data(), c->d_Kr, sizeof(double) * ne * ne, cudaMemcpyDeviceToHost));
    QPB_CUDA(cudaMemcpy(rho.data(), c->d_rho, sizeof(double) * ne, cudaMemcpyDeviceToHost));
    QPB_CUDA(cudaMemcpy(idd.data(), c->d_idxd, sizeof(int32_t) * ne * ne, cudaMemcpyDeviceToHost));
    QPB_CUDA(cudaMemcpy(ids.data(), c->d_idxs, sizeof(int32_t) * ne * ne, cudaMemcpyDeviceToHost));
    QPB_CUDA(cudaMemcpy(sg.data(), c->d_sign, sizeof(int8_t) * ne * ne, cudaMemcpyDeviceToHost));
    const double dE = cf.dE;
    auto ke = [&](int i, int j) {   // dE * Ks[i,j] * Np[i,j], zero diagonal (solver.py:730-732)
        if (i == j || !scat) return 0.0;
        const double nd = ph[idd[(size_t)i * ne + j]];
        return dE * (Ks[(size_t)i * ne + j] * (sg[(size_t)i * ne + j] > 0 ? 1.0 + nd : nd));
    };
    std::vector<double> K4((size_t)4 * nep * nep, 0.0), rhop(nep, 0.0);
    for (int i = 0; i < ne; ++i) {
        rhop[i] = rho[i];
        for (int j = 0; j < ne; ++j) {
            double *o = &K4[((size_t)i * nep + j) * 4];
            o[0] = ke(i, j);
            o[1] = ke(j, i);
            if (rec) {
                const double ns = ph[ids[(size_t)i * ne + j]], kr = Kr[(size_t)i * ne + j];
                o[2] = 2.0 * dE * (kr * (1.0 + ns));
                o[3] = 2.0 * dE * (kr * ns);
            }
        }
    }
    if (c->d_K4) qpb_dev_free(c->d_K4);
    c->d_K4 = nullptr;
    QPB_CUDA(qpb_dev_malloc((void **)&c->d_K4, sizeof(double) * (K4.size() + nep)));
    QPB_CUDA(cudaMemcpy(c->d_K4, K4.data(), sizeof(double) * K4.size(), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(c->d_K4 + K4.size(), rhop.data(), sizeof(double) * nep, cudaMemcpyHostToDevice));
    c->uniform_ph = true;
    c->h_ph_bins = ph;
    // ---- tensor-core form (qpb_collide_gemm.cuh) where the bin count makes the products a real GEMM ----
    if (c->d_Mg) qpb_dev_free(c->d_Mg);
    if (c->d_Xn) qpb_dev_free(c->d_Xn);
    if (c->d_Xp) qpb_dev_free(c->d_Xp);
    c->d_Mg = c->d_Xn = c->d_Xp = nullptr;
    c->gemm_ready = false;
    int min_ne = 64;
    if (const char *e = getenv("QPB_GEMM_MIN_NE")) min_ne = atoi(e);
    if (ne >= min_ne && !(getenv("QPB_NO_GEMM") && getenv("QPB_NO_GEMM")[0] == '1')) {
        const int ng = ((ne + GM_BM - 1) / GM_BM) * GM_BM;
        const size_t npadc = ((size_t)cf.ncell + GM_BN - 1) / GM_BN * GM_BN;
        std::vector<double> M((size_t)4 * ng * ng + ng, 0.0);
        for (int i = 0; i < ne; ++i) {
            M[(size_t)4 * ng * ng + i] = rho[i];
            for (int j = 0; j < ne; ++j) {
                const double *o = &K4[((size_t)i * nep + j) * 4];   // (dE Ke_ij, dE Ke_ji, 2dE Kb_ij, 2dE Ka_ij)
                M[((size_t)0 * ng + i) * ng + j] = o[0];
                M[((size_t)1 * ng + i) * ng + j] = o[2];
                M[((size_t)2 * ng + i) * ng + j] = o[3];
                M[((size_t)3 * ng + i) * ng + j] = o[1];
            }
        }
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_Mg, sizeof(double) * M.size()));
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_Xn, sizeof(double) * ng * npadc));
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_Xp, sizeof(double) * ng * npadc));
        QPB_CUDA(cudaMemcpy(c->d_Mg, M.data(), sizeof(double) * M.size(), cudaMemcpyHostToDevice));
        // the padding stays zero for good (ordered on the context's stream: the legacy stream does not order it)
        QPB_CUDA(cudaMemsetAsync(c->d_Xn, 0, sizeof(double) * ng * npadc, c->stream));
        QPB_CUDA(cudaMemsetAsync(c->d_Xp, 0, sizeof(double) * ng * npadc, c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));
        c->gemm_nep = ng;
        c->gemm_npadc = (long long)npadc;
        c->gemm_ready = true;
    }
    return QPB_OK;
}

static int launch_gemm_args(qpb_ctx *c, const GemmArgs &G);

static int launch_gemm(qpb_ctx *c, double dt) {
    const auto &cf = c->cfg;
    GemmArgs G;
    G.ne = cf.ne; G.nep = c->gemm_nep; G.ncell = cf.ncell; G.npadc = (int)c->gemm_npadc; G.ncd = c->ncd;
    G.S = c->d_S; G.c2d = c->d_cell2dense; G.M = c->d_Mg; G.Xn = c->d_Xn; G.Xp = c->d_Xp;
    G.rho = c->d_Mg + (size_t)4 * G.nep * G.nep;
    G.dt = dt;
    G.mode = 0;
    G.gth = nullptr;
    return launch_gemm_args(c, G);
}

static int launch_gemm_args(qpb_ctx *c, const GemmArgs &G) {
    const auto &cf = c->cfg;
    static bool configured = false;
    const size_t smem = (size_t)GM_ST * GM_STAGE_BYTES;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(k_collide_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long long total = (long long)cf.ne * cf.ncell;
    const int pack_blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    k_gemm_pack<<<pack_blocks, 256, 0, c->stream>>>(G);
    QPB_CHECK_LAUNCH();
    // cell blocks along x so that CTAs running together share the row block's matrix tiles in L2
    dim3 grid((unsigned)(G.npadc / GM_BN), (unsigned)(G.nep / GM_BM));
    k_collide_gemm<<<grid, 256, smem, c->stream>>>(G);
    QPB_CHECK_LAUNCH();
    c->diag.kernel_launches++;
    return QPB_OK;
}

template <int CC, int NT>
static int launch_uniform(qpb_ctx *c, const UniformArgs &A) {
    auto kern = k_collide_uniform<CC, NT>;
    const size_t smem = sizeof(double) * (size_t)2 * A.nep * CC + (size_t)(NT / 32) * (32 / CC) * NSTAGE * (TI * TJ * 32);
    static bool configured = false;
    if (!configured) {
        QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    if (smem > 227 * 1024) return 1;
    kern<<<(A.ncell + CC - 1) / CC, NT, smem, c->stream>>>(A);
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

// ---- fixed-bath forward-Euler forms (solver.py:551-605): the same products with other matrices and epilogues ------
int qpbk_euler_step(qpb_ctx *c, int kind, const double *K, const double *vec, double dt) {
    const auto &cf = c->cfg;
    const int ne = cf.ne;
    const double dE = cf.dE;
    int min_ne = 64;
    if (const char *e = getenv("QPB_GEMM_MIN_NE")) min_ne = atoi(e);
    const bool gemm = ne >= min_ne && !(getenv("QPB_NO_GEMM") && getenv("QPB_NO_GEMM")[0] == '1');
    c->diag.kernel_launches++;
    ScopedTimer tm(c, 2);
    if (gemm) {
        const int ng = ((ne + GM_BM - 1) / GM_BM) * GM_BM;
        const size_t npadc = ((size_t)cf.ncell + GM_BN - 1) / GM_BN * GM_BN;
        // [4 matrices | rho | gth], operands Xn, Xp
        std::vector<double> M((size_t)4 * ng * ng + 2 * ng, 0.0);
        double *rho = M.data() + (size_t)4 * ng * ng, *gth = rho + ng;
        for (int i = 0; i < ne; ++i) {
            if (kind == 1) rho[i] = vec[i]; else gth[i] = vec[i];
            for (int j = 0; j < ne; ++j) {
                if (kind == 1) {
                    M[((size_t)0 * ng + i) * ng + j] = dE * K[(size_t)i * ne + j];        // L: dE Ks_ij p_j
                    M[((size_t)3 * ng + i) * ng + j] = dE * K[(size_t)j * ne + i];        // G: dE Ks_ji n_j
                } else {
                    M[((size_t)1 * ng + i) * ng + j] = 2.0 * dE * K[(size_t)i * ne + j];  // L: 2dE Kr_ij n_j
                }
            }
        }
        const size_t need = sizeof(double) * (M.size() + 2 * (size_t)ng * npadc);
        if (c->euler_bytes < need) {
            if (c->d_euler) qpb_dev_free(c->d_euler);
            c->d_euler = nullptr;
            QPB_CUDA(qpb_dev_malloc((void **)&c->d_euler, need));
            c->euler_bytes = need;
            QPB_CUDA(cudaMemsetAsync(c->d_euler, 0, need, c->stream));   // operand padding stays zero
        }
        QPB_CUDA(cudaMemcpyAsync(c->d_euler, M.data(), sizeof(double) * M.size(), cudaMemcpyHostToDevice, c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));   // M is a local vector
        GemmArgs G;
        G.ne = ne; G.nep = ng; G.ncell = cf.ncell; G.npadc = (int)npadc; G.ncd = c->ncd;
        G.S = c->d_S; G.c2d = c->d_cell2dense; G.M = c->d_euler;
        G.rho = c->d_euler + (size_t)4 * ng * ng;
        G.gth = G.rho + ng;
        G.Xn = c->d_euler + M.size();
        G.Xp = G.Xn + (size_t)ng * npadc;
        G.dt = dt;
        G.mode = kind;
        return launch_gemm_args(c, G);
    }
    // fused GEMV: packed [nep][nep][4] = (L<-p, G<-n, L<-n, G<-p) + rho + gth
    const int nep = ((ne + TI - 1) / TI) * TI;
    std::vector<double> K4((size_t)4 * nep * nep + 2 * nep, 0.0);
    double *rho = K4.data() + (size_t)4 * nep * nep, *gth = rho + nep;
    for (int i = 0; i < ne; ++i) {
        if (kind == 1) rho[i] = vec[i]; else gth[i] = vec[i];
        for (int j = 0; j < ne; ++j) {
            double *o = &K4[((size_t)i * nep + j) * 4];
            if (kind == 1) {
                o[0] = dE * K[(size_t)i * ne + j];
                o[1] = dE * K[(size_t)j * ne + i];
            } else {
                o[2] = 2.0 * dE * K[(size_t)i * ne + j];
            }
        }
    }
    const size_t need = sizeof(double) * K4.size();
    if (c->euler_bytes < need) {
        if (c->d_euler) qpb_dev_free(c->d_euler);
        c->d_euler = nullptr;
        QPB_CUDA(qpb_dev_malloc((void **)&c->d_euler, need));
        c->euler_bytes = need;
    }
    QPB_CUDA(cudaMemcpyAsync(c->d_euler, K4.data(), need, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    UniformArgs U;
    U.ne = ne; U.nep = nep; U.ncell = cf.ncell; U.ncd = c->ncd;
    U.S = c->d_S; U.c2d = c->d_cell2dense;
    U.K4 = reinterpret_cast<const double4 *>(c->d_euler);
    U.rho = c->d_euler + (size_t)4 * nep * nep;
    U.gth = U.rho + nep;
    U.dt = dt;
    U.mode = kind;
    int rc = launch_uniform<32, 512>(c, U);
    if (rc <= 0) return rc;
    rc = launch_uniform<8, 256>(c, U);
    if (rc <= 0) return rc;
    qpb_set_error("qpb_euler_step: %d energy bins do not fit the fused GEMV kernel", ne);
    return QPB_E_INVALID;
}

int qpbk_collide(qpb_ctx *c, double dt, int xmode) {
    const auto &cf = c->cfg;
    const bool scat = cf.flags & QPB_F_SCATTERING, rec = cf.flags & QPB_F_RECOMBINATION;
    const bool ph = !(cf.flags & QPB_F_FREEZE_PHONONS);
    ScopedTimer tm(c, 2);
    c->diag.kernel_launches++;
    const size_t smem_cap = 227 * 1024;
    if (xmode != 0 && !(c->structured && (scat || rec) && c->x_nranks > 0 && !(c->uniform_ph && !ph))) {
        qpb_set_error("fused exchange needs the structured collision kernel and qpb_set_exchange");
        return QPB_E_INVALID;
    }
    if (c->uniform_ph && !ph && (scat || rec)) {
        UniformArgs U;
        U.ne = cf.ne; U.nep = ((cf.ne + TI - 1) / TI) * TI; U.ncell = cf.ncell; U.ncd = c->ncd;
        U.S = c->d_S; U.c2d = c->d_cell2dense;
        U.K4 = reinterpret_cast<const double4 *>(c->d_K4);
        U.rho = c->d_K4 + (size_t)4 * U.nep * U.nep;
        U.dt = dt;
        U.mode = 0;
        U.gth = nullptr;
        if (c->gemm_ready) return launch_gemm(c, dt);
        int rc = launch_uniform<32, 512>(c, U);
        if (rc <= 0) return rc;
        rc = launch_uniform<8, 256>(c, U);
        if (rc <= 0) return rc;
        // too many energy bins for the shared-memory columns: general kernels below
    }
    if (c->structured && (scat || rec)) {
        StructTables t = carve_tables(c);
        StructArgs A;
        A.ne = cf.ne; A.nep = t.nep; A.nw = cf.nw; A.ncell = cf.ncell; A.ncd = c->ncd;
        A.S = c->d_S; A.P = c->d_P; A.c2d = c->d_cell2dense;
        A.K2 = t.K2; A.KsD = t.KsD; A.KrA = t.KrA; A.rho = t.rho;
        for (int k = 0; k < NEPMAX; ++k) A.dmap[k] = A.mofk[k] = -1;
        for (int m = 0; m < 2 * NEPMAX; ++m) A.smap[m] = A.kofm[m] = -1;
        for (int k = 0; k < cf.ne; ++k) {
            A.dmap[k] = (int16_t)c->h_dmap[k];
            A.mofk[k] = (int16_t)c->h_mof[c->h_dmap[k]];
        }
        for (int m = 0; m < 2 * cf.ne - 1; ++m) {
            A.smap[m] = (int16_t)c->h_smap[m];
            A.kofm[m] = (int16_t)c->h_kof[c->h_smap[m]];
        }
        A.dt = dt;
        A.cperm = c->d_cperm;
        A.ggid = c->d_ggid;
        A.xmode = xmode;
        A.xncd = c->x_ncd;
        A.xdense = c->d_xdense;
        for (int r = 0; r < QPB_MAX_RANKS; ++r) A.xpeer[r] = c->x_peer[r];
        for (int i = 0; i < NEPMAX; ++i) A.route[i] = 0;
        if (xmode != 0)
            for (int i = 0; i < cf.ne; ++i) A.route[i] = c->x_route[i];
        // One 32-cell, 512-thread CTA per SM measured faster (1.19 ms at C2) than two 16-cell, 256-thread CTAs
        // (1.47 ms: half-warps read different kernel-matrix tiles); QPB_COLL_CC=16 selects the narrow variant
        const char *ecc = getenv("QPB_COLL_CC");
        const bool narrow = ecc && ecc[0] == '1';
        if (narrow && 2 * (StructCfg<16>::smem(t.nep) + 1024) <= 228 * 1024)
            return dispatch_struct<16>(c, A, StructCfg<16>::smem(t.nep), scat, rec, ph);
        if (c->d_cperm) {   // regrouped cells: the group width was fixed at setup
            switch (c->group_cc) {
                case 32: return dispatch_struct<32>(c, A, StructCfg<32>::smem(t.nep), scat, rec, ph);
                case 16: return dispatch_struct<16>(c, A, StructCfg<16>::smem(t.nep), scat, rec, ph);
                case 8: return dispatch_struct<8>(c, A, StructCfg<8>::smem(t.nep), scat, rec, ph);
                default: return dispatch_struct<4>(c, A, StructCfg<4>::smem(t.nep), scat, rec, ph);
            }
        }
        if (StructCfg<32>::smem(t.nep) <= smem_cap) return dispatch_struct<32>(c, A, StructCfg<32>::smem(t.nep), scat, rec, ph);
        if (StructCfg<16>::smem(t.nep) <= smem_cap) return dispatch_struct<16>(c, A, StructCfg<16>::smem(t.nep), scat, rec, ph);
        if (StructCfg<8>::smem(t.nep) <= smem_cap) return dispatch_struct<8>(c, A, StructCfg<8>::smem(t.nep), scat, rec, ph);
        if (StructCfg<4>::smem(t.nep) <= smem_cap) return dispatch_struct<4>(c, A, StructCfg<4>::smem(t.nep), scat, rec, ph);
        // energy grids too large for the shared-memory columns fall through to the generic kernel
    }
    GenericArgs G;
    G.ne = cf.ne; G.nw = cf.nw; G.ncell = cf.ncell; G.ncd = c->ncd; G.ngap = cf.ngap;
    G.S = c->d_S; G.P = c->d_P; G.c2d = c->d_cell2dense;
    G.Ks = c->d_Ks; G.KsT = c->d_KsT; G.Kr = c->d_Kr; G.KrT = c->d_KrT; G.rho = c->d_rho;
    G.gapid = c->d_gapid; G.idxd = c->d_idxd; G.idxdT = c->d_idxdT; G.idxs = c->d_idxs;
    G.sign = c->d_sign; G.signT = c->d_signT;
    G.dE = cf.dE; G.dt = dt; G.scat = scat; G.rec = rec; G.update_ph = ph;
    auto need = [&](int cg) { return sizeof(double) * (size_t)cg * (3 * (size_t)cf.ne + 3 * (size_t)cf.nw); };
    if (need(4) <= 96 * 1024) return launch_generic<4>(c, G, need(4));
    if (need(2) <= smem_cap) return launch_generic<2>(c, G, need(2));
    if (need(1) <= smem_cap) return launch_generic<1>(c, G, need(1));
    qpb_set_error("qpb_collide: %d energy bins / %d phonon bins exceed the shared-memory budget", cf.ne, cf.nw);
    return QPB_E_INVALID;
}
