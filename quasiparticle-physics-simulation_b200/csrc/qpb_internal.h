// Internal declarations shared by the qpb translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "qpb.h"

// per-cell geometry flags (dense grid)
#define QPB_LK_L 1u   // cell (y,x) linked to (y,x-1)
#define QPB_LK_R 2u   // linked to (y,x+1)
#define QPB_LK_U 4u   // linked to (y-1,x)
#define QPB_LK_D 8u   // linked to (y+1,x)
#define QPB_IN   16u  // cell is inside the mask

void qpb_set_error(const char *fmt, ...);

// cached device allocations (qpb_api.cu): drop-in for cudaMalloc / cudaFree of library-owned buffers
cudaError_t qpb_dev_malloc(void **p, size_t bytes);
void qpb_dev_free(void *p);

#define QPB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            qpb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,     \
                          __LINE__, cudaGetErrorString(_e));                                \
            return QPB_E_CUDA;                                                              \
        }                                                                                   \
    } while (0)

#define QPB_CHECK_LAUNCH() QPB_CUDA(cudaGetLastError())

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// launch plan of the persistent TMA-pipelined sweeps (qpb_sweep_pipe.cu)
struct PipePlan {
    bool x_ok = false, y_ok = false;
    int nsm = 148;
    int nt = 256;                           // threads per CTA of both kernels
    int x_qp = 0, x_ns = 0, x_tpb = 0;      // lanes per row (power of two), input stages, tiles per bin
    int y_cw = 0, y_ns = 0, y_tpb = 0;      // columns per strip, input stages, tiles per bin
    // segmentation of long lines (PipeArgs::qs ...): chunks per tile, interior chunks, halo chunks, segments per line
    int x_qs = 0, x_qi = 0, x_halo = 0, x_nseg = 1;
    int y_qs = 0, y_qi = 0, y_halo = 0, y_nseg = 1;
    bool x_inplace = false, y_inplace = false;   // output over an input tile of the stage (PipeArgs::inplace)
    std::vector<unsigned char> xmaps, ymaps;  // host copies of the CUtensorMap triples
};

// one prepared Crank-Nicolson solve (a step length)
struct DiffSlot {
    bool ready = false;
    double dt = 0.0;
    int mode = 0;              // 0 iterate, 1 direct sweep along x, 2 direct sweep along y
    bool commuting = false;
    int jmax = 0;              // row length of the shift table
    int launch_iters = 0;      // iterations launched before the first host check
    int known_iters = 0;       // iterations the previous solve needed (0: unknown)
    long long solves = 0;      // solves done with this slot
    std::vector<double> a_bin; // 0.5*dt*D_i/dx^2 per bin (uniform D)
    std::vector<int> jlen;     // shifts per bin
    double *d_a = nullptr;     // [ne]
    double *d_shift = nullptr; // [ne][jmax]
    int *d_jlen = nullptr;     // [ne]
    double *d_tol = nullptr;   // [ne] componentwise residual tolerance per bin of the sweep iteration
    double *d_tolk = nullptr;  // [ne] max-norm tolerance of the Krylov path, floored at what fp64 can resolve
    int *d_known = nullptr;    // [ne] iterations the previous solve needed per bin (0: unknown)
    // direct spectral solve on full rectangles with reflective left / right walls (qpb_spectral.cu)
    bool spectral = false;
    int sp_logn = 0;
    double2 *d_sp_tw = nullptr, *d_sp_tw2 = nullptr;   // exp(-2 pi i k / nx), exp(-i pi k / (2 nx))
    double *d_sp_lam = nullptr, *d_sp_bcy = nullptr;   // eigenvalues of Gx [nx], wall diagonal of every row [ny]
    double *d_sp_piv = nullptr;                        // [ne][sp_T][nx] first pivots of the Thomas pass (sp_T = 0: none)
    int sp_T = 0;
    double sp_wall_last = 0.0, sp_wall0 = 0.0;
    bool sp_fused = false;             // the Thomas pass forms the right-hand side itself (no k_build_rhs pass, no b)
    double *d_sp_srchat = nullptr;     // [ny][nx] cosine transform of the boundary sources (null: they are all zero)
    bool krylov = false;       // stiff non-commuting solve: preconditioned BiCGStab instead of the sweep iteration (qpb_krylov.cu)
    double *d_kshift = nullptr; // [ne] shift sqrt(lo hi) of the line-solve preconditioner
    // variable-D coefficient fields (dense, per bin): links to the left / up neighbour, boundary diagonals
    double *d_ex = nullptr, *d_ey = nullptr, *d_gbx = nullptr, *d_gby = nullptr;
    // source term dt*D*s, dense [ncd] (uniform: multiplied by D_i on the fly) or [ne][ncd] (variable)
    double *d_src = nullptr;
    // fast path tables (uniform D, chunked sweeps)
    struct FastDir {
        int n = 0, S = 0, Q = 0, npad = 0, nclass = 0;
        int carry_depth = 0;       // chunks a carry must be propagated through (products of g below 1e-18 beyond)
        int *d_cls = nullptr;      // class of every line
        double *d_tab = nullptr;   // [ne][jmax][nclass][npad] pivot reciprocals m (x: chunk-interleaved), 0 outside the mask
        double *d_tabg = nullptr;  // same layout: g_t = e_{t+1} m_t
        bool use_tma = false;      // x sweep staged by TMA (nx % 16 == 0, nx <= 512)
        std::vector<unsigned char> tma;  // host copy of the CUtensorMap triple passed as a kernel parameter
    } fx, fy;
    bool fast = false;
    PipePlan pipe;
    // bin-resident solve (qpb_resident.cu): a cluster of CS CTAs keeps a bin in shared memory for the whole iteration
    struct Resident {
        bool ok = false;
        int RP = 0, QP = 0, CS = 0;   // rows per CTA, lanes per row of the row solve, CTAs per cluster
        int pitch = 0, reach = 0;     // shared-memory row pitch (doubles), CTAs a column carry reaches
        int nclusters = 0;            // clusters the device keeps resident at once
        size_t smem = 0;
    } res;
    int *d_resq = nullptr;            // bin queue counter of the resident solve
    uint8_t *d_rescode = nullptr;     // [ncd] geometry code of every cell
    double *d_reslut = nullptr;       // [256] code -> linked neighbours + boundary diagonals
    double *d_respax = nullptr, *d_respay = nullptr;   // chunk products of the row / column multipliers (double2 each)
};

struct Timer {
    double ms = 0.0;
    int64_t launches = 0;
};

struct qpb_ctx {
    qpb_config cfg{};
    int ncd = 0;  // dense cells ny*nx
    cudaStream_t stream = nullptr;      // stream every kernel of the context is enqueued on
    cudaStream_t own_stream = nullptr;  // the stream created with the context (stream == own_stream unless qpb_set_stream)
    // geometry
    bool have_geom = false;
    std::vector<uint8_t> h_flags;
    std::vector<double> h_bcx, h_bcy, h_src;
    std::vector<int32_t> h_cell2dense;
    uint8_t *d_flags = nullptr;      // [ncd]
    double *d_bcx = nullptr, *d_bcy = nullptr, *d_srcgeom = nullptr;  // [ncd]
    double *d_cx = nullptr, *d_cy = nullptr;  // [ncd] linked neighbours along x / y + boundary diagonal (0 outside the mask)
    int32_t *d_cell2dense = nullptr; // [ncell]
    bool thin_x = false, thin_y = false;  // no links along y / along x anywhere
    bool commuting = false;
    double gmax_x = 0.0, gmax_y = 0.0;    // Gershgorin bounds of -Lx, -Ly (grid units)
    double gmin = 0.0;                    // lower Gershgorin bound of both (negative only with a negative Robin beta)
    // diffusion
    bool have_D = false;
    std::vector<double> h_D;          // [ne] or [ne][ncell]
    double *d_Dcell = nullptr;        // variable D: dense [ne][ncd]
    DiffSlot slot[2];
    double *d_S = nullptr;            // QP state dense [ne][ncd]
    double *d_B = nullptr, *d_T1 = nullptr, *d_T2 = nullptr;  // CN work arrays
    unsigned long long *d_res = nullptr;   // [maxit][ne] residual max-norms (bit patterns of doubles)
    unsigned long long *d_unorm = nullptr; // [maxit][ne]
    int *d_done = nullptr;                 // [ne]
    int maxit = 0;
    double *d_kry[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // BiCGStab vectors, dense [ne][ncd], allocated on first use
    double *d_kry_small = nullptr;         // block partials + per-bin scalars of the same
    // collision
    bool have_coll = false;
    bool structured = false;          // Toeplitz/Hankel index maps, symmetric kernels
    double *d_Kr = nullptr, *d_Ks = nullptr, *d_KrT = nullptr, *d_KsT = nullptr;  // [ngap][ne][ne]
    double *d_rho = nullptr;          // [ngap][ne]
    int32_t *d_gapid = nullptr;       // [ncell] or null
    int32_t *d_idxd = nullptr, *d_idxs = nullptr, *d_idxdT = nullptr;  // [ne][ne]
    int8_t *d_sign = nullptr, *d_signT = nullptr;
    int32_t *d_dmap = nullptr;        // structured: diff index k -> phonon bin  [ne]
    int32_t *d_smap = nullptr;        // structured: sum index m -> phonon bin   [2ne-1]
    int32_t *d_kof = nullptr, *d_mof = nullptr;  // phonon bin -> k / m or -1   [nw]
    std::vector<int32_t> h_dmap, h_smap, h_kof, h_mof;  // host copies (ride in the kernel parameters)
    double *d_P = nullptr;            // phonon state [nw][ncell]
    bool uniform_ph = false;          // frozen phonons, identical in every cell: packed effective kernels in d_K4
    double *d_K4 = nullptr;           // [nep][nep][4] + rho[nep]
    std::vector<double> h_ph_bins;    // the per-bin occupations d_K4 / d_Mg were packed from (re-packed on a table upload)
    // tensor-core form of the same products (qpb_collide_gemm.cuh): 4 padded row-major matrices + rho, packed operands
    bool gemm_ready = false;
    double *d_Mg = nullptr, *d_Xn = nullptr, *d_Xp = nullptr;
    int gemm_nep = 0;
    long long gemm_npadc = 0;
    // several gap tables on the structured kernel: cells regrouped by table (qpbk_collision_setup)
    std::vector<int32_t> h_gapid;     // [ncell] host copy
    int32_t *d_cperm = nullptr;       // [ngroups * group_cc] cell index or -1
    int32_t *d_ggid = nullptr;        // [ngroups] gap table of the group
    int ngroups = 0, group_cc = 0;
    double *d_scratch = nullptr;      // collision scratch
    size_t scratch_bytes = 0;
    // fused layout exchange (qpb_set_exchange): peer diffusion states, routing of bins, dense index of my cells
    int x_nranks = 0;
    long long x_ncd = 0;
    double *x_peer[8] = {nullptr};
    int32_t *d_xdense = nullptr;      // [ncell]
    std::vector<int16_t> x_route;     // [ne] owner << 10 | row
    double *d_euler = nullptr;        // tables + operands of the fixed-bath Euler forms (qpbk_euler_step)
    size_t euler_bytes = 0;
    // generation array
    double *d_gen = nullptr;          // [ne][ncell]
    bool gen_resident = false;        // d_gen holds the array of the last QPB_GEN_ARRAY batch
    // device-side custom generation body (qpb_upload_generation_program): postfix program, its inputs, its verdict
    void *d_genprog = nullptr;        // qpb_gen_op[nops]
    int gen_nops = 0;
    double *d_gen_E = nullptr, *d_gen_x = nullptr, *d_gen_y = nullptr;   // [ne], [ncell], [ncell]
    int *d_gen_flag = nullptr;        // bit 0: a value was not finite, bit 1: a value was negative
    // reductions
    double *d_integrated = nullptr;   // [ncell]
    qpb_pauli_rec *d_pauli = nullptr; // [capacity]
    int pauli_cap = 0;
    void *d_pauli_part = nullptr;     // block partials
    int pauli_blocks = 0;
    // diagnostics
    qpb_diag diag{};
    bool timers_on = false;
    Timer timer[3];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // asynchronous snapshots (qpb_frames_snapshot / qpb_frames_download): device copy of the NaN-padded frames, the
    // event that marks it complete and the stream its download runs on while the context keeps stepping
    double *d_snap = nullptr;
    cudaEvent_t ev_snap = nullptr;
    cudaStream_t copy_stream = nullptr;
};

// ---- launch wrappers implemented in the .cu files (all enqueue on ctx->stream) ----
int qpbk_build_rhs(qpb_ctx *c, DiffSlot &s);
int qpbk_sweep_generic(qpb_ctx *c, DiffSlot &s, int dir, int iter, int mode);
int qpbk_diffuse(qpb_ctx *c, DiffSlot &s);
int qpbk_prepare_spectral(qpb_ctx *c, DiffSlot &s);
int qpbk_diffuse_spectral(qpb_ctx *c, DiffSlot &s);  // A u = b directly, b in d_B, result in d_S
int qpbk_diffuse_krylov(qpb_ctx *c, DiffSlot &s);   // A u = b, b in d_B, guess/result in d_S
void qpbk_free_krylov(qpb_ctx *c);
int qpbk_prepare_fast(qpb_ctx *c, DiffSlot &s);
int qpbr_plan(qpb_ctx *c, DiffSlot &s);                               // bin-resident solve: eligibility + launch shape
int qpbr_solve(qpb_ctx *c, DiffSlot &s, std::vector<int> &h_done);    // QPB_E_NOCONV (b left in d_B) when a bin stalls
int qpbk_sweep_fast(qpb_ctx *c, DiffSlot &s, int dir, int iter, int mode, bool check = true);
int qpbp_chunk(int n, int dir);
int qpbp_plan(qpb_ctx *c, DiffSlot &s, PipePlan &p);
int qpbp_sweep(qpb_ctx *c, DiffSlot &s, int dir, int iter, bool check);
void qpbk_free_slot(DiffSlot &s);

int qpbk_collide(qpb_ctx *c, double dt, int xmode = 0);
int qpbk_collision_setup(qpb_ctx *c);
int qpbk_euler_step(qpb_ctx *c, int kind, const double *K, const double *vec, double dt);
int qpbk_uniform_setup(qpb_ctx *c, const double *n_ph, bool per_bin = false);   // host phonon state [nw][ncell] ([nw] when per_bin) or null
int qpbk_broadcast_phonons(qpb_ctx *c, const double *d_bins);  // P[o][q] = bins[o]

int qpbk_add_generation(qpb_ctx *c, double scale, double rate, const double *d_array);
int qpbk_generation_program(qpb_ctx *c, double scale, double t, double *d_out);   // d_out: write g there instead of adding scale*g to the state
int qpbk_pauli(qpb_ctx *c, qpb_pauli_rec *d_out);
int qpbk_integrate(qpb_ctx *c);
int qpbk_scatter_state(qpb_ctx *c, const double *d_compact);  // [ne][ncell] -> dense
int qpbk_outer_state(qpb_ctx *c, const double *d_weights, const double *d_spatial);  // dense S[i][cell] = spatial[cell] * weights[i]
int qpbk_gather_state(qpb_ctx *c, double *d_compact);
int qpbk_frames(qpb_ctx *c, double *d_out);                   // dense NaN-padded frames [ne][ncd]

struct ScopedTimer {
    qpb_ctx *c;
    int which;
    bool on;
    ScopedTimer(qpb_ctx *ctx, int w) : c(ctx), which(w), on(ctx->timers_on) {
        if (on) cudaEventRecord(c->ev0, c->stream);
    }
    ~ScopedTimer() {
        if (on) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            c->timer[which].ms += ms;
            c->timer[which].launches += 1;
        }
    }
};
