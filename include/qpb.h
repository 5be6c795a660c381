/*
 * qpb.h — C ABI of the B200-native qpsim time-stepping hot path ("qpb" = quasiparticle-B200).
 *
 * The reference (Soren-O/Quasiparticle-Physics-Simulation, pure Python) has no FFI layer; its
 * boundary for this path is the Python function qpsim.solver.run_2d_crank_nicolson
 * (qpsim/solver.py:999-1036) and the in-place helpers apply_collision_step_fischer_catelani_uniform /
 * _nonuniform (qpsim/solver.py:794-875).  This header declares what a ctypes binding placed behind those
 * functions calls (see INTEGRATION.md for the binding).  Every entry point cites the reference lines whose
 * work it takes over.
 *
 * Conventions
 *  - plain C, caller-owned HOST buffers, library-owned DEVICE buffers, no exceptions across the boundary;
 *  - every function returns QPB_OK (0) or a negative QPB_E_* code; qpb_last_error() gives the message of the
 *    last failure on the calling thread;
 *  - host arrays use the reference's layouts: quasiparticle state  n[NE][N]  and phonon state  n_ph[Nw][N],
 *    C-contiguous float64, N = number of mask cells in np.argwhere(mask) (row-major) order
 *    (qpsim/solver.py:53-58); geometry arrays are dense [ny][nx];
 *  - there is NO CPU fallback: qpb_create fails when no sm_100 device is usable.
 */
#ifndef QPB_H
#define QPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPB_ABI_VERSION 1

/* status codes */
#define QPB_OK            0
#define QPB_E_INVALID    -1   /* bad argument / call order            */
#define QPB_E_CUDA       -2   /* CUDA runtime error (message has it)  */
#define QPB_E_NOMEM      -3
#define QPB_E_NODEVICE   -4   /* no usable CUDA device                */
#define QPB_E_NOCONV     -5   /* diffusion iteration did not converge */

/* qpb_config.flags */
#define QPB_F_DIFFUSION      (1u << 0)  /* enable_diffusion                      */
#define QPB_F_SCATTERING     (1u << 1)  /* enable_scattering                     */
#define QPB_F_RECOMBINATION  (1u << 2)  /* enable_recombination                  */
#define QPB_F_FREEZE_PHONONS (1u << 3)  /* freeze_phonon_dynamics (solver.py:1406) */
#define QPB_F_VARIABLE_D     (1u << 4)  /* per-cell D(E,x): solver.py:1145-1164   */
#define QPB_F_PAULI          (1u << 5)  /* record occupancy diagnostics each step (solver.py:967-996) */
#define QPB_F_SCALAR         (1u << 6)  /* legacy scalar mode, NE == 1, no dE factor (solver.py:1517-1571) */

typedef struct qpb_ctx qpb_ctx; /* opaque */

typedef struct qpb_config {
    int32_t abi_version;  /* QPB_ABI_VERSION */
    int32_t device;       /* CUDA ordinal */
    int32_t ny, nx;       /* mask shape */
    int32_t ne;           /* energy bins on this context (all bins on one GPU) */
    int32_t nw;           /* phonon bins (0 when collisions are off) */
    int32_t ncell;        /* N = mask cells */
    int32_t ngap;         /* number of distinct gap tables (1 = uniform gap) */
    uint32_t flags;       /* QPB_F_* */
    int32_t reserved;
    double dx;            /* mesh size */
    double dE;            /* energy bin width (1.0 for a single bin, solver.py:75-78) */
    double diff_tol;      /* relative max-norm residual tolerance of the CN solve (<=0: default 1e-12) */
    double pauli_floor;   /* pauli_density_floor of solver.py:1031 (forbidden-state test)          */
} qpb_config;

/* diagnostics of the most recent qpb_advance */
typedef struct qpb_diag {
    int64_t steps_done;
    int64_t sweeps;          /* directional tridiagonal sweeps executed (all bins count as one sweep) */
    int64_t bin_sweeps;      /* sum over bins of sweeps actually applied to that bin                    */
    int64_t pr_iterations;   /* Peaceman-Rachford double-sweeps (max over bins), summed over steps      */
    double  last_delta;      /* largest relative update of the last PR iteration                       */
    int32_t direct_mode;     /* 1: one-cell-thick geometry, CN solved by one direct sweep              */
    int32_t commuting;       /* 1: Lx,Ly commute (planned Wachspress sequence), 0: cyclic shifts        */
    int64_t kernel_launches; /* CUDA kernels launched by the library since creation                     */
    double  last_advance_ms; /* device time of the most recent qpb_advance step loop (CUDA events on the
                                library's stream, first launch to last kernel)                          */
    int32_t sweep_path;      /* kernels of the prepared CN solve: 0 generic (one thread per line), 1 chunked table
                                kernels, 2 persistent TMA-pipelined kernels (x and y), 3 the same with lines cut
                                into overlapping segments (lines longer than 512 cells), 4 direct spectral
                                solve (cosine transform along x + one tridiagonal solve along y per mode: full
                                rectangles with reflective left / right walls), 5 bin-resident solve (masks of
                                up to 256 x 256 cells: a thread-block cluster keeps a bin in shared memory for
                                the whole iteration, one launch per solve)                              */
    int32_t reserved;
} qpb_diag;

/* one record per time step, mirrors _pauli_occupancy_stats (solver.py:967-996) */
typedef struct qpb_pauli_rec {
    double  max_occ;     /* max n/rho over cells with rho > 1e-30 (0 elsewhere)          */
    int64_t max_index;   /* flat index i*N + cell of the first maximum (np.argmax order)  */
    int64_t forbidden;   /* flat index of the first forbidden-state cell, or -1           */
} qpb_pauli_rec;

const char *qpb_last_error(void);
int qpb_abi_version(void);
/* number of usable sm_100 devices, or a negative code */
int qpb_device_count(void);

int qpb_create(const qpb_config *cfg, qpb_ctx **out);
void qpb_destroy(qpb_ctx *ctx);

/*
 * Geometry and boundary terms — replaces build_laplacian_with_boundaries /_apply_boundary_contribution
 * (solver.py:112-212).  All arrays dense [ny*nx]:
 *   mask   1 = cell in the domain
 *   bcx    sum over the cell's boundary faces in x (left/right) of the face's diagonal term in 1/dx^2
 *          units (absorbing/dirichlet: 2, robin: beta*dx, reflective/neumann: 0); bcy likewise for up/down
 *   source s of  rhs = B u + dt*D*s  (dirichlet 2g/dx^2, neumann q/dx, robin gamma/dx), solver.py:130-148
 */
int qpb_upload_geometry(qpb_ctx *ctx, const uint8_t *mask, const double *bcx, const double *bcy,
                        const double *source);

/*
 * Diffusion coefficients — D[ne] (uniform gap, solver.py:1135,1167) or, with QPB_F_VARIABLE_D, D[ne][N]
 * (precomputed["D_array"], solver.py:1132, harmonic-mean faces solver.py:283).
 */
int qpb_upload_diffusion(qpb_ctx *ctx, const double *D);

/*
 * Prepare the Crank-Nicolson solve for a step length (slot 0 = dt, slot 1 = remainder_dt): replaces the
 * per-bin operator build + splu of solver.py:1143-1174.  Chooses direct / planned / cyclic mode and builds
 * the pivot tables of the tridiagonal sweeps.
 */
int qpb_prepare_diffusion(qpb_ctx *ctx, int slot, double dt);

/*
 * Collision tables (solver.py:1203-1238, 668-683).  K_r0/K_s0: [ngap][ne][ne] (NULL when that process is
 * off), rho: [ngap][ne], gap_id: [N] (NULL = all zero), idx_diff/idx_sum: [ne][ne] phonon-bin index maps,
 * sign: [ne][ne] int8 sign(E_i - E_j).
 */
int qpb_upload_collision(qpb_ctx *ctx, const double *K_r0, const double *K_s0, const double *rho,
                         const int32_t *gap_id, const int32_t *idx_diff, const int32_t *idx_sum,
                         const int8_t *sign);

/* state in the reference layouts; n_ph may be NULL when nw == 0 */
int qpb_set_state(qpb_ctx *ctx, const double *n, const double *n_ph);
int qpb_get_state(qpb_ctx *ctx, double *n, double *n_ph);
/* same as qpb_set_state for the reference's default initial phonon state n_ph_eq[:, None] * ones((1, N))
 * (solver.py:1183-1185): one occupation per phonon bin [Nw], broadcast over the cells on the device */
int qpb_set_state_uniform_phonons(qpb_ctx *ctx, const double *n, const double *n_ph_bins);
/* the reference's default initial quasiparticle state  state[i] = spatial_values * weights[i]  (solver.py:1281-1283)
 * formed on the device from its two factors (weights: [NE], spatial: [N]; one fp64 multiply per element, the same
 * rounding as the host product), together with the default phonon state of qpb_set_state_uniform_phonons
 * (n_ph_bins: [Nw], NULL when nw == 0): the call uploads NE + N + Nw doubles instead of NE*N */
int qpb_set_state_separable(qpb_ctx *ctx, const double *weights, const double *spatial, const double *n_ph_bins);
/* energy-integrated field  sum_i n[i][cell]*dE  (solver.py:1480), [N] */
int qpb_get_integrated(qpb_ctx *ctx, double *out);
/* the NE stored energy frames of one snapshot, dense [NE][ny][nx] with NaN outside the mask: what the reference
 * builds with NE calls of reconstruct_field (solver.py:215-218, 1484-1486), assembled on the device */
int qpb_get_frames(qpb_ctx *ctx, double *frames);
/* the same frames without stalling the run (the storage branch of the loop, solver.py:1479-1494, off the critical
 * path): qpb_frames_snapshot assembles them into a device buffer of their own, stream ordered and non-blocking;
 * qpb_frames_download waits for that snapshot only and copies it out on a second stream.  It may be called from
 * another host thread while this context keeps stepping; one snapshot is outstanding at a time. */
int qpb_frames_snapshot(qpb_ctx *ctx);
int qpb_frames_download(qpb_ctx *ctx, double *frames);

/* external generation (solver.py:878-964, 1459-1464) */
#define QPB_GEN_NONE     0
#define QPB_GEN_CONSTANT 1   /* rate                                  */
#define QPB_GEN_PULSE    2   /* rate while t0 <= t < t0 + duration    */
#define QPB_GEN_ARRAY    3   /* host-evaluated g[ne][N] (custom bodies, solver.py:918-962): uploaded by this call and
                              * applied in every step of the batch; it stays resident on the device afterwards */
#define QPB_GEN_RESIDENT 4   /* the array the last QPB_GEN_ARRAY call left on the device: a time-independent custom body
                              * is evaluated and uploaded once per run */

#define QPB_GEN_PROGRAM  5   /* the program of qpb_upload_generation_program, evaluated on the device at the time of
                              * every step: time-dependent custom bodies without a host evaluation and an NE x N
                              * upload per step */

typedef struct qpb_generation {
    int32_t mode;
    int32_t reserved;
    double  rate;
    double  pulse_start;
    double  pulse_duration;
    const double *array;   /* QPB_GEN_ARRAY only */
} qpb_generation;

/*
 * A custom generation body g(E, x, y, t, params) (evaluate_external_generation, solver.py:918-962) as a postfix program
 * over a stack of doubles, one value per (energy bin, cell).  The host layer translates the body's expression tree
 * (userexpr.compile_program); params are folded into constants.  Binary operators pop b then a and push a OP b.
 * cell_x / cell_y [ncell]: the normalised cell-centre coordinates the reference passes as x and y; E_bins [ne].
 * Values that are not finite or negative make qpb_advance fail with the reference's messages.
 */
typedef struct qpb_gen_op {
    int32_t op;       /* QPB_OP_* */
    int32_t reserved;
    double  value;    /* QPB_OP_CONST */
} qpb_gen_op;
enum {
    QPB_OP_CONST = 0, QPB_OP_E, QPB_OP_X, QPB_OP_Y, QPB_OP_T,
    QPB_OP_ADD, QPB_OP_SUB, QPB_OP_MUL, QPB_OP_DIV, QPB_OP_POW, QPB_OP_MOD, QPB_OP_FLOORDIV,
    QPB_OP_NEG, QPB_OP_NOT, QPB_OP_TRUTH,
    QPB_OP_LT, QPB_OP_LE, QPB_OP_GT, QPB_OP_GE, QPB_OP_EQ, QPB_OP_NE,
    QPB_OP_AND,       /* Python: a and b  ->  b if a else a */
    QPB_OP_OR,        /* Python: a or b   ->  a if a else b */
    QPB_OP_SELECT,    /* pops b, a, cond; pushes a where cond is true else b  (np.where, a if c else b) */
    QPB_OP_MIN, QPB_OP_MAX,         /* Python builtins on two values */
    QPB_OP_NPMIN, QPB_OP_NPMAX,     /* np.minimum / np.maximum: NaN propagates */
    QPB_OP_HEAVISIDE,
    QPB_OP_ABS, QPB_OP_SQRT, QPB_OP_EXP, QPB_OP_LOG, QPB_OP_LOG10, QPB_OP_SIN, QPB_OP_COS, QPB_OP_TAN,
    QPB_OP_ASIN, QPB_OP_ACOS, QPB_OP_ATAN, QPB_OP_SINH, QPB_OP_COSH, QPB_OP_TANH, QPB_OP_FLOOR, QPB_OP_CEIL,
    QPB_OP_TRUNC,
    QPB_OP_COUNT
};
#define QPB_GEN_MAX_OPS   512
#define QPB_GEN_MAX_STACK 24
int qpb_upload_generation_program(qpb_ctx *ctx, const qpb_gen_op *ops, int32_t nops, const double *E_bins,
                                  const double *cell_x, const double *cell_y);
/* g(E, x, y, t) of the uploaded program on the host, [ne][ncell] (tests; the run itself never downloads it) */
int qpb_eval_generation_program(qpb_ctx *ctx, double t, double *out);

/*
 * Advance nsteps time steps of length dt starting at time t_start — the loop body of
 * solver.py:1454-1478: generation, then C(dt/2) D(dt) C(dt/2) when both collisions and diffusion are on,
 * otherwise C(dt) D(dt); Pauli record per step when QPB_F_PAULI.  `slot` selects the prepared diffusion
 * operator.  pauli_out (may be NULL) receives nsteps records.
 */
int qpb_advance(qpb_ctx *ctx, int32_t nsteps, double dt, int32_t slot, double t_start,
                const qpb_generation *gen, qpb_pauli_rec *pauli_out);

/* single stages, used by the multi-GPU driver and by the in-place collision helpers (solver.py:794-875) */
int qpb_collide(qpb_ctx *ctx, double dt);
int qpb_diffuse(qpb_ctx *ctx, int32_t slot);
int qpb_pauli(qpb_ctx *ctx, qpb_pauli_rec *out);
/* the same record taken on the stream without a host synchronisation (slot-th record of a device-side list), and the
 * download of the first `count` records: the sharded driver checks the occupancy once per batch of steps, like
 * qpb_advance does on one GPU */
int qpb_pauli_record(qpb_ctx *ctx, int32_t slot);
int qpb_pauli_fetch(qpb_ctx *ctx, int32_t count, qpb_pauli_rec *out);

int qpb_get_diag(qpb_ctx *ctx, qpb_diag *out);
int qpb_synchronize(qpb_ctx *ctx);

/*
 * Device-side timing of the dominant kernels since the last qpb_reset_timers (CUDA events on the library's
 * stream): ms spent and launches, for bench.py's roofline.  which: 0 = x sweeps, 1 = y sweeps, 2 = collision.
 * Timers are off by default (they serialise); qpb_enable_timers(ctx, 1) turns them on.
 */
int qpb_enable_timers(qpb_ctx *ctx, int on);
int qpb_reset_timers(qpb_ctx *ctx);
int qpb_get_timer(qpb_ctx *ctx, int which, double *ms, int64_t *launches);

/*
 * Roofline denominators measured with the library's own micro-kernels on `device`:
 *   qpb_measure_fp64  dependent-chain-free DFMA loop on every SM -> TFLOP/s (FMA = 2 flop)
 *   qpb_measure_copy  float64 copy of `bytes` (read + write counted) -> GB/s
 */
int qpb_measure_fp64(int device, double *tflops);
int qpb_measure_copy(int device, int64_t bytes, double *gbs);

/* raw device pointers for NCCL plumbing (torch wraps them): which 0 = dense QP state [ne][ny*nx],
 * 1 = phonon state [nw][N]; *bytes receives the allocation size */
int qpb_device_ptr(qpb_ctx *ctx, int which, void **ptr, int64_t *bytes);

/*
 * Multi-GPU plumbing: diffusion is sharded by energy bin, collisions by cell.  A rank
 * holds a diffusion context (its bins, all cells) and a collision context (all bins, its cells as a 1 x N_local
 * strip); between the stages the host moves blocks with NCCL and these two calls convert between a received
 * block  d_block[ne][count]  (cells cell0 .. cell0+count-1 of the compressed ordering, DEVICE memory) and the
 * dense state of the context.  qpb_add_generation applies  state += scale * rate  (solver.py:1464) as a stage.
 */
int qpb_scatter_block(qpb_ctx *ctx, const double *d_block, int32_t cell0, int32_t count);
int qpb_gather_block(qpb_ctx *ctx, double *d_block, int32_t cell0, int32_t count);
int qpb_add_generation(qpb_ctx *ctx, double scale, double rate);
/* The custom forms of the same stage (evaluate_external_generation, solver.py:918-962) for a context that holds a slice
 * of the cells:  state += scale * g  with g the uploaded program evaluated at time t (qpb_upload_generation_program with
 * the slice's coordinates), or a host array g[ne][ncell] (uploaded by the call and kept; NULL = the array of the last
 * call: a time-independent body).  Stream ordered.  qpb_generation_status waits for the stream and returns and clears
 * the program's verdict: bit 0 = a value was not finite, bit 1 = a value was negative. */
int qpb_add_generation_program(qpb_ctx *ctx, double scale, double t);
int qpb_add_generation_array(qpb_ctx *ctx, double scale, const double *array);
int qpb_generation_status(qpb_ctx *ctx, int32_t *flags);
/* Enqueue all further work of the context on the caller's CUDA stream (a cudaStream_t; NULL = back to the stream
 * the library created).  The host driver passes the stream its NCCL calls are ordered against, so stages and
 * exchanges need no host synchronisation between them.  scatter/gather_block and add_generation are stream
 * ordered (no host sync); every call that returns data to the host synchronises that stream. */
int qpb_set_stream(qpb_ctx *ctx, void *cuda_stream);

/* Fixed-bath forward-Euler collision forms of the reference (apply_scattering_step, solver.py:551-581;
 * apply_recombination_step, solver.py:584-605) on the context's state: the K(E,E') contraction and the n.R.n form as
 * one FP64 tensor-core GEMM from 64 bins on, a fused GEMV below.
 *   kind 1: K = K_s [NE][NE], vec = rho_bins [NE];   kind 2: K = K_r [NE][NE], vec = G_therm [NE] */
int qpb_euler_step(qpb_ctx *ctx, int32_t kind, const double *K, const double *vec, double dt);

/* ---- layout exchange fused into the collision kernel (multi-GPU) -------------------------------------
 * The cell-sharded collision context of a rank can store the updated n(E) of its cells straight into the bin-sharded
 * diffusion states of all ranks (mode 1: collide, then scatter over NVLink peer memory) and read its cells from there
 * (mode 2: gather, then collide), which replaces the all-to-all between the two layouts and its packing kernels.
 *   qpb_set_exchange    peer_state[r] = device address (in THIS process) of rank r's diffusion state [rows][peer_ncd];
 *                       bin_owner/bin_row[NE] route bin i to (rank, row); cell_dense[N] = dense grid index of my cells
 *   qpb_ipc_export/open make another process's state array addressable (cudaIpc*); handles are 64 bytes
 * The caller orders the phases across ranks (a barrier between writers and readers of a diffusion state). */
int qpb_set_exchange(qpb_ctx *ctx, int32_t nranks, void *const *peer_state, int64_t peer_ncd, const int16_t *bin_owner,
                     const int16_t *bin_row, const int32_t *cell_dense);
int qpb_collide_exchange(qpb_ctx *ctx, double dt, int32_t mode);
int qpb_ipc_export(qpb_ctx *ctx, int which, void *handle64);
int qpb_ipc_open(int device, const void *handle64, void **ptr);
int qpb_ipc_close(int device, void *ptr);

/* Device buffers of destroyed contexts are parked in a bounded per-process cache (QPB_CACHE_MB, default 4096) so
 * that back-to-back runs of the same shape do not pay cudaMalloc/cudaFree again; this returns them to the driver. */
int qpb_trim_cache(void);

#ifdef __cplusplus
}
#endif
#endif /* QPB_H */
