// Layout conversion, external generation, Pauli diagnostics and energy integration.
//   generation  solver.py:1459-1464        Pauli stats  solver.py:967-996        integration  solver.py:1480
#include "qpb_internal.h"

#include <algorithm>
#include <cfloat>
#include <climits>

namespace {

__global__ void k_scatter_state(int ne, int ncell, int ncd, const double *__restrict__ compact,
                                double *__restrict__ dense, const int32_t *__restrict__ c2d) {
    const long long total = (long long)ne * ncell;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(g / ncell);
        const int q = (int)(g - (long long)i * ncell);
        dense[(long long)i * ncd + c2d[q]] = compact[g];
    }
}

// state[i] = spatial_values * weights[i]  (solver.py:1281-1283), written straight into the dense layout
__global__ void k_outer_state(int ne, int ncell, int ncd, const double *__restrict__ weights,
                              const double *__restrict__ spatial, double *__restrict__ dense,
                              const int32_t *__restrict__ c2d) {
    const long long total = (long long)ne * ncell;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(g / ncell);
        const int q = (int)(g - (long long)i * ncell);
        dense[(long long)i * ncd + c2d[q]] = spatial[q] * weights[i];
    }
}

__global__ void k_gather_state(int ne, int ncell, int ncd, double *__restrict__ compact,
                               const double *__restrict__ dense, const int32_t *__restrict__ c2d) {
    const long long total = (long long)ne * ncell;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(g / ncell);
        const int q = (int)(g - (long long)i * ncell);
        compact[g] = dense[(long long)i * ncd + c2d[q]];
    }
}

// frames[i][p] = cell p in the mask ? S[i][p] : NaN   (reconstruct_field for every bin, solver.py:215-218)
__global__ void k_frames(long long total, int ncd, const double *__restrict__ S, const uint8_t *__restrict__ flags,
                         double *__restrict__ out) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(g % ncd);
        out[g] = (flags[p] & QPB_IN) ? S[g] : nan;
    }
}

// state += scale * g   (g: one rate for every bin and cell, or a host-evaluated array [ne][ncell])
// grid.y = energy bin: no per-element index division, the dense index of a cell is loaded once per thread and bin
__global__ void k_add_generation(int ne, int ncell, int ncd, double *__restrict__ S,
                                 const int32_t *__restrict__ c2d, double scale, double rate,
                                 const double *__restrict__ arr) {
    const int i = blockIdx.y;
    double *row = S + (long long)i * ncd;
    const double *g = arr ? arr + (long long)i * ncell : nullptr;
    const double add = scale * rate;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += gridDim.x * blockDim.x)
        row[c2d[q]] += g ? scale * g[q] : add;
}

// state += scale * g(E, x, y, t) with g given as a postfix program (qpb.h: qpb_gen_op) - a custom generation body evaluated
// where the state lives (evaluate_external_generation, solver.py:918-962, 1459-1464).  One thread per (bin, cell); every
// thread of the grid runs the same operator sequence (no divergence), the program sits in shared memory.  Python's
// scalar semantics (the reference falls back to a per-value loop whenever a body is not vectorisable).
__global__ void k_generation_program(int ne, int ncell, int ncd, double *__restrict__ S, const int32_t *__restrict__ c2d,
                                     double scale, double t, int nops, const qpb_gen_op *__restrict__ prog,
                                     const double *__restrict__ Eb, const double *__restrict__ cx,
                                     const double *__restrict__ cy, double *__restrict__ out, int *__restrict__ flag) {
    __shared__ qpb_gen_op ops[QPB_GEN_MAX_OPS];
    for (int k = threadIdx.x; k < nops; k += blockDim.x) ops[k] = prog[k];
    __syncthreads();
    const int i = blockIdx.y;
    const double E = Eb[i];
    int bad = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += gridDim.x * blockDim.x) {
        const double x = cx[q], y = cy[q];
        double st[QPB_GEN_MAX_STACK];
        int sp = 0;   // number of values on the stack
        for (int k = 0; k < nops; ++k) {
            const int op = ops[k].op;
            if (op <= QPB_OP_T) {
                st[sp++] = op == QPB_OP_CONST ? ops[k].value : op == QPB_OP_E ? E : op == QPB_OP_X ? x : op == QPB_OP_Y ? y : t;
            } else if (op == QPB_OP_SELECT) {
                const double b = st[--sp], a = st[--sp], c = st[sp - 1];
                st[sp - 1] = c != 0.0 ? a : b;
            } else if (op >= QPB_OP_ABS || op == QPB_OP_NEG || op == QPB_OP_NOT || op == QPB_OP_TRUTH) {
                const double a = st[sp - 1];
                double r;
                switch (op) {
                    case QPB_OP_NEG: r = -a; break;
                    case QPB_OP_NOT: r = a == 0.0 ? 1.0 : 0.0; break;
                    case QPB_OP_TRUTH: r = a != 0.0 ? 1.0 : 0.0; break;
                    case QPB_OP_ABS: r = fabs(a); break;
                    case QPB_OP_SQRT: r = sqrt(a); break;
                    case QPB_OP_EXP: r = exp(a); break;
                    case QPB_OP_LOG: r = log(a); break;
                    case QPB_OP_LOG10: r = log10(a); break;
                    case QPB_OP_SIN: r = sin(a); break;
                    case QPB_OP_COS: r = cos(a); break;
                    case QPB_OP_TAN: r = tan(a); break;
                    case QPB_OP_ASIN: r = asin(a); break;
                    case QPB_OP_ACOS: r = acos(a); break;
                    case QPB_OP_ATAN: r = atan(a); break;
                    case QPB_OP_SINH: r = sinh(a); break;
                    case QPB_OP_COSH: r = cosh(a); break;
                    case QPB_OP_TANH: r = tanh(a); break;
                    case QPB_OP_FLOOR: r = floor(a); break;
                    case QPB_OP_CEIL: r = ceil(a); break;
                    default: r = trunc(a); break;
                }
                st[sp - 1] = r;
            } else {
                const double b = st[--sp], a = st[sp - 1];
                double r;
                switch (op) {
                    case QPB_OP_ADD: r = a + b; break;
                    case QPB_OP_SUB: r = a - b; break;
                    case QPB_OP_MUL: r = a * b; break;
                    case QPB_OP_DIV: r = a / b; break;
                    case QPB_OP_POW: r = pow(a, b); break;
                    case QPB_OP_MOD: {   // Python / numpy: the result takes the sign of the divisor
                        r = fmod(a, b);
                        if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
                        break;
                    }
                    case QPB_OP_FLOORDIV: r = floor(a / b); break;
                    case QPB_OP_LT: r = a < b ? 1.0 : 0.0; break;
                    case QPB_OP_LE: r = a <= b ? 1.0 : 0.0; break;
                    case QPB_OP_GT: r = a > b ? 1.0 : 0.0; break;
                    case QPB_OP_GE: r = a >= b ? 1.0 : 0.0; break;
                    case QPB_OP_EQ: r = a == b ? 1.0 : 0.0; break;
                    case QPB_OP_NE: r = a != b ? 1.0 : 0.0; break;
                    case QPB_OP_AND: r = a != 0.0 ? b : a; break;
                    case QPB_OP_OR: r = a != 0.0 ? a : b; break;
                    case QPB_OP_MIN: r = b < a ? b : a; break;     // min(a, b): the first smallest
                    case QPB_OP_MAX: r = b > a ? b : a; break;
                    case QPB_OP_NPMIN: r = (a != a || b != b) ? a + b : (b < a ? b : a); break;
                    case QPB_OP_NPMAX: r = (a != a || b != b) ? a + b : (b > a ? b : a); break;
                    default: r = a != a ? a : (a < 0.0 ? 0.0 : (a == 0.0 ? b : 1.0)); break;   // heaviside(a, b)
                }
                st[sp - 1] = r;
            }
        }
        const double g = st[0];
        if (!(fabs(g) <= 1.79769313486231570e308)) bad |= 1;
        else if (g < 0.0) bad |= 2;
        if (out) out[(long long)i * ncell + q] = g;
        else S[(long long)i * ncd + c2d[q]] += scale * g;
    }
    if (bad) atomicOr(flag, bad);
}

// integrated[q] = (sum_i n[i][q]) * dE, bins added in order like np.sum(state, axis=0)
__global__ void k_integrate(int ne, int ncell, int ncd, const double *__restrict__ S,
                            const int32_t *__restrict__ c2d, double dE, double *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ncell) return;
    const int d = c2d[q];
    double acc = 0.0;
    for (int i = 0; i < ne; ++i) acc += S[(long long)i * ncd + d];
    out[q] = acc * dE;
}

struct PauliPart {
    double val;
    long long idx;
    long long forb;
};

__device__ __forceinline__ void pauli_merge(PauliPart &a, const PauliPart &b) {
    if (b.val > a.val || (b.val == a.val && b.idx < a.idx)) {
        a.val = b.val;
        a.idx = b.idx;
    }
    if (b.forb < a.forb) a.forb = b.forb;
}

__device__ void pauli_block_reduce(PauliPart &p) {
    __shared__ PauliPart sh[32];
    for (int o = 16; o > 0; o >>= 1) {
        PauliPart q;
        q.val = __shfl_down_sync(0xffffffffu, p.val, o);
        q.idx = __shfl_down_sync(0xffffffffu, p.idx, o);
        q.forb = __shfl_down_sync(0xffffffffu, p.forb, o);
        pauli_merge(p, q);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = p;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        PauliPart q = l < nw ? sh[l] : PauliPart{-DBL_MAX, LLONG_MAX, LLONG_MAX};
        for (int o = 16; o > 0; o >>= 1) {
            PauliPart r;
            r.val = __shfl_down_sync(0xffffffffu, q.val, o);
            r.idx = __shfl_down_sync(0xffffffffu, q.idx, o);
            r.forb = __shfl_down_sync(0xffffffffu, q.forb, o);
            pauli_merge(q, r);
        }
        p = q;
    }
}

// grid.y = energy bin, four cells per thread and pass (independent loads); flattened index g = bin*ncell + cell grows
// along a thread's loop, so ">" keeps the first maximum of the np.argmax order (solver.py:967-996)
__global__ void k_pauli_stage1(int ne, int ncell, int ncd, const double *__restrict__ S,
                               const int32_t *__restrict__ c2d, const double *__restrict__ rho,
                               const int32_t *__restrict__ gapid, double floor_, PauliPart *__restrict__ part) {
    PauliPart p{-DBL_MAX, LLONG_MAX, LLONG_MAX};
    const int i = blockIdx.y;
    const double *row = S + (long long)i * ncd;
    const long long g0 = (long long)i * ncell;
    const double r0 = rho[i];
    constexpr int U = 4;
    const int stride = gridDim.x * blockDim.x;
    for (int qb = blockIdx.x * blockDim.x + threadIdx.x; qb < ncell; qb += U * stride) {
        double n[U], r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = qb + u * stride;
            n[u] = q < ncell ? row[c2d[q]] : 0.0;
            r[u] = (gapid && q < ncell) ? rho[gapid[q] * ne + i] : r0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = qb + u * stride;
            if (q >= ncell) break;
            const bool ok = r[u] > 1e-30;
            const double f = ok ? n[u] / fmax(r[u], 1e-30) : 0.0;
            if (f > p.val) {
                p.val = f;
                p.idx = g0 + q;
            }
            if (!ok && n[u] > floor_ && g0 + q < p.forb) p.forb = g0 + q;
        }
    }
    pauli_block_reduce(p);
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = p;
}

__global__ void k_pauli_stage2(int nparts, const PauliPart *__restrict__ part, qpb_pauli_rec *__restrict__ out) {
    PauliPart p{-DBL_MAX, LLONG_MAX, LLONG_MAX};
    for (int k = threadIdx.x; k < nparts; k += blockDim.x) pauli_merge(p, part[k]);
    pauli_block_reduce(p);
    if (threadIdx.x == 0) {
        out->max_occ = p.val;
        out->max_index = p.idx;
        out->forbidden = p.forb == LLONG_MAX ? -1 : p.forb;
    }
}

inline int grid_for(long long total, int threads, int cap) {
    return (int)std::max<long long>(1, std::min<long long>(ceil_div64(total, threads), cap));
}

}  // namespace

int qpbk_scatter_state(qpb_ctx *c, const double *d_compact) {
    const auto &cf = c->cfg;
    const long long total = (long long)cf.ne * cf.ncell;
    k_scatter_state<<<grid_for(total, 256, 148 * 16), 256, 0, c->stream>>>(cf.ne, cf.ncell, c->ncd, d_compact, c->d_S,
                                                                          c->d_cell2dense);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_outer_state(qpb_ctx *c, const double *d_weights, const double *d_spatial) {
    const auto &cf = c->cfg;
    const long long total = (long long)cf.ne * cf.ncell;
    k_outer_state<<<grid_for(total, 256, 148 * 16), 256, 0, c->stream>>>(cf.ne, cf.ncell, c->ncd, d_weights, d_spatial,
                                                                        c->d_S, c->d_cell2dense);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_gather_state(qpb_ctx *c, double *d_compact) {
    const auto &cf = c->cfg;
    const long long total = (long long)cf.ne * cf.ncell;
    k_gather_state<<<grid_for(total, 256, 148 * 16), 256, 0, c->stream>>>(cf.ne, cf.ncell, c->ncd, d_compact, c->d_S,
                                                                         c->d_cell2dense);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

__global__ void k_broadcast_rows(long long total, int ncell, const double *__restrict__ bins, double *__restrict__ P) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x)
        P[g] = bins[g / ncell];
}

int qpbk_broadcast_phonons(qpb_ctx *c, const double *d_bins) {
    const long long total = (long long)c->cfg.nw * c->cfg.ncell;
    k_broadcast_rows<<<grid_for(total, 256, 148 * 16), 256, 0, c->stream>>>(total, c->cfg.ncell, d_bins, c->d_P);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_frames(qpb_ctx *c, double *d_out) {
    const long long total = (long long)c->cfg.ne * c->ncd;
    k_frames<<<grid_for(total, 256, 148 * 16), 256, 0, c->stream>>>(total, c->ncd, c->d_S, c->d_flags, d_out);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_add_generation(qpb_ctx *c, double scale, double rate, const double *d_array) {
    const auto &cf = c->cfg;
    const dim3 grid((unsigned)std::max(1, std::min((cf.ncell + 255) / 256, 64)), (unsigned)cf.ne);
    k_add_generation<<<grid, 256, 0, c->stream>>>(cf.ne, cf.ncell, c->ncd, c->d_S, c->d_cell2dense, scale, rate,
                                                  d_array);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_generation_program(qpb_ctx *c, double scale, double t, double *d_out) {
    const auto &cf = c->cfg;
    const dim3 grid((unsigned)std::max(1, std::min((cf.ncell + 255) / 256, 64)), (unsigned)cf.ne);
    k_generation_program<<<grid, 256, 0, c->stream>>>(cf.ne, cf.ncell, c->ncd, c->d_S, c->d_cell2dense, scale, t, c->gen_nops,
                                                      (const qpb_gen_op *)c->d_genprog, c->d_gen_E, c->d_gen_x, c->d_gen_y,
                                                      d_out, c->d_gen_flag);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_integrate(qpb_ctx *c) {
    const auto &cf = c->cfg;
    k_integrate<<<(cf.ncell + 127) / 128, 128, 0, c->stream>>>(cf.ne, cf.ncell, c->ncd, c->d_S, c->d_cell2dense, cf.dE,
                                                               c->d_integrated);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;
}

int qpbk_pauli(qpb_ctx *c, qpb_pauli_rec *d_out) {
    const auto &cf = c->cfg;
    // grid.x blocks of 256 threads x 4 cells per bin, about 8 CTAs per SM in total
    const int gx = std::max(1, std::min((cf.ncell + 1023) / 1024, std::max(1, 148 * 8 / cf.ne)));
    const int blocks = gx * cf.ne;
    if (!c->d_pauli_part || c->pauli_blocks < blocks) {
        if (c->d_pauli_part) qpb_dev_free(c->d_pauli_part);
        c->d_pauli_part = nullptr;
        c->pauli_blocks = blocks;
        QPB_CUDA(qpb_dev_malloc(&c->d_pauli_part, sizeof(PauliPart) * c->pauli_blocks));
    }
    k_pauli_stage1<<<dim3((unsigned)gx, (unsigned)cf.ne), 256, 0, c->stream>>>(
        cf.ne, cf.ncell, c->ncd, c->d_S, c->d_cell2dense, c->d_rho, c->d_gapid, cf.pauli_floor,
        (PauliPart *)c->d_pauli_part);
    QPB_CHECK_LAUNCH();
    k_pauli_stage2<<<1, 256, 0, c->stream>>>(blocks, (const PauliPart *)c->d_pauli_part, d_out);
    QPB_CHECK_LAUNCH();
    c->diag.kernel_launches += 2;
    return QPB_OK;
}

// ---- block <-> dense conversion for the energy-sharded / cell-sharded exchange -----------------------------
namespace {
__global__ void k_block(int ne, int count, int ncd, int cell0, double *__restrict__ block, double *__restrict__ dense,
                        const int32_t *__restrict__ c2d, int to_dense) {
    const long long total = (long long)ne * count;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(g / count);
        const int q = (int)(g - (long long)i * count);
        const long long d = (long long)i * ncd + c2d[cell0 + q];
        if (to_dense) dense[d] = block[g];
        else block[g] = dense[d];
    }
}
}  // namespace

static int block_call(qpb_ctx *c, double *d_block, int cell0, int count, int to_dense) {
    if (!c || !d_block || cell0 < 0 || count < 0 || cell0 + count > c->cfg.ncell || !c->have_geom) {
        qpb_set_error("qpb_scatter/gather_block: bad arguments (cell0=%d count=%d)", cell0, count);
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaSetDevice(c->cfg.device));
    if (count == 0) return QPB_OK;
    const long long total = (long long)c->cfg.ne * count;
    k_block<<<grid_for(total, 256, 148 * 16), 256, 0, c->stream>>>(c->cfg.ne, count, c->ncd, cell0, d_block, c->d_S,
                                                                  c->d_cell2dense, to_dense);
    c->diag.kernel_launches++;
    QPB_CHECK_LAUNCH();
    return QPB_OK;   // stream ordered: the caller synchronises (qpb_synchronize) or enqueues on the same stream
}

extern "C" int qpb_scatter_block(qpb_ctx *c, const double *d_block, int32_t cell0, int32_t count) {
    return block_call(c, const_cast<double *>(d_block), cell0, count, 1);
}

extern "C" int qpb_gather_block(qpb_ctx *c, double *d_block, int32_t cell0, int32_t count) {
    return block_call(c, d_block, cell0, count, 0);
}

extern "C" int qpb_add_generation(qpb_ctx *c, double scale, double rate) {
    if (!c || !c->have_geom) {
        qpb_set_error("qpb_add_generation: context without geometry");
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaSetDevice(c->cfg.device));
    return qpbk_add_generation(c, scale, rate, nullptr);
}

extern "C" int qpb_add_generation_program(qpb_ctx *c, double scale, double t) {
    if (!c || !c->have_geom || !c->d_genprog || c->gen_nops <= 0) {
        qpb_set_error("qpb_add_generation_program: context without geometry or without an uploaded program");
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaSetDevice(c->cfg.device));
    return qpbk_generation_program(c, scale, t, nullptr);
}

extern "C" int qpb_add_generation_array(qpb_ctx *c, double scale, const double *array) {
    if (!c || !c->have_geom) {
        qpb_set_error("qpb_add_generation_array: context without geometry");
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaSetDevice(c->cfg.device));
    const size_t n = (size_t)c->cfg.ne * c->cfg.ncell;
    if (array) {
        if (!c->d_gen) QPB_CUDA(qpb_dev_malloc((void **)&c->d_gen, sizeof(double) * n));
        c->gen_resident = false;
        QPB_CUDA(cudaMemcpyAsync(c->d_gen, array, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));   // the caller's buffer is free again
        c->gen_resident = true;
    } else if (!c->d_gen || !c->gen_resident) {
        qpb_set_error("qpb_add_generation_array: no generation array is resident");
        return QPB_E_INVALID;
    }
    return qpbk_add_generation(c, scale, 0.0, c->d_gen);
}

extern "C" int qpb_generation_status(qpb_ctx *c, int32_t *flags) {
    if (!c || !flags) {
        qpb_set_error("qpb_generation_status: null argument");
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaSetDevice(c->cfg.device));
    *flags = 0;
    if (!c->d_gen_flag) return QPB_OK;
    int v = 0;
    QPB_CUDA(cudaMemcpyAsync(&v, c->d_gen_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaMemsetAsync(c->d_gen_flag, 0, sizeof(int), c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    *flags = v;
    return QPB_OK;
}

// ---- roofline denominators ------------------------------------------------------------------------------
namespace {
__global__ void k_fp64_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_copy(const double2 *__restrict__ in, double2 *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = in[i];
}
}  // namespace

extern "C" int qpb_measure_fp64(int device, double *tflops) {
    if (!tflops) return QPB_E_INVALID;
    QPB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    QPB_CUDA(cudaGetDeviceProperties(&p, device));
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *d = nullptr;
    QPB_CUDA(cudaMalloc((void **)&d, sizeof(double) * blocks * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        k_fp64_peak<<<blocks, threads>>>(d, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep >= 1 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    QPB_CHECK_LAUNCH();
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return QPB_OK;
}

extern "C" int qpb_measure_copy(int device, int64_t bytes, double *gbs) {
    if (!gbs || bytes < 1024) return QPB_E_INVALID;
    QPB_CUDA(cudaSetDevice(device));
    const long long n = bytes / 16;
    double2 *x = nullptr, *y = nullptr;
    QPB_CUDA(cudaMalloc((void **)&x, n * 16));
    QPB_CUDA(cudaMalloc((void **)&y, n * 16));
    QPB_CUDA(cudaMemset(x, 0, n * 16));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(a);
        k_copy<<<148 * 16, 512>>>(x, y, n);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(x);
    cudaFree(y);
    QPB_CHECK_LAUNCH();
    *gbs = 2.0 * n * 16 / (best * 1e-3) / 1e9;
    return QPB_OK;
}
