// C ABI of libqpb: context, uploads, the time-step loop.  See include/qpb.h for the contract.
#include "qpb_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <unordered_map>

static thread_local std::string g_err;

void qpb_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

extern "C" const char *qpb_last_error(void) { return g_err.c_str(); }
extern "C" int qpb_abi_version(void) { return QPB_ABI_VERSION; }

extern "C" int qpb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        qpb_set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return QPB_E_NODEVICE;
    }
    int usable = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) usable++;
    }
    return usable;
}

// ---- device block cache ------------------------------------------------------------------------------------
// cudaMalloc / cudaFree of the state-sized arrays cost ~100 ms each way per run on the reference-facing path (one
// context per run_2d_crank_nicolson call).  Freed blocks are parked here (exact-size reuse, per device) so that the
// next run with the same shapes pays nothing; the cache is bounded (QPB_CACHE_MB, default 4096; 0 disables it),
// evicts oldest first, and is emptied on an allocation failure before the allocation is retried.
namespace {
struct CachedBlock {
    int device;
    size_t bytes;
    void *ptr;
};
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;                       // oldest first
std::unordered_map<void *, size_t> g_live;              // bytes of every block handed out
size_t g_cache_bytes = 0;

size_t cache_cap() {
    static size_t cap = [] {
        const char *e = getenv("QPB_CACHE_MB");
        const long long mb = e ? atoll(e) : 4096;
        return (size_t)(mb < 0 ? 0 : mb) << 20;
    }();
    return cap;
}

void cache_release_all_locked() {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &b : g_cache) {
        cudaSetDevice(b.device);
        cudaFree(b.ptr);
    }
    g_cache.clear();
    g_cache_bytes = 0;
    cudaSetDevice(cur);
}
}  // namespace

cudaError_t qpb_dev_malloc(void **p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = g_cache.size(); i-- > 0;)
        if (g_cache[i].device == dev && g_cache[i].bytes == bytes) {
            *p = g_cache[i].ptr;
            g_cache_bytes -= bytes;
            g_cache.erase(g_cache.begin() + i);
            g_live[*p] = bytes;
            // a block from the driver arrives zeroed; keep that property for recycled ones
            if (getenv("QPB_DEBUG_ALLOC")) {
                auto t0 = std::chrono::steady_clock::now();
                cudaMemsetAsync(*p, 0, bytes, 0);
                cudaError_t e2 = cudaStreamSynchronize(0);
                const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                if (ms > 1.0) fprintf(stderr, "[qpb] recycle %zu bytes: memset+sync %.2f ms\n", bytes, ms);
                return e2;
            }
            cudaMemsetAsync(*p, 0, bytes, 0);
            return cudaStreamSynchronize(0);
        }
    if (getenv("QPB_DEBUG_ALLOC")) fprintf(stderr, "[qpb] cache miss %zu bytes (cached %zu in %zu blocks)\n", bytes, g_cache_bytes, g_cache.size());
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && !g_cache.empty()) {
        cudaGetLastError();
        cache_release_all_locked();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) g_live[*p] = bytes;
    return e;
}

void qpb_dev_free(void *p) {
    if (!p) return;
    // what cudaFree guarantees implicitly: nothing in flight still uses the block when somebody else gets it
    cudaDeviceSynchronize();
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    size_t bytes = 0;
    auto it = g_live.find(p);
    if (it != g_live.end()) {
        bytes = it->second;
        g_live.erase(it);
    }
    const size_t cap = cache_cap();
    if (bytes == 0 || bytes > cap) {
        cudaFree(p);
        return;
    }
    while (g_cache_bytes + bytes > cap && !g_cache.empty()) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(g_cache.front().device);
        cudaFree(g_cache.front().ptr);
        cudaSetDevice(cur);
        g_cache_bytes -= g_cache.front().bytes;
        g_cache.erase(g_cache.begin());
    }
    g_cache.push_back({dev, bytes, p});
    g_cache_bytes += bytes;
}

extern "C" int qpb_trim_cache(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_release_all_locked();
    return QPB_OK;
}

template <class T>
static int dev_alloc(T **p, size_t count) {
    *p = nullptr;
    if (count == 0) return QPB_OK;
    cudaError_t e = qpb_dev_malloc((void **)p, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        qpb_set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return QPB_E_NOMEM;
    }
    return QPB_OK;
}

#define QPB_ALLOC(ptr, count)                              \
    do {                                                   \
        int _rc = dev_alloc(&(ptr), (size_t)(count));      \
        if (_rc != QPB_OK) return _rc;                     \
    } while (0)

template <class T>
static void dev_free(T *&p) {
    if (p) qpb_dev_free(p);
    p = nullptr;
}

void qpbk_free_slot(DiffSlot &s) {
    dev_free(s.d_a);
    dev_free(s.d_shift);
    dev_free(s.d_jlen);
    dev_free(s.d_tol);
    dev_free(s.d_tolk);
    dev_free(s.d_kshift);
    dev_free(s.d_sp_tw);
    dev_free(s.d_sp_tw2);
    dev_free(s.d_sp_lam);
    dev_free(s.d_sp_bcy);
    dev_free(s.d_sp_piv);
    dev_free(s.d_sp_srchat);
    s.sp_fused = false;
    s.sp_T = 0;
    s.spectral = false;
    dev_free(s.d_known);
    dev_free(s.d_ex);
    dev_free(s.d_ey);
    dev_free(s.d_gbx);
    dev_free(s.d_gby);
    dev_free(s.d_src);
    dev_free(s.fx.d_cls);
    dev_free(s.fx.d_tab);
    dev_free(s.fx.d_tabg);
    dev_free(s.fy.d_tabg);
    dev_free(s.fy.d_cls);
    dev_free(s.fy.d_tab);
    dev_free(s.d_resq);
    dev_free(s.d_rescode);
    dev_free(s.d_reslut);
    dev_free(s.d_respax);
    dev_free(s.d_respay);
    s.res = DiffSlot::Resident();
    s.ready = false;
    s.fast = false;
    s.pipe = PipePlan();
}

extern "C" int qpb_create(const qpb_config *cfg, qpb_ctx **out) {
    if (!cfg || !out) {
        qpb_set_error("qpb_create: null argument");
        return QPB_E_INVALID;
    }
    *out = nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (getenv("QPB_DEBUG_ALLOC"))
            fprintf(stderr, "[qpb] create: %s at %.2f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    if (cfg->abi_version != QPB_ABI_VERSION) {
        qpb_set_error("qpb_create: ABI version mismatch (caller %d, library %d)", cfg->abi_version, QPB_ABI_VERSION);
        return QPB_E_INVALID;
    }
    if (cfg->ny <= 0 || cfg->nx <= 0 || cfg->ne <= 0 || cfg->ncell <= 0 || cfg->ncell > (int64_t)cfg->ny * cfg->nx ||
        cfg->nw < 0 || cfg->ngap < 1 || !(cfg->dx > 0.0)) {
        qpb_set_error("qpb_create: invalid dimensions ny=%d nx=%d ne=%d nw=%d ncell=%d ngap=%d dx=%g", cfg->ny, cfg->nx,
                      cfg->ne, cfg->nw, cfg->ncell, cfg->ngap, cfg->dx);
        return QPB_E_INVALID;
    }
    if ((int64_t)cfg->ny * cfg->nx >= (int64_t)1 << 31) {
        qpb_set_error("qpb_create: grid too large for 32-bit cell indices");
        return QPB_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        qpb_set_error("qpb_create: no CUDA device is visible; this library has no CPU fallback");
        return QPB_E_NODEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        qpb_set_error("qpb_create: device %d out of range (0..%d)", cfg->device, ndev - 1);
        return QPB_E_NODEVICE;
    }
    // two attribute queries instead of cudaGetDeviceProperties (measured 4-40 ms per call on this box)
    int cc_major = 0, cc_minor = 0;
    QPB_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device));
    QPB_CUDA(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, cfg->device));
    if (cc_major != 10) {
        qpb_set_error("qpb_create: device %d is sm_%d%d; libqpb is built for sm_100a only", cfg->device, cc_major,
                      cc_minor);
        return QPB_E_NODEVICE;
    }
    lap("device properties");
    QPB_CUDA(cudaSetDevice(cfg->device));
    lap("set device");
    qpb_ctx *c = new qpb_ctx();
    c->cfg = *cfg;
    if (!(c->cfg.diff_tol > 0.0)) c->cfg.diff_tol = 5e-14;   // componentwise: see qpb_prepare_diffusion
    c->ncd = cfg->ny * cfg->nx;
    c->maxit = 512;
    if (const char *e = getenv("QPB_MAXIT")) {   // iteration cap of the sweep iteration (tests: force the Krylov fall-back)
        const int v = atoi(e);
        if (v >= 1 && v <= 512) c->maxit = v;
    }
    auto fail = [&](int rc) {
        qpb_destroy(c);
        return rc;
    };
#define TRY(x)                          \
    do {                                \
        int _r = (x);                   \
        if (_r != QPB_OK) return fail(_r); \
    } while (0)
#define TRYCUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            qpb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,     \
                          cudaGetErrorString(_e));                                                     \
            return fail(QPB_E_CUDA);                                                                   \
        }                                                                                              \
    } while (0)
    TRYCUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    TRYCUDA(cudaEventCreate(&c->ev0));
    TRYCUDA(cudaEventCreate(&c->ev1));
    lap("stream and events");
    const size_t nstate = (size_t)cfg->ne * c->ncd;
    TRY(dev_alloc(&c->d_S, nstate));
    TRYCUDA(cudaMemsetAsync(c->d_S, 0, nstate * sizeof(double), c->stream));
    TRY(dev_alloc(&c->d_flags, (size_t)c->ncd));
    TRY(dev_alloc(&c->d_cell2dense, (size_t)cfg->ncell));
    TRY(dev_alloc(&c->d_integrated, (size_t)cfg->ncell));
    // T1 doubles as the staging buffer for state upload/download even when diffusion is off
    TRY(dev_alloc(&c->d_T1, nstate));
    if (cfg->flags & QPB_F_DIFFUSION) {
        TRY(dev_alloc(&c->d_B, nstate));
        TRY(dev_alloc(&c->d_T2, nstate));
        TRY(dev_alloc(&c->d_bcx, (size_t)c->ncd));
        TRY(dev_alloc(&c->d_bcy, (size_t)c->ncd));
        TRY(dev_alloc(&c->d_srcgeom, (size_t)c->ncd));
        TRY(dev_alloc(&c->d_cx, (size_t)c->ncd));
        TRY(dev_alloc(&c->d_cy, (size_t)c->ncd));
        TRY(dev_alloc(&c->d_res, (size_t)c->maxit * cfg->ne));
        TRY(dev_alloc(&c->d_unorm, (size_t)c->maxit * cfg->ne));
        TRY(dev_alloc(&c->d_done, 2 * (size_t)cfg->ne));
        if (cfg->flags & QPB_F_VARIABLE_D) TRY(dev_alloc(&c->d_Dcell, nstate));
    }
    if (cfg->nw > 0) {
        TRY(dev_alloc(&c->d_P, (size_t)cfg->nw * cfg->ncell));
        TRYCUDA(cudaMemsetAsync(c->d_P, 0, (size_t)cfg->nw * cfg->ncell * sizeof(double), c->stream));
    }
    c->pauli_cap = 1024;
    TRY(dev_alloc(&c->d_pauli, (size_t)c->pauli_cap));
    lap("allocations");
    TRYCUDA(cudaStreamSynchronize(c->stream));
    lap("done");
#undef TRY
#undef TRYCUDA
    *out = c;
    return QPB_OK;
}

extern "C" void qpb_destroy(qpb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    qpbk_free_slot(c->slot[0]);
    qpbk_free_slot(c->slot[1]);
    dev_free(c->d_cx); dev_free(c->d_cy); dev_free(c->d_flags); dev_free(c->d_bcx); dev_free(c->d_bcy); dev_free(c->d_srcgeom);
    dev_free(c->d_cell2dense); dev_free(c->d_Dcell); dev_free(c->d_S); dev_free(c->d_B);
    dev_free(c->d_T1); dev_free(c->d_T2); dev_free(c->d_res); dev_free(c->d_unorm); dev_free(c->d_done);
    dev_free(c->d_Kr); dev_free(c->d_Ks); dev_free(c->d_KrT); dev_free(c->d_KsT); dev_free(c->d_rho);
    dev_free(c->d_gapid); dev_free(c->d_idxd); dev_free(c->d_idxs); dev_free(c->d_idxdT);
    dev_free(c->d_sign); dev_free(c->d_signT); dev_free(c->d_dmap); dev_free(c->d_smap);
    dev_free(c->d_kof); dev_free(c->d_mof); dev_free(c->d_P); dev_free(c->d_K4); dev_free(c->d_Mg); dev_free(c->d_Xn); dev_free(c->d_Xp); dev_free(c->d_scratch); dev_free(c->d_gen); dev_free(c->d_genprog); dev_free(c->d_gen_E); dev_free(c->d_gen_x); dev_free(c->d_gen_y); dev_free(c->d_gen_flag);
    qpbk_free_krylov(c);
    dev_free(c->d_integrated); dev_free(c->d_pauli); dev_free(c->d_xdense); dev_free(c->d_cperm); dev_free(c->d_ggid); dev_free(c->d_euler);
    if (c->d_pauli_part) qpb_dev_free(c->d_pauli_part);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    dev_free(c->d_snap);
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

#define QPB_ENTER(c)                                   \
    do {                                               \
        if (!(c)) {                                    \
            qpb_set_error("null context");             \
            return QPB_E_INVALID;                      \
        }                                              \
        QPB_CUDA(cudaSetDevice((c)->cfg.device));      \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
extern "C" int qpb_upload_geometry(qpb_ctx *c, const uint8_t *mask, const double *bcx, const double *bcy,
                                   const double *source) {
    QPB_ENTER(c);
    const int ny = c->cfg.ny, nx = c->cfg.nx, ncd = c->ncd;
    if (!mask) {
        qpb_set_error("qpb_upload_geometry: mask is null");
        return QPB_E_INVALID;
    }
    c->h_flags.assign(ncd, 0);
    c->h_cell2dense.clear();
    c->h_cell2dense.reserve(c->cfg.ncell);
    bool anyx = false, anyy = false;
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            const int p = y * nx + x;
            if (!mask[p]) continue;
            unsigned f = QPB_IN;
            if (x > 0 && mask[p - 1]) f |= QPB_LK_L;
            if (x + 1 < nx && mask[p + 1]) f |= QPB_LK_R;
            if (y > 0 && mask[p - nx]) f |= QPB_LK_U;
            if (y + 1 < ny && mask[p + nx]) f |= QPB_LK_D;
            if (f & (QPB_LK_L | QPB_LK_R)) anyx = true;
            if (f & (QPB_LK_U | QPB_LK_D)) anyy = true;
            c->h_flags[p] = (uint8_t)f;
            c->h_cell2dense.push_back(p);
        }
    if ((int)c->h_cell2dense.size() != c->cfg.ncell) {
        qpb_set_error("qpb_upload_geometry: mask has %zu cells, config says %d", c->h_cell2dense.size(), c->cfg.ncell);
        return QPB_E_INVALID;
    }
    c->thin_x = !anyy;  // no vertical links: every row is an independent 1-D problem along x
    c->thin_y = !anyx;
    QPB_CUDA(cudaMemcpyAsync(c->d_flags, c->h_flags.data(), ncd, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(c->d_cell2dense, c->h_cell2dense.data(), sizeof(int32_t) * c->cfg.ncell,
                             cudaMemcpyHostToDevice, c->stream));
    if (c->cfg.flags & QPB_F_DIFFUSION) {
        if (!bcx || !bcy || !source) {
            qpb_set_error("qpb_upload_geometry: boundary arrays are required when diffusion is enabled");
            return QPB_E_INVALID;
        }
        {   // host copies (qpb_prepare_diffusion reads them), first touched by three threads at once
            std::thread ta([&] { c->h_bcx.assign(bcx, bcx + ncd); }), tb([&] { c->h_bcy.assign(bcy, bcy + ncd); });
            c->h_src.assign(source, source + ncd);
            ta.join();
            tb.join();
        }
        QPB_CUDA(cudaMemcpyAsync(c->d_bcx, bcx, sizeof(double) * ncd, cudaMemcpyHostToDevice, c->stream));
        QPB_CUDA(cudaMemcpyAsync(c->d_bcy, bcy, sizeof(double) * ncd, cudaMemcpyHostToDevice, c->stream));
        QPB_CUDA(cudaMemcpyAsync(c->d_srcgeom, source, sizeof(double) * ncd, cudaMemcpyHostToDevice, c->stream));
        // Gershgorin bounds and a commutator probe of Gx, Gy (unit coefficient); the passes over the grid run on a few
        // host threads (row blocks): at 2048 x 2048 they were 0.15 s of every run's setup
        const int nth = std::max(1, std::min(8, std::min(ny / 64, (int)std::thread::hardware_concurrency())));
        auto rows_parallel = [&](auto &&fn) {
            if (nth <= 1) {
                fn(0, 0, ny);
                return;
            }
            std::vector<std::thread> th;
            for (int t = 0; t < nth; ++t)
                th.emplace_back([&, t] { fn(t, (int)((long long)ny * t / nth), (int)((long long)ny * (t + 1) / nth)); });
            for (auto &x : th) x.join();
        };
        std::vector<double> pgx(nth, 0.0), pgy(nth, 0.0), pgm(nth, 0.0);
        std::vector<double> r((size_t)ncd);
        rows_parallel([&](int t, int ya, int yb) {
            double gx = 0.0, gy = 0.0, gmin = 0.0;
            for (int p = ya * nx; p < yb * nx; ++p) {
                const unsigned f = c->h_flags[p];
                double rv = 0.0;
                if (f & QPB_IN) {
                    const int dxl = ((f & QPB_LK_L) ? 1 : 0) + ((f & QPB_LK_R) ? 1 : 0);
                    const int dyl = ((f & QPB_LK_U) ? 1 : 0) + ((f & QPB_LK_D) ? 1 : 0);
                    gx = std::max(gx, 2.0 * dxl + std::fabs(bcx[p]));
                    gy = std::max(gy, 2.0 * dyl + std::fabs(bcy[p]));
                    gmin = std::min(gmin, std::min(bcx[p], bcy[p]));
                    // probe vector: a fixed pseudo-random value in [0.5, 1.5) per cell (splitmix64 of the index)
                    unsigned long long z = (unsigned long long)p * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
                    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                    z ^= z >> 31;
                    rv = 0.5 + (double)(z >> 11) * (1.0 / 9007199254740992.0);
                }
                r[p] = rv;
            }
            pgx[t] = gx; pgy[t] = gy; pgm[t] = gmin;
        });
        double gx = 0.0, gy = 0.0, gmin = 0.0;
        for (int t = 0; t < nth; ++t) {
            gx = std::max(gx, pgx[t]);
            gy = std::max(gy, pgy[t]);
            gmin = std::min(gmin, pgm[t]);
        }
        c->gmin = gmin;
        c->gmax_x = gx;
        c->gmax_y = gy;
        // (Gy Gx - Gx Gy) r in one pass: both products of a cell from the 3 x 3 neighbourhood of r, no grid-sized
        // temporaries (they were 130 MB of first-touch page faults at 2048 x 2048)
        auto gxr = [&](int p) {   // (Gx r)(p)
            const unsigned f = c->h_flags[p];
            if (!(f & QPB_IN)) return 0.0;
            double v = bcx[p] * r[p];
            if (f & QPB_LK_L) v += r[p] - r[p - 1];
            if (f & QPB_LK_R) v += r[p] - r[p + 1];
            return v;
        };
        auto gyr = [&](int p) {   // (Gy r)(p)
            const unsigned f = c->h_flags[p];
            if (!(f & QPB_IN)) return 0.0;
            double v = bcy[p] * r[p];
            if (f & QPB_LK_U) v += r[p] - r[p - nx];
            if (f & QPB_LK_D) v += r[p] - r[p + nx];
            return v;
        };
        std::vector<double> pd(nth, 0.0), pv(nth, 0.0);
        rows_parallel([&](int t, int ya, int yb) {
            double dm = 0.0, vm = 0.0;
            for (int p = ya * nx; p < yb * nx; ++p) {
                const unsigned f = c->h_flags[p];
                if (!(f & QPB_IN)) continue;
                const double ax = gxr(p), ay = gyr(p);
                double yx = bcy[p] * ax, xy = bcx[p] * ay;   // Gy (Gx r), Gx (Gy r)
                if (f & QPB_LK_U) yx += ax - gxr(p - nx);
                if (f & QPB_LK_D) yx += ax - gxr(p + nx);
                if (f & QPB_LK_L) xy += ay - gyr(p - 1);
                if (f & QPB_LK_R) xy += ay - gyr(p + 1);
                dm = std::max(dm, std::fabs(yx - xy));
                vm = std::max(vm, std::max(std::fabs(yx), std::fabs(xy)));
            }
            pd[t] = dm; pv[t] = vm;
        });
        double dmax = 0.0, vmax = 0.0;
        for (int t = 0; t < nth; ++t) {
            dmax = std::max(dmax, pd[t]);
            vmax = std::max(vmax, pv[t]);
        }
        c->commuting = dmax <= 1e-12 * std::max(vmax, 1.0);
    }
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_geom = true;
    qpbk_free_slot(c->slot[0]);
    qpbk_free_slot(c->slot[1]);
    return QPB_OK;
}

extern "C" int qpb_upload_diffusion(qpb_ctx *c, const double *D) {
    QPB_ENTER(c);
    if (!(c->cfg.flags & QPB_F_DIFFUSION) || !D) {
        qpb_set_error("qpb_upload_diffusion: diffusion is disabled on this context or D is null");
        return QPB_E_INVALID;
    }
    const bool vard = c->cfg.flags & QPB_F_VARIABLE_D;
    const size_t n = vard ? (size_t)c->cfg.ne * c->cfg.ncell : (size_t)c->cfg.ne;
    c->h_D.assign(D, D + n);
    for (size_t i = 0; i < n; ++i)
        if (!(c->h_D[i] >= 0.0) || !std::isfinite(c->h_D[i])) {
            qpb_set_error("qpb_upload_diffusion: D must be finite and non-negative");
            return QPB_E_INVALID;
        }
    c->have_D = true;
    qpbk_free_slot(c->slot[0]);
    qpbk_free_slot(c->slot[1]);
    return QPB_OK;
}

// ---- shift parameters -----------------------------------------------------------------------------------
// Jacobi dn(u|m) and K(m) by the arithmetic-geometric mean (Abramowitz & Stegun 16.4, 17.6).
static void agm_dn(double m, const std::vector<double> &frac, std::vector<double> &dn_out) {
    double a[32], cc[32];
    a[0] = 1.0;
    double b = std::sqrt(std::max(1.0 - m, 0.0));
    cc[0] = std::sqrt(m);
    int N = 0;
    while (N < 30 && std::fabs(cc[N]) > 1e-16 * a[N]) {
        const double an = 0.5 * (a[N] + b);
        cc[N + 1] = 0.5 * (a[N] - b);
        b = std::sqrt(a[N] * b);
        a[N + 1] = an;
        ++N;
    }
    const double K = M_PI / (2.0 * a[N]);
    dn_out.resize(frac.size());
    for (size_t i = 0; i < frac.size(); ++i) {
        const double u = frac[i] * K;
        double phi = std::ldexp(a[N] * u, N);
        double phi_prev = phi;
        for (int n = N; n >= 1; --n) {
            phi_prev = phi;
            phi = 0.5 * (phi + std::asin(cc[n] / a[n] * std::sin(phi)));
        }
        const double den = std::cos(phi_prev - phi);
        dn_out[i] = (N == 0 || std::fabs(den) < 1e-300) ? 1.0 : std::cos(phi) / den;
    }
}

// Optimal J-parameter ADI shifts on [lo,hi] (Wachspress): r_j = hi * dn((2j-1)K/(2J), 1-(lo/hi)^2).
static std::vector<double> wachspress(double lo, double hi, int J) {
    std::vector<double> out(J);
    if (hi <= lo * (1.0 + 1e-12)) {
        std::fill(out.begin(), out.end(), std::sqrt(lo * hi));
        return out;
    }
    const double kp = lo / hi;
    std::vector<double> frac(J), dn;
    for (int j = 0; j < J; ++j) frac[j] = (2.0 * j + 1.0) / (2.0 * J);
    agm_dn(1.0 - kp * kp, frac, dn);
    for (int j = 0; j < J; ++j) out[j] = std::min(hi, std::max(lo, hi * dn[j]));
    return out;
}

static double adi_bound(double lo, double hi, const std::vector<double> &sh) {
    const int NP = 600;
    double worst = 0.0;
    for (int i = 0; i <= NP; ++i) {
        const double h = lo * std::pow(hi / lo, (double)i / NP);
        double f = 1.0;
        for (double r : sh) f *= std::fabs((r - h) / (r + h));
        worst = std::max(worst, f);
    }
    return worst * worst;
}

static std::vector<double> plan_shifts(double lo, double hi, double target, int jcap) {
    if (hi <= lo * (1.0 + 1e-9)) return {std::sqrt(lo * hi)};
    int J = (int)std::ceil(std::log(4.0 / target) * std::log(4.0 * hi / lo) / (M_PI * M_PI));
    J = std::max(1, std::min(J, jcap));
    while (J > 1 && adi_bound(lo, hi, wachspress(lo, hi, J - 1)) <= target) --J;
    while (J < jcap && adi_bound(lo, hi, wachspress(lo, hi, J)) > target) ++J;
    return wachspress(lo, hi, J);
}

extern "C" int qpb_prepare_diffusion(qpb_ctx *c, int slot, double dt) {
    QPB_ENTER(c);
    if (slot < 0 || slot > 1 || !(dt > 0.0)) {
        qpb_set_error("qpb_prepare_diffusion: bad slot %d or dt %g", slot, dt);
        return QPB_E_INVALID;
    }
    if (!(c->cfg.flags & QPB_F_DIFFUSION) || !c->have_geom || !c->have_D) {
        qpb_set_error("qpb_prepare_diffusion: geometry and diffusion coefficients must be uploaded first");
        return QPB_E_INVALID;
    }
    DiffSlot &s = c->slot[slot];
    qpbk_free_slot(s);
    const auto &cf = c->cfg;
    const int ne = cf.ne, ncd = c->ncd, nx = cf.nx;
    const bool vard = cf.flags & QPB_F_VARIABLE_D;
    const double inv_dx = 1.0 / cf.dx, inv_dx2 = inv_dx * inv_dx;
    s.dt = dt;
    s.mode = c->thin_x ? 1 : (c->thin_y ? 2 : 0);
    s.commuting = c->commuting && !vard;
    s.a_bin.assign(ne, 0.0);
    std::vector<double> hi_bin(ne, 0.5);
    std::vector<double> srccoef(ne, 0.0);
    if (!vard) {
        for (int i = 0; i < ne; ++i) {
            const double a = 0.5 * dt * c->h_D[i] * inv_dx2;
            s.a_bin[i] = a;
            hi_bin[i] = 0.5 + a * std::max(c->gmax_x, c->gmax_y);
            srccoef[i] = dt * c->h_D[i];
        }
        QPB_ALLOC(s.d_src, ne);
        QPB_CUDA(cudaMemcpyAsync(s.d_src, srccoef.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
    } else {
        // dense per-bin coefficient fields (solver.py:275-318): harmonic-mean faces, D-scaled boundary terms
        const size_t nst = (size_t)ne * ncd;
        std::vector<double> ex(nst, 0.0), ey(nst, 0.0), gbx(nst, 0.0), gby(nst, 0.0), src(nst, 0.0), Dd((size_t)ncd);
        for (int i = 0; i < ne; ++i) {
            std::fill(Dd.begin(), Dd.end(), 0.0);
            for (int q = 0; q < cf.ncell; ++q) Dd[c->h_cell2dense[q]] = c->h_D[(size_t)i * cf.ncell + q];
            double hx = 0.0, hy = 0.0;
            const size_t o = (size_t)i * ncd;
            for (int p = 0; p < ncd; ++p) {
                const unsigned f = c->h_flags[p];
                if (!(f & QPB_IN)) continue;
                const double Dp = Dd[p];
                if (f & QPB_LK_L) {
                    const double Dq = Dd[p - 1];
                    ex[o + p] = 0.5 * dt * (2.0 * Dp * Dq / std::max(Dp + Dq, 1e-30) * inv_dx2);
                }
                if (f & QPB_LK_U) {
                    const double Dq = Dd[p - nx];
                    ey[o + p] = 0.5 * dt * (2.0 * Dp * Dq / std::max(Dp + Dq, 1e-30) * inv_dx2);
                }
                gbx[o + p] = 0.5 * dt * (Dp * c->h_bcx[p] * inv_dx2);
                gby[o + p] = 0.5 * dt * (Dp * c->h_bcy[p] * inv_dx2);
                src[o + p] = dt * (Dp * c->h_src[p]);
            }
            for (int p = 0; p < ncd; ++p) {
                const unsigned f = c->h_flags[p];
                if (!(f & QPB_IN)) continue;
                const double eL = ex[o + p], eR = (f & QPB_LK_R) ? ex[o + p + 1] : 0.0;
                const double eU = ey[o + p], eD = (f & QPB_LK_D) ? ey[o + p + nx] : 0.0;
                hx = std::max(hx, 2.0 * (eL + eR) + std::fabs(gbx[o + p]));
                hy = std::max(hy, 2.0 * (eU + eD) + std::fabs(gby[o + p]));
            }
            hi_bin[i] = 0.5 + std::max(hx, hy);
        }
        QPB_ALLOC(s.d_ex, nst);
        QPB_ALLOC(s.d_ey, nst);
        QPB_ALLOC(s.d_gbx, nst);
        QPB_ALLOC(s.d_gby, nst);
        QPB_ALLOC(s.d_src, nst);
        QPB_CUDA(cudaMemcpy(s.d_ex, ex.data(), sizeof(double) * nst, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(s.d_ey, ey.data(), sizeof(double) * nst, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(s.d_gbx, gbx.data(), sizeof(double) * nst, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(s.d_gby, gby.data(), sizeof(double) * nst, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(s.d_src, src.data(), sizeof(double) * nst, cudaMemcpyHostToDevice));
    }
    // shift tables
    std::vector<std::vector<double>> sh(ne);
    s.jlen.assign(ne, 1);
    int jmax = 1;
    const double target = std::min(1e-2, std::max(cf.diff_tol * 0.1, 1e-15));
    for (int i = 0; i < ne; ++i) {
        // spectrum of H, V = I/2 + alpha*G inside [lo, hi]; G is positive semi-definite unless a Robin face has a negative
        // beta, which lowers the bound by alpha*|beta| (vard: the dense coefficient fields carry the same sign)
        const double amax = vard ? (hi_bin[i] - 0.5) / std::max(1.0, std::max(c->gmax_x, c->gmax_y)) : s.a_bin[i];
        const double lo = 0.5 + amax * c->gmin, hi = std::max(hi_bin[i], 0.5);
        if (s.mode == 0 && !(lo > 0.05)) {
            qpb_set_error("qpb_prepare_diffusion: a boundary face with a negative Robin coefficient makes the split "
                          "operators indefinite at this step length (bin %d: lower spectral bound %.3g); the sweep "
                          "iteration does not cover that case", i, lo);
            return QPB_E_INVALID;
        }
        if (s.mode != 0) sh[i] = {0.0};
        else if (s.commuting) sh[i] = plan_shifts(lo, hi, target, 48);
        else {
            const int J = 4;  // cyclic geometric set (box D); more shifts do not help a
                              // non-commuting pair (CPU study: 57 / 81 / 101 iterations for J = 4 / 8 / 16 at a = 40)
            sh[i].resize(J);
            for (int k = 1; k <= J; ++k) sh[i][k - 1] = hi * std::pow(lo / hi, (2.0 * k - 1.0) / (2.0 * J));
        }
        s.jlen[i] = (int)sh[i].size();
        jmax = std::max(jmax, s.jlen[i]);
    }
    s.jmax = jmax;
    // Stiff steps on a non-commuting geometry leave the sweep iteration for the preconditioned Krylov solve
    // (qpb_krylov.cu): the cyclic-shift iteration stops contracting once hi/lo reaches a few hundred.
    {
        double ratio = 1.0;
        std::vector<double> ksh(ne, 1.0);
        for (int i = 0; i < ne; ++i) {
            const double amax = vard ? (hi_bin[i] - 0.5) / std::max(1.0, std::max(c->gmax_x, c->gmax_y)) : s.a_bin[i];
            const double lo = std::max(0.05, 0.5 + amax * c->gmin), hi = std::max(hi_bin[i], 0.5);
            ratio = std::max(ratio, hi / lo);
            ksh[i] = std::sqrt(lo * hi);
        }
        double limit = 200.0;
        if (const char *e = getenv("QPB_KRYLOV_RATIO")) limit = atof(e);
        s.krylov = s.mode == 0 && !s.commuting && ratio > limit;
        QPB_ALLOC(s.d_kshift, ne);
        QPB_CUDA(cudaMemcpyAsync(s.d_kshift, ksh.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));
    }
    std::vector<double> flat((size_t)ne * jmax, 0.5);
    for (int i = 0; i < ne; ++i)
        for (int k = 0; k < s.jlen[i]; ++k) flat[(size_t)i * jmax + k] = sh[i][k];
    QPB_ALLOC(s.d_a, ne);
    QPB_ALLOC(s.d_shift, (size_t)ne * jmax);
    QPB_ALLOC(s.d_jlen, ne);
    // Stop test of the sweep iteration, componentwise (Oettli-Prager): |b - Au|_i <= tol (|A||u| + |b|)_i in EVERY
    // cell.  A max-norm test ||b - Au|| <= tol ||u|| leaves cells far below the bin's peak unconstrained (boundary
    // sources next to a cold interior, lognormal fields): measured on such a case the element-wise error still fell
    // from 1e-7 to 1e-12 while the max-norm residual sat at its rounding floor.  The componentwise residual tracks the
    // element-wise relative error of the solution (within 10x on masks, up to a few 1000x on the first steps from a
    // rough field), so tol = 5e-14 keeps every cell within the 1e-9 bar.  Its rounding floor is a few ulps whatever alpha is.
    // d_tolk is the max-norm tolerance of the Krylov path (qpb_krylov.cu), floored at the fp64 resolution of a
    // residual whose terms are (1 + 2 max row sum of alpha*G) |u| large.
    std::vector<double> tolb(ne), tolk(ne);
    for (int i = 0; i < ne; ++i) {
        tolb[i] = std::max(cf.diff_tol, 1e-14);
        tolk[i] = std::max(cf.diff_tol, 1e-15 * (1.0 + 2.0 * (hi_bin[i] - 0.5)));
    }
    QPB_ALLOC(s.d_tol, ne);
    QPB_ALLOC(s.d_tolk, ne);
    QPB_CUDA(cudaMemcpyAsync(s.d_tol, tolb.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(s.d_tolk, tolk.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
    QPB_ALLOC(s.d_known, ne);
    QPB_CUDA(cudaMemsetAsync(s.d_known, 0, sizeof(int) * ne, c->stream));
    s.known_iters = 0;
    s.solves = 0;
    QPB_CUDA(cudaMemcpyAsync(s.d_a, s.a_bin.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(s.d_shift, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(s.d_jlen, s.jlen.data(), sizeof(int) * ne, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    s.launch_iters = s.commuting ? std::min(jmax, 16) : 8;
    s.ready = true;
    s.fast = false;
    if (!vard) {
        // a full rectangle that takes the direct spectral solve needs none of the sweep tables
        int rc = qpbk_prepare_spectral(c, s);
        if (rc != QPB_OK) return rc;
        if (!s.spectral) {
            if ((rc = qpbk_prepare_fast(c, s)) != QPB_OK) return rc;
            if ((rc = qpbr_plan(c, s)) != QPB_OK) return rc;
        }
    } else if (!s.krylov) {
        int rc = qpbk_prepare_fast(c, s);   // chunked sweeps with per-line pivot tables
        if (rc != QPB_OK) return rc;
    }
    QPB_CUDA(cudaDeviceSynchronize());   // blocking legacy-stream copies of the table setup (see qpb_upload_collision)
    c->diag.direct_mode = s.mode != 0;
    c->diag.commuting = s.commuting;
    c->diag.sweep_path = s.spectral ? 4 : s.res.ok ? 5 : !s.fast ? 0
                         : !(s.pipe.x_ok && s.pipe.y_ok) ? 1
                         : (s.pipe.x_nseg > 1 || s.pipe.y_nseg > 1) ? 3 : 2;
    return QPB_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int qpb_upload_collision(qpb_ctx *c, const double *K_r0, const double *K_s0, const double *rho,
                                    const int32_t *gap_id, const int32_t *idx_diff, const int32_t *idx_sum,
                                    const int8_t *sign) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    const int ne = cf.ne, ng = cf.ngap;
    if (!rho) {
        qpb_set_error("qpb_upload_collision: rho is required");
        return QPB_E_INVALID;
    }
    const bool scat = cf.flags & QPB_F_SCATTERING, rec = cf.flags & QPB_F_RECOMBINATION;
    if ((scat || rec) && (cf.nw <= 0 || !idx_diff || !idx_sum || !sign)) {
        qpb_set_error("qpb_upload_collision: phonon index maps are required when collisions are enabled");
        return QPB_E_INVALID;
    }
    if ((scat && !K_s0) || (rec && !K_r0)) {
        qpb_set_error("qpb_upload_collision: kernel matrix missing for an enabled process");
        return QPB_E_INVALID;
    }
    const size_t nk = (size_t)ng * ne * ne;
    auto up_mat = [&](const double *src, double *&d, double *&dT) -> int {
        if (!src) return QPB_OK;
        std::vector<double> T(nk);
        for (int g = 0; g < ng; ++g)
            for (int i = 0; i < ne; ++i)
                for (int j = 0; j < ne; ++j)
                    T[((size_t)g * ne + j) * ne + i] = src[((size_t)g * ne + i) * ne + j];
        dev_free(d);
        dev_free(dT);
        QPB_ALLOC(d, nk);
        QPB_ALLOC(dT, nk);
        QPB_CUDA(cudaMemcpy(d, src, sizeof(double) * nk, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(dT, T.data(), sizeof(double) * nk, cudaMemcpyHostToDevice));
        return QPB_OK;
    };
    int rc;
    if ((rc = up_mat(rec ? K_r0 : nullptr, c->d_Kr, c->d_KrT)) != QPB_OK) return rc;
    if ((rc = up_mat(scat ? K_s0 : nullptr, c->d_Ks, c->d_KsT)) != QPB_OK) return rc;
    dev_free(c->d_rho);
    QPB_ALLOC(c->d_rho, (size_t)ng * ne);
    QPB_CUDA(cudaMemcpy(c->d_rho, rho, sizeof(double) * ng * ne, cudaMemcpyHostToDevice));
    dev_free(c->d_gapid);
    c->h_gapid.clear();
    if (gap_id && ng > 1) {
        c->h_gapid.assign(gap_id, gap_id + cf.ncell);
        for (int q = 0; q < cf.ncell; ++q)
            if (gap_id[q] < 0 || gap_id[q] >= ng) {
                qpb_set_error("qpb_upload_collision: gap_id[%d]=%d out of range", q, gap_id[q]);
                return QPB_E_INVALID;
            }
        QPB_ALLOC(c->d_gapid, cf.ncell);
        QPB_CUDA(cudaMemcpy(c->d_gapid, gap_id, sizeof(int32_t) * cf.ncell, cudaMemcpyHostToDevice));
    }
    c->structured = false;
    if (scat || rec) {
        const size_t nn = (size_t)ne * ne;
        for (size_t k = 0; k < nn; ++k)
            if (idx_diff[k] < 0 || idx_diff[k] >= cf.nw || idx_sum[k] < 0 || idx_sum[k] >= cf.nw) {
                qpb_set_error("qpb_upload_collision: phonon index out of range at pair %zu", k);
                return QPB_E_INVALID;
            }
        std::vector<int32_t> idT(nn);
        std::vector<int8_t> sgT(nn);
        for (int i = 0; i < ne; ++i)
            for (int j = 0; j < ne; ++j) {
                idT[(size_t)j * ne + i] = idx_diff[(size_t)i * ne + j];
                sgT[(size_t)j * ne + i] = sign[(size_t)i * ne + j];
            }
        dev_free(c->d_idxd); dev_free(c->d_idxs); dev_free(c->d_idxdT); dev_free(c->d_sign); dev_free(c->d_signT);
        QPB_ALLOC(c->d_idxd, nn);
        QPB_ALLOC(c->d_idxs, nn);
        QPB_ALLOC(c->d_idxdT, nn);
        QPB_ALLOC(c->d_sign, nn);
        QPB_ALLOC(c->d_signT, nn);
        QPB_CUDA(cudaMemcpy(c->d_idxd, idx_diff, sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(c->d_idxs, idx_sum, sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(c->d_idxdT, idT.data(), sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(c->d_sign, sign, sizeof(int8_t) * nn, cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(c->d_signT, sgT.data(), sizeof(int8_t) * nn, cudaMemcpyHostToDevice));
        // Structure probe: idx_diff depends on |i-j| only, idx_sum on i+j only, both injective, sign = sign(i-j),
        // kernels symmetric, idx_sum symmetric.  True for every uniform energy grid the solver builds.
        bool ok = true;
        std::vector<int32_t> dmap(ne, -1), smap(2 * ne - 1, -1);
        for (int i = 0; i < ne && ok; ++i)
            for (int j = 0; j < ne && ok; ++j) {
                const int k = std::abs(i - j), m = i + j;
                const int32_t d = idx_diff[(size_t)i * ne + j], sidx = idx_sum[(size_t)i * ne + j];
                if (dmap[k] < 0) dmap[k] = d; else if (dmap[k] != d) ok = false;
                if (smap[m] < 0) smap[m] = sidx; else if (smap[m] != sidx) ok = false;
                const int sg = (i > j) - (i < j);
                if (sign[(size_t)i * ne + j] != sg) ok = false;
            }
        std::vector<int32_t> kof(cf.nw, -1), mof(cf.nw, -1);
        for (int k = 0; k < ne && ok; ++k) {
            if (kof[dmap[k]] >= 0) ok = false;
            kof[dmap[k]] = k;
        }
        for (int m = 0; m < 2 * ne - 1 && ok; ++m) {
            if (mof[smap[m]] >= 0) ok = false;
            mof[smap[m]] = m;
        }
        auto symmetric = [&](const double *K) {
            if (!K) return true;
            for (int g = 0; g < ng; ++g)
                for (int i = 0; i < ne; ++i)
                    for (int j = 0; j < i; ++j)
                        if (K[((size_t)g * ne + i) * ne + j] != K[((size_t)g * ne + j) * ne + i]) return false;
            return true;
        };
        ok = ok && symmetric(scat ? K_s0 : nullptr) && symmetric(rec ? K_r0 : nullptr);
        if (const char *env = getenv("QPB_FORCE_GENERIC"))   // test hook: exercise the generic kernels
            if (env[0] == '1') ok = false;
        if (ok && scat)
            for (int g = 0; g < ng && ok; ++g)
                for (int i = 0; i < ne; ++i)
                    if (K_s0[((size_t)g * ne + i) * ne + i] != 0.0) ok = false;
        dev_free(c->d_dmap); dev_free(c->d_smap); dev_free(c->d_kof); dev_free(c->d_mof);
        if (ok) {
            QPB_ALLOC(c->d_dmap, ne);
            QPB_ALLOC(c->d_smap, 2 * ne - 1);
            QPB_ALLOC(c->d_kof, cf.nw);
            QPB_ALLOC(c->d_mof, cf.nw);
            QPB_CUDA(cudaMemcpy(c->d_dmap, dmap.data(), sizeof(int32_t) * ne, cudaMemcpyHostToDevice));
            QPB_CUDA(cudaMemcpy(c->d_smap, smap.data(), sizeof(int32_t) * (2 * ne - 1), cudaMemcpyHostToDevice));
            QPB_CUDA(cudaMemcpy(c->d_kof, kof.data(), sizeof(int32_t) * cf.nw, cudaMemcpyHostToDevice));
            QPB_CUDA(cudaMemcpy(c->d_mof, mof.data(), sizeof(int32_t) * cf.nw, cudaMemcpyHostToDevice));
            c->h_dmap = dmap; c->h_smap = smap; c->h_kof = kof; c->h_mof = mof;
            c->structured = true;
        }
    }
    c->have_coll = true;
    int rcs = qpbk_collision_setup(c);
    // packed effective kernels (frozen cell-independent phonons) were built from the PREVIOUS tables: re-pack them from
    // the occupations they were made of, so that a table upload on a live context cannot leave stale products behind
    const bool was_uniform = c->uniform_ph;
    c->uniform_ph = false;
    c->gemm_ready = false;
    if (rcs == QPB_OK && was_uniform && (int)c->h_ph_bins.size() == cf.nw) {
        const std::vector<double> bins = c->h_ph_bins;
        rcs = qpbk_uniform_setup(c, bins.data(), true);
    }
    // the setup code uses blocking copies on the legacy stream, which the context's non-blocking stream is not
    // ordered against: everything has landed before the caller can enqueue a step
    QPB_CUDA(cudaDeviceSynchronize());
    return rcs;
}

// ---------------------------------------------------------------------------------------------------------
// ---- staged device -> host copies --------------------------------------------------------------------------
// The caller's buffers are fresh pageable numpy arrays: a plain cudaMemcpy of a 64 MiB snapshot ran at ~5 GB/s
// (driver-side staging plus the first-touch page faults of the destination, all on one thread).  Large downloads go
// through two process-wide pinned chunks instead: chunk k+1 crosses PCIe while chunk k is copied out by a few host
// threads (which also spreads the page faults; pre-faulting the destination with MADV_POPULATE_WRITE measured slower:
// 10 ms instead of 6.3 ms per 64 MiB).  QPB_STAGED_D2H=0 restores the plain copy.
namespace {
constexpr size_t kStageChunk = (size_t)8 << 20;
std::mutex g_stage_mu;
void *g_stage_pin[2] = {nullptr, nullptr};

int host_copy_threads() {
    static const int n = [] {
        const char *e = getenv("QPB_COPY_THREADS");
        const int v = e ? atoi(e) : 8;   // measured on the 16-core GPU box: 4 threads 10.5 GB/s, 8 threads 12.8 GB/s
        return std::max(1, std::min(v, 16));
    }();
    return n;
}

void host_copy_parallel(char *dst, const char *src, size_t bytes) {
    const int nthr = host_copy_threads();
    const size_t per = ((bytes / nthr) + 4095) & ~(size_t)4095;
    std::thread th[16];
    int started = 0;
    for (int t = 1; t < nthr; ++t) {
        const size_t off = per * t;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th[started++] = std::thread([=]() { memcpy(dst + off, src + off, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (int t = 0; t < started; ++t) th[t].join();
}
}  // namespace

// copy `bytes` from device memory to a pageable host buffer, ordered on `stream`; returns after the data has landed
static cudaError_t d2h_staged(void *dst, const void *dsrc, size_t bytes, cudaStream_t stream) {
    static const bool enabled = !(getenv("QPB_STAGED_D2H") && getenv("QPB_STAGED_D2H")[0] == '0');
    if (!enabled || bytes < 2 * kStageChunk) {
        cudaError_t e = cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, stream);
        return e != cudaSuccess ? e : cudaStreamSynchronize(stream);
    }
    std::lock_guard<std::mutex> lk(g_stage_mu);
    if (!g_stage_pin[0]) {
        for (int b = 0; b < 2; ++b) {
            cudaError_t e = cudaHostAlloc(&g_stage_pin[b], kStageChunk, cudaHostAllocPortable);
            if (e != cudaSuccess) {
                if (b == 1) cudaFreeHost(g_stage_pin[0]);
                g_stage_pin[0] = g_stage_pin[1] = nullptr;
                return e;
            }
        }
    }
    cudaEvent_t ev[2];   // events belong to the current device: made per call, the pinned chunks are portable
    for (int b = 0; b < 2; ++b) {
        cudaError_t e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            if (b == 1) cudaEventDestroy(ev[0]);
            return e;
        }
    }
    cudaError_t rc = cudaSuccess;
    const size_t nchunk = (bytes + kStageChunk - 1) / kStageChunk;
    for (size_t k = 0; k <= nchunk && rc == cudaSuccess; ++k) {
        if (k < nchunk) {   // chunk k-2, the previous user of this pinned buffer, was copied out an iteration ago
            const size_t off = k * kStageChunk, len = std::min(kStageChunk, bytes - off);
            rc = cudaMemcpyAsync(g_stage_pin[k & 1], (const char *)dsrc + off, len, cudaMemcpyDeviceToHost, stream);
            if (rc == cudaSuccess) rc = cudaEventRecord(ev[k & 1], stream);
            if (rc != cudaSuccess) break;
        }
        if (k >= 1) {
            const size_t off = (k - 1) * kStageChunk, len = std::min(kStageChunk, bytes - off);
            rc = cudaEventSynchronize(ev[(k - 1) & 1]);
            if (rc != cudaSuccess) break;
            host_copy_parallel((char *)dst + off, (const char *)g_stage_pin[(k - 1) & 1], len);
        }
    }
    if (rc != cudaSuccess) cudaStreamSynchronize(stream);   // nothing may still target the pinned chunks
    for (int b = 0; b < 2; ++b) cudaEventDestroy(ev[b]);
    return rc;
}

extern "C" int qpb_set_state(qpb_ctx *c, const double *n, const double *n_ph) {
    QPB_ENTER(c);
    if (!c->have_geom || !n) {
        qpb_set_error("qpb_set_state: upload the geometry first and pass a state");
        return QPB_E_INVALID;
    }
    const auto &cf = c->cfg;
    QPB_CUDA(cudaMemcpyAsync(c->d_T1, n, sizeof(double) * (size_t)cf.ne * cf.ncell, cudaMemcpyHostToDevice, c->stream));
    int rc = qpbk_scatter_state(c, c->d_T1);
    if (rc != QPB_OK) return rc;
    if (cf.nw > 0 && n_ph)
        QPB_CUDA(cudaMemcpyAsync(c->d_P, n_ph, sizeof(double) * (size_t)cf.nw * cf.ncell, cudaMemcpyHostToDevice,
                                 c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    if (cf.nw > 0 && n_ph) {
        const int rcu = qpbk_uniform_setup(c, n_ph);
        QPB_CUDA(cudaDeviceSynchronize());
        return rcu;
    }
    return QPB_OK;
}

extern "C" int qpb_set_state_uniform_phonons(qpb_ctx *c, const double *n, const double *n_ph_bins) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (!c->have_geom || !n || (cf.nw > 0 && !n_ph_bins)) {
        qpb_set_error("qpb_set_state_uniform_phonons: upload the geometry first and pass both arrays");
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaMemcpyAsync(c->d_T1, n, sizeof(double) * (size_t)cf.ne * cf.ncell, cudaMemcpyHostToDevice, c->stream));
    int rc = qpbk_scatter_state(c, c->d_T1);
    if (rc != QPB_OK) return rc;
    if (cf.nw > 0) {
        double *d_bins = nullptr;
        QPB_ALLOC(d_bins, cf.nw);
        QPB_CUDA(cudaMemcpyAsync(d_bins, n_ph_bins, sizeof(double) * cf.nw, cudaMemcpyHostToDevice, c->stream));
        rc = qpbk_broadcast_phonons(c, d_bins);
        QPB_CUDA(cudaStreamSynchronize(c->stream));
        dev_free(d_bins);
        if (rc != QPB_OK) return rc;
        rc = qpbk_uniform_setup(c, n_ph_bins, true);
        QPB_CUDA(cudaDeviceSynchronize());
        return rc;
    }
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_set_state_separable(qpb_ctx *c, const double *weights, const double *spatial,
                                       const double *n_ph_bins) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (!c->have_geom || !weights || !spatial || (cf.nw > 0 && !n_ph_bins)) {
        qpb_set_error("qpb_set_state_separable: upload the geometry first and pass weights, spatial values and "
                      "(with collisions) the phonon occupations");
        return QPB_E_INVALID;
    }
    // the factors ride in the front of the work array T1 (free between steps): [weights | spatial | phonon bins];
    // tiny grids (a 1 x 1 mask with one bin) take a block of their own
    const size_t nfac = (size_t)cf.ne + cf.ncell + cf.nw;
    double *d_own = nullptr;
    if (nfac > (size_t)cf.ne * c->ncd) QPB_ALLOC(d_own, nfac);
    struct Release {
        double *p;
        ~Release() { dev_free(p); }
    } release{d_own};
    double *d_w = d_own ? d_own : c->d_T1, *d_sp = d_w + cf.ne, *d_bins = d_sp + cf.ncell;
    QPB_CUDA(cudaMemcpyAsync(d_w, weights, sizeof(double) * cf.ne, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(d_sp, spatial, sizeof(double) * cf.ncell, cudaMemcpyHostToDevice, c->stream));
    int rc = qpbk_outer_state(c, d_w, d_sp);
    if (rc != QPB_OK) return rc;
    if (cf.nw > 0) {
        QPB_CUDA(cudaMemcpyAsync(d_bins, n_ph_bins, sizeof(double) * cf.nw, cudaMemcpyHostToDevice, c->stream));
        rc = qpbk_broadcast_phonons(c, d_bins);
        QPB_CUDA(cudaStreamSynchronize(c->stream));
        if (rc != QPB_OK) return rc;
        rc = qpbk_uniform_setup(c, n_ph_bins, true);
        QPB_CUDA(cudaDeviceSynchronize());
        return rc;
    }
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_get_state(qpb_ctx *c, double *n, double *n_ph) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (n) {
        int rc = qpbk_gather_state(c, c->d_T1);
        if (rc != QPB_OK) return rc;
        QPB_CUDA(d2h_staged(n, c->d_T1, sizeof(double) * (size_t)cf.ne * cf.ncell, c->stream));
    }
    if (n_ph && cf.nw > 0) QPB_CUDA(d2h_staged(n_ph, c->d_P, sizeof(double) * (size_t)cf.nw * cf.ncell, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_get_integrated(qpb_ctx *c, double *out) {
    QPB_ENTER(c);
    if (!out) {
        qpb_set_error("qpb_get_integrated: null output");
        return QPB_E_INVALID;
    }
    int rc = qpbk_integrate(c);
    if (rc != QPB_OK) return rc;
    QPB_CUDA(cudaMemcpyAsync(out, c->d_integrated, sizeof(double) * c->cfg.ncell, cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_get_frames(qpb_ctx *c, double *frames) {
    QPB_ENTER(c);
    if (!frames || !c->have_geom) {
        qpb_set_error("qpb_get_frames: null output or no geometry");
        return QPB_E_INVALID;
    }
    int rc = qpbk_frames(c, c->d_T1);   // T1 is free between solves (it also stages qpb_get_state)
    if (rc != QPB_OK) return rc;
    QPB_CUDA(d2h_staged(frames, c->d_T1, sizeof(double) * (size_t)c->cfg.ne * c->ncd, c->stream));
    // no NaN may survive in a work array: 0 * NaN would leak into cells outside the mask
    QPB_CUDA(cudaMemsetAsync(c->d_T1, 0, sizeof(double) * (size_t)c->cfg.ne * c->ncd, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

// Snapshot now, download while the run goes on: the frames of a stored step are assembled into a buffer of their own
// on the context's stream (26 us at C2) and cross PCIe on a second stream, so the 5 ms of a 64 MiB download hide behind
// the following time steps.  qpb_frames_download may be called from another host thread than the one that steps.
extern "C" int qpb_frames_snapshot(qpb_ctx *c) {
    QPB_ENTER(c);
    if (!c->have_geom) {
        qpb_set_error("qpb_frames_snapshot: no geometry");
        return QPB_E_INVALID;
    }
    if (!c->d_snap) QPB_ALLOC(c->d_snap, (size_t)c->cfg.ne * c->ncd);
    if (!c->ev_snap) QPB_CUDA(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
    if (!c->copy_stream) QPB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    int rc = qpbk_frames(c, c->d_snap);
    if (rc != QPB_OK) return rc;
    QPB_CUDA(cudaEventRecord(c->ev_snap, c->stream));
    return QPB_OK;
}

extern "C" int qpb_frames_download(qpb_ctx *c, double *frames) {
    QPB_ENTER(c);
    if (!frames || !c->d_snap || !c->ev_snap || !c->copy_stream) {
        qpb_set_error("qpb_frames_download: null output or no snapshot taken");
        return QPB_E_INVALID;
    }
    QPB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    QPB_CUDA(d2h_staged(frames, c->d_snap, sizeof(double) * (size_t)c->cfg.ne * c->ncd, c->copy_stream));
    QPB_CUDA(cudaStreamSynchronize(c->copy_stream));
    return QPB_OK;
}

extern "C" int qpb_collide(qpb_ctx *c, double dt) {
    QPB_ENTER(c);
    if (!c->have_coll) {
        qpb_set_error("qpb_collide: collision tables were not uploaded");
        return QPB_E_INVALID;
    }
    if (!(dt > 0.0)) return QPB_OK;  // solver.py:1385
    return qpbk_collide(c, dt);
}

extern "C" int qpb_euler_step(qpb_ctx *c, int32_t kind, const double *K, const double *vec, double dt) {
    QPB_ENTER(c);
    if (kind < 1 || kind > 2 || !K || !vec || !c->have_geom) {
        qpb_set_error("qpb_euler_step: kind 1 (scattering) or 2 (recombination), non-null tables, geometry uploaded");
        return QPB_E_INVALID;
    }
    return qpbk_euler_step(c, kind, K, vec, dt);
}

extern "C" int qpb_set_exchange(qpb_ctx *c, int32_t nranks, void *const *peer_state, int64_t peer_ncd,
                                const int16_t *bin_owner, const int16_t *bin_row, const int32_t *cell_dense) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (nranks < 1 || nranks > 8 || !peer_state || !bin_owner || !bin_row || !cell_dense || peer_ncd <= 0) {
        qpb_set_error("qpb_set_exchange: bad arguments (1..8 ranks, non-null tables)");
        return QPB_E_INVALID;
    }
    if (!c->have_coll || !c->structured || !(cf.flags & (QPB_F_SCATTERING | QPB_F_RECOMBINATION)) ||
        (c->uniform_ph && (cf.flags & QPB_F_FREEZE_PHONONS))) {
        qpb_set_error("qpb_set_exchange: only the structured collision kernel carries the exchange "
                      "(upload the collision tables and the state first)");
        return QPB_E_INVALID;
    }
    c->x_route.assign(cf.ne, 0);
    for (int i = 0; i < cf.ne; ++i) {
        if (bin_owner[i] < 0 || bin_owner[i] >= nranks || bin_row[i] < 0 || bin_row[i] > 1023) {
            qpb_set_error("qpb_set_exchange: bin %d routed to rank %d row %d", i, bin_owner[i], bin_row[i]);
            return QPB_E_INVALID;
        }
        c->x_route[i] = (int16_t)((bin_owner[i] << 10) | bin_row[i]);
    }
    for (int q = 0; q < cf.ncell; ++q)
        if (cell_dense[q] < 0 || cell_dense[q] >= peer_ncd) {
            qpb_set_error("qpb_set_exchange: cell %d has dense index %d outside the peer grid", q, cell_dense[q]);
            return QPB_E_INVALID;
        }
    for (int r = 0; r < 8; ++r) c->x_peer[r] = r < nranks ? (double *)peer_state[r] : nullptr;
    for (int r = 0; r < nranks; ++r)
        if (!c->x_peer[r]) {
            qpb_set_error("qpb_set_exchange: null state pointer for rank %d", r);
            return QPB_E_INVALID;
        }
    dev_free(c->d_xdense);
    QPB_ALLOC(c->d_xdense, cf.ncell);
    QPB_CUDA(cudaMemcpy(c->d_xdense, cell_dense, sizeof(int32_t) * cf.ncell, cudaMemcpyHostToDevice));
    QPB_CUDA(cudaDeviceSynchronize());
    c->x_nranks = nranks;
    c->x_ncd = peer_ncd;
    return QPB_OK;
}

extern "C" int qpb_collide_exchange(qpb_ctx *c, double dt, int32_t mode) {
    QPB_ENTER(c);
    if (!c->have_coll || mode < 1 || mode > 2 || c->x_nranks <= 0) {
        qpb_set_error("qpb_collide_exchange: needs collision tables, qpb_set_exchange and mode 1 or 2");
        return QPB_E_INVALID;
    }
    if (!(dt > 0.0)) {
        qpb_set_error("qpb_collide_exchange: dt must be positive (the exchange rides on the collision kernel)");
        return QPB_E_INVALID;
    }
    return qpbk_collide(c, dt, mode);
}

extern "C" int qpb_ipc_export(qpb_ctx *c, int which, void *handle64) {
    QPB_ENTER(c);
    void *ptr = which == 0 ? (void *)c->d_S : (which == 1 ? (void *)c->d_P : nullptr);
    if (!ptr || !handle64) {
        qpb_set_error("qpb_ipc_export: nothing to export");
        return QPB_E_INVALID;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    QPB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, ptr));
    return QPB_OK;
}

extern "C" int qpb_ipc_open(int device, const void *handle64, void **ptr) {
    if (!handle64 || !ptr) return QPB_E_INVALID;
    QPB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    QPB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QPB_OK;
}

extern "C" int qpb_ipc_close(int device, void *ptr) {
    if (!ptr) return QPB_OK;
    QPB_CUDA(cudaSetDevice(device));
    QPB_CUDA(cudaIpcCloseMemHandle(ptr));
    return QPB_OK;
}

extern "C" int qpb_diffuse(qpb_ctx *c, int32_t slot) {
    QPB_ENTER(c);
    if (slot < 0 || slot > 1 || !c->slot[slot].ready) {
        qpb_set_error("qpb_diffuse: slot %d was not prepared", slot);
        return QPB_E_INVALID;
    }
    return qpbk_diffuse(c, c->slot[slot]);
}

extern "C" int qpb_pauli(qpb_ctx *c, qpb_pauli_rec *out) {
    QPB_ENTER(c);
    if (!c->d_rho || !out) {
        qpb_set_error("qpb_pauli: density of states not uploaded or null output");
        return QPB_E_INVALID;
    }
    int rc = qpbk_pauli(c, c->d_pauli);
    if (rc != QPB_OK) return rc;
    QPB_CUDA(cudaMemcpyAsync(out, c->d_pauli, sizeof(qpb_pauli_rec), cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_pauli_record(qpb_ctx *c, int32_t slot) {
    QPB_ENTER(c);
    if (!c->d_rho || slot < 0) {
        qpb_set_error("qpb_pauli_record: density of states not uploaded or negative slot");
        return QPB_E_INVALID;
    }
    if (slot >= c->pauli_cap) {   // grow, keeping the records already taken
        const int cap = std::max(2 * c->pauli_cap, slot + 1);
        qpb_pauli_rec *bigger = nullptr;
        QPB_ALLOC(bigger, (size_t)cap);
        QPB_CUDA(cudaMemcpyAsync(bigger, c->d_pauli, sizeof(qpb_pauli_rec) * c->pauli_cap, cudaMemcpyDeviceToDevice,
                                 c->stream));
        QPB_CUDA(cudaStreamSynchronize(c->stream));
        dev_free(c->d_pauli);
        c->d_pauli = bigger;
        c->pauli_cap = cap;
    }
    return qpbk_pauli(c, c->d_pauli + slot);   // stream ordered, no host synchronisation
}

extern "C" int qpb_pauli_fetch(qpb_ctx *c, int32_t count, qpb_pauli_rec *out) {
    QPB_ENTER(c);
    if (!out || count < 0 || count > c->pauli_cap) {
        qpb_set_error("qpb_pauli_fetch: bad count %d (capacity %d) or null output", count, c->pauli_cap);
        return QPB_E_INVALID;
    }
    if (count > 0)
        QPB_CUDA(cudaMemcpyAsync(out, c->d_pauli, sizeof(qpb_pauli_rec) * count, cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_advance(qpb_ctx *c, int32_t nsteps, double dt, int32_t slot, double t_start,
                           const qpb_generation *gen, qpb_pauli_rec *pauli_out) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (nsteps < 0 || !(dt > 0.0)) {
        qpb_set_error("qpb_advance: bad nsteps %d or dt %g", nsteps, dt);
        return QPB_E_INVALID;
    }
    const bool diff = cf.flags & QPB_F_DIFFUSION;
    const bool coll = cf.flags & (QPB_F_SCATTERING | QPB_F_RECOMBINATION);
    const bool pauli = (cf.flags & QPB_F_PAULI) && c->d_rho;
    if (diff && (slot < 0 || slot > 1 || !c->slot[slot].ready)) {
        qpb_set_error("qpb_advance: diffusion slot %d was not prepared", slot);
        return QPB_E_INVALID;
    }
    if (coll && !c->have_coll) {
        qpb_set_error("qpb_advance: collision tables were not uploaded");
        return QPB_E_INVALID;
    }
    const int gmode = gen ? gen->mode : QPB_GEN_NONE;
    if (gmode == QPB_GEN_ARRAY) {
        if (!gen->array) {
            qpb_set_error("qpb_advance: array generation needs a non-null array");
            return QPB_E_INVALID;
        }
        const size_t n = (size_t)cf.ne * cf.ncell;
        if (!c->d_gen) QPB_ALLOC(c->d_gen, n);
        c->gen_resident = false;
        QPB_CUDA(cudaMemcpyAsync(c->d_gen, gen->array, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        c->gen_resident = true;
    } else if (gmode == QPB_GEN_RESIDENT) {
        if (!c->d_gen || !c->gen_resident) {
            qpb_set_error("qpb_advance: no generation array is resident (run a QPB_GEN_ARRAY batch first)");
            return QPB_E_INVALID;
        }
    } else if (gmode == QPB_GEN_PROGRAM) {
        if (!c->d_genprog || c->gen_nops <= 0) {
            qpb_set_error("qpb_advance: no generation program was uploaded (qpb_upload_generation_program)");
            return QPB_E_INVALID;
        }
        QPB_CUDA(cudaMemsetAsync(c->d_gen_flag, 0, sizeof(int), c->stream));
    } else if (gmode < QPB_GEN_NONE || gmode > QPB_GEN_PROGRAM) {
        qpb_set_error("qpb_advance: unknown generation mode %d", gmode);
        return QPB_E_INVALID;
    }
    if (pauli && nsteps > c->pauli_cap) {
        dev_free(c->d_pauli);
        c->pauli_cap = nsteps;
        QPB_ALLOC(c->d_pauli, (size_t)c->pauli_cap);
    }
    double t = t_start;
    int rc;
    cudaEvent_t ea = nullptr, eb = nullptr;
    QPB_CUDA(cudaEventCreate(&ea));
    QPB_CUDA(cudaEventCreate(&eb));
    struct EvGuard {
        cudaEvent_t a, b;
        ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); }
    } guard{ea, eb};
    QPB_CUDA(cudaEventRecord(ea, c->stream));
    for (int s = 0; s < nsteps; ++s) {
        if (gmode == QPB_GEN_CONSTANT) {
            if ((rc = qpbk_add_generation(c, dt, gen->rate, nullptr)) != QPB_OK) return rc;
        } else if (gmode == QPB_GEN_PULSE) {
            if (gen->pulse_start <= t && t < gen->pulse_start + gen->pulse_duration)  // solver.py:914
                if ((rc = qpbk_add_generation(c, dt, gen->rate, nullptr)) != QPB_OK) return rc;
        } else if (gmode == QPB_GEN_ARRAY || gmode == QPB_GEN_RESIDENT) {
            if ((rc = qpbk_add_generation(c, dt, 0.0, c->d_gen)) != QPB_OK) return rc;
        } else if (gmode == QPB_GEN_PROGRAM) {
            if ((rc = qpbk_generation_program(c, dt, t, nullptr)) != QPB_OK) return rc;
        }
        if (coll && diff) {  // solver.py:1469-1472
            if ((rc = qpbk_collide(c, 0.5 * dt)) != QPB_OK) return rc;
            if ((rc = qpbk_diffuse(c, c->slot[slot])) != QPB_OK) return rc;
            if ((rc = qpbk_collide(c, 0.5 * dt)) != QPB_OK) return rc;
        } else {             // solver.py:1474-1475
            if (coll && (rc = qpbk_collide(c, dt)) != QPB_OK) return rc;
            if (diff && (rc = qpbk_diffuse(c, c->slot[slot])) != QPB_OK) return rc;
        }
        if (pauli && (rc = qpbk_pauli(c, c->d_pauli + s)) != QPB_OK) return rc;
        t += dt;
        c->diag.steps_done++;
    }
    QPB_CUDA(cudaEventRecord(eb, c->stream));
    if (pauli && pauli_out && nsteps > 0)
        QPB_CUDA(cudaMemcpyAsync(pauli_out, c->d_pauli, sizeof(qpb_pauli_rec) * nsteps, cudaMemcpyDeviceToHost,
                                 c->stream));
    int gflag = 0;
    if (gmode == QPB_GEN_PROGRAM)
        QPB_CUDA(cudaMemcpyAsync(&gflag, c->d_gen_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    QPB_CUDA(cudaEventElapsedTime(&ms, ea, eb));
    c->diag.last_advance_ms = ms;
    if (gflag & 1) {   // the reference's checks of evaluate_external_generation (solver.py:954-962), same messages
        qpb_set_error("External generation mode 'custom' produced non-finite values.");
        return QPB_E_INVALID;
    }
    if (gflag & 2) {
        qpb_set_error("External generation mode 'custom' produced negative values. Generation rates must be non-negative.");
        return QPB_E_INVALID;
    }
    return QPB_OK;
}

extern "C" int qpb_upload_generation_program(qpb_ctx *c, const qpb_gen_op *ops, int32_t nops, const double *E_bins,
                                             const double *cell_x, const double *cell_y) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (!ops || nops <= 0 || nops > QPB_GEN_MAX_OPS || !E_bins || !cell_x || !cell_y) {
        qpb_set_error("qpb_upload_generation_program: null argument or a program of %d operators (1..%d)", nops,
                      QPB_GEN_MAX_OPS);
        return QPB_E_INVALID;
    }
    // the program must leave exactly one value and never outgrow the stack: checked here, not per thread
    int depth = 0, peak = 0;
    for (int k = 0; k < nops; ++k) {
        const int op = ops[k].op;
        if (op < 0 || op >= QPB_OP_COUNT) {
            qpb_set_error("qpb_upload_generation_program: unknown operator %d at %d", op, k);
            return QPB_E_INVALID;
        }
        const int pops = op <= QPB_OP_T ? 0 : op == QPB_OP_SELECT ? 3
                         : (op >= QPB_OP_ABS || op == QPB_OP_NEG || op == QPB_OP_NOT || op == QPB_OP_TRUTH) ? 1 : 2;
        if (depth < pops) {
            qpb_set_error("qpb_upload_generation_program: operator %d at %d finds %d values on the stack", op, k, depth);
            return QPB_E_INVALID;
        }
        depth += 1 - pops;
        peak = std::max(peak, depth);
    }
    if (depth != 1 || peak > QPB_GEN_MAX_STACK) {
        qpb_set_error("qpb_upload_generation_program: the program leaves %d values (must be 1), stack depth %d (max %d)",
                      depth, peak, QPB_GEN_MAX_STACK);
        return QPB_E_INVALID;
    }
    dev_free(c->d_genprog);
    QPB_CUDA(qpb_dev_malloc(&c->d_genprog, sizeof(qpb_gen_op) * (size_t)nops));
    if (!c->d_gen_E) QPB_ALLOC(c->d_gen_E, (size_t)cf.ne);
    if (!c->d_gen_x) QPB_ALLOC(c->d_gen_x, (size_t)cf.ncell);
    if (!c->d_gen_y) QPB_ALLOC(c->d_gen_y, (size_t)cf.ncell);
    if (!c->d_gen_flag) QPB_CUDA(qpb_dev_malloc((void **)&c->d_gen_flag, sizeof(int)));
    QPB_CUDA(cudaMemcpyAsync(c->d_genprog, ops, sizeof(qpb_gen_op) * (size_t)nops, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(c->d_gen_E, E_bins, sizeof(double) * cf.ne, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(c->d_gen_x, cell_x, sizeof(double) * cf.ncell, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemcpyAsync(c->d_gen_y, cell_y, sizeof(double) * cf.ncell, cudaMemcpyHostToDevice, c->stream));
    QPB_CUDA(cudaMemsetAsync(c->d_gen_flag, 0, sizeof(int), c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    c->gen_nops = nops;
    return QPB_OK;
}

extern "C" int qpb_eval_generation_program(qpb_ctx *c, double t, double *out) {
    QPB_ENTER(c);
    const auto &cf = c->cfg;
    if (!out || !c->d_genprog || c->gen_nops <= 0) {
        qpb_set_error("qpb_eval_generation_program: no program uploaded or null output");
        return QPB_E_INVALID;
    }
    const size_t n = (size_t)cf.ne * cf.ncell;
    if (!c->d_gen) QPB_ALLOC(c->d_gen, n);
    c->gen_resident = false;
    int rc = qpbk_generation_program(c, 0.0, t, c->d_gen);
    if (rc != QPB_OK) return rc;
    QPB_CUDA(cudaMemcpyAsync(out, c->d_gen, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_get_diag(qpb_ctx *c, qpb_diag *out) {
    if (!c || !out) {
        qpb_set_error("qpb_get_diag: null argument");
        return QPB_E_INVALID;
    }
    *out = c->diag;
    return QPB_OK;
}

extern "C" int qpb_synchronize(qpb_ctx *c) {
    QPB_ENTER(c);
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    return QPB_OK;
}

extern "C" int qpb_set_stream(qpb_ctx *c, void *cuda_stream) {
    QPB_ENTER(c);
    QPB_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return QPB_OK;
}

extern "C" int qpb_enable_timers(qpb_ctx *c, int on) {
    if (!c) return QPB_E_INVALID;
    c->timers_on = on != 0;
    return QPB_OK;
}

extern "C" int qpb_reset_timers(qpb_ctx *c) {
    if (!c) return QPB_E_INVALID;
    for (auto &t : c->timer) t = Timer();
    return QPB_OK;
}

extern "C" int qpb_get_timer(qpb_ctx *c, int which, double *ms, int64_t *launches) {
    if (!c || which < 0 || which > 2) return QPB_E_INVALID;
    if (ms) *ms = c->timer[which].ms;
    if (launches) *launches = c->timer[which].launches;
    return QPB_OK;
}

extern "C" int qpb_device_ptr(qpb_ctx *c, int which, void **ptr, int64_t *bytes) {
    if (!c || !ptr) return QPB_E_INVALID;
    if (which == 0) {
        *ptr = c->d_S;
        if (bytes) *bytes = (int64_t)sizeof(double) * c->cfg.ne * c->ncd;
    } else if (which == 1) {
        *ptr = c->d_P;
        if (bytes) *bytes = (int64_t)sizeof(double) * c->cfg.nw * c->cfg.ncell;
    } else
        return QPB_E_INVALID;
    return QPB_OK;
}
