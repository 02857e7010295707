// gb200_api.cu -- the C ABI of include/gradus_b200.h: context, validation, host-side problem
// set-up (observer LNRF constants, special radii), device buffer management and launches.
// No CPU implementation of the path exists in this library: every compute entry point runs
// the sm_100a kernels of gb200_trace.cu or fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types and enums only: the library is opened at run time (see gb200_comm_init)
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>
#include <functional>
#include <thread>
#include <vector>
#include "gb200_device.cuh"
#include "gb200_internal.h"

// ---------------------------------------------------------------- errors
static thread_local std::string g_last_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct gb200_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr; // the context's own stream
    cudaStream_t cur = nullptr;    // stream of the call in flight (own stream unless the caller passed one)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::string err;
    gb200_stats stats{};
    std::vector<DevBuf> pool;
    unsigned long long* d_queue = nullptr; // [0] queue, [1..3] counters
    double fp64_peak = 0.0;
    std::vector<cudaStream_t> pool_streams; // gb200_trace_batch, gb200_render_batch, pipelined gb200_render
    std::vector<cudaEvent_t> chunk_events;  // pipelined gb200_render: one per chunk
    void* stage = nullptr;                  // pinned host staging for gb200_trace_batch
    size_t stage_cap = 0;
    // An asynchronous call (gb200_render_device / gb200_lineprofile_device with async != 0) returns while its kernels still
    // use the context's work queue and pool buffers: it records `inflight` at its end, and every entry point makes the
    // stream it works on wait for that event before touching the queue or the pool.
    cudaEvent_t inflight = nullptr;
    bool have_inflight = false;
    // cross-section table of GB200_GEOMETRY_THICK_TABLE (gb200_set_cross_section)
    double* d_cs = nullptr; // [rho (n) | height (n)]
    int cs_n = 0;
    double cs_max = 0.0;
};

static int fail(gb200_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}
#define CU(ctx, call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) return fail(ctx, GB200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// ---------------------------------------------------------------- host-only dual number (d/dr), for the ISCO root find
struct D1 {
    double v, d;
    GB_HD D1(double v_ = 0.0, double d_ = 0.0) : v(v_), d(d_) {}
    GB_HD friend D1 operator+(const D1& a, const D1& b) { return D1(a.v + b.v, a.d + b.d); }
    GB_HD friend D1 operator-(const D1& a, const D1& b) { return D1(a.v - b.v, a.d - b.d); }
    GB_HD friend D1 operator-(const D1& a) { return D1(-a.v, -a.d); }
    GB_HD friend D1 operator*(const D1& a, const D1& b) { return D1(a.v * b.v, a.d * b.v + a.v * b.d); }
    GB_HD friend D1 operator/(const D1& a, const D1& b) { double q = a.v / b.v; return D1(q, (a.d - q * b.d) / b.v); }
};
GB_HD static inline D1 gb_rcp(const D1& a) { return D1(1.0) / a; }
static inline D1 dsqrt(const D1& a) { double s = std::sqrt(a.v); return D1(s, a.d / (2.0 * s)); }
static inline double dsqrt(double a) { return std::sqrt(a); }
static inline double val(const D1& a) { return a.v; }
static inline double val(double a) { return a; }

// E = -u_t of the equatorial circular orbit at radius r (CircularOrbits.energy, src/orbits/circular-orbits.jl:11-52)
template <class S>
static S circular_energy_host(int kind, const double* mp, S r) {
    S g[5], dr[5], dth[5];
    metric_jacobian_kind<S>(kind, mp, r, S(1.0), S(0.0), g, dr, dth);
    const S D = g[0] * g[3] - g[4] * g[4];
    const S gitt = g[3] / D, giphph = g[0] / D, gitph = -g[4] / D;
    const S Om = -(dr[4] - dsqrt(dr[4] * dr[4] - dr[0] * dr[3])) / dr[3];
    const S A = -(Om * gitt - gitph);
    const S B = (Om * gitph - giphph);
    const S denom = B * B * gitt + 2.0 * A * B * gitph + A * A * giphph;
    const double sg = val(denom) > 0 ? 1.0 : (val(denom) < 0 ? -1.0 : 0.0);
    const S ad = val(denom) < 0 ? -denom : denom;
    const S d = -sg * dsqrt(S(1.0) / ad);
    return -(B * d);
}

static double kerr_isco_host(double M, double a) { // Bardeen et al. (1972), kerr-metric-first-order.jl:297-337
    const double x = a / M;
    const double Z1 = 1 + std::cbrt(1 - x * x) * (std::cbrt(1 + x) + std::cbrt(1 - x));
    const double Z2 = std::sqrt(3 * x * x + Z1 * Z1);
    const double s = std::sqrt((3 - Z1) * (3 + Z1 + 2 * Z2));
    return a > 0.0 ? M * (3 + Z2 - s) : M * (3 + Z2 + s);
}

// isco(::AbstractStaticAxisSymmetric): bracket from find_isco_bounds, then the root of dE/dr (special-radii.jl:14-60)
static int generic_isco_host(int kind, const double* mp, double* out) {
    double lower = 0, upper = 0;
    for (long n = 0;; ++n) {
        const double r = 100.0 - 0.005 * (double)n;
        if (r < 1.0) break;
        const double en = circular_energy_host<double>(kind, mp, r);
        if (std::fabs(en) > 1.0) { lower = r; upper = 100.0; break; }
    }
    if (lower == upper) return GB200_ERR_INVALID_ARGUMENT;
    auto dE = [&](double r) { return circular_energy_host<D1>(kind, mp, D1(r, 1.0)).d; };
    double lo = lower, hi = upper, flo = dE(lo);
    for (int it = 0; it < 200; ++it) {
        const double mid = lo + (hi - lo) / 2;
        if (!(mid > lo && mid < hi)) break;
        const double fm = dE(mid);
        if (fm == 0) { lo = hi = mid; break; }
        if ((fm > 0) == (flo > 0)) { lo = mid; flo = fm; } else hi = mid;
    }
    *out = lo + (hi - lo) / 2;
    return GB200_OK;
}

// Start of every entry point that uses the context's queue / pool on `stream`: order it after an earlier asynchronous call.
static int begin_call(gb200_ctx* ctx, cudaStream_t stream) {
    CU(ctx, cudaSetDevice(ctx->device));
    ctx->stats = gb200_stats{};
    ctx->cur = stream;
    if (ctx->have_inflight) CU(ctx, cudaStreamWaitEvent(stream, ctx->inflight, 0));
    return GB200_OK;
}
// End of an asynchronous call: later calls (on any stream) wait for the work enqueued so far.
static int end_async_call(gb200_ctx* ctx, cudaStream_t stream) {
    CU(ctx, cudaEventRecord(ctx->inflight, stream));
    ctx->have_inflight = true;
    return GB200_OK;
}

// ---------------------------------------------------------------- validation
static int validate(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic) {
    if (!p || !ic) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null problem or ic");
    if (p->metric_kind < 0 || p->metric_kind >= GB200_METRIC_COUNT)
        return fail(ctx, GB200_ERR_UNSUPPORTED, "metric kind %d has no closed-form right-hand side in this library", p->metric_kind);
    const double M = p->metric_params[0], a = p->metric_params[1];
    if (p->metric_kind == GB200_METRIC_MORRIS_THORNE) { // metric_params[0] is the throat size b, there is no mass or spin
        if (!(M == M)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "Morris-Thorne throat size b is NaN");
    } else if (!(M > 0) || !(std::fabs(a) <= M)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "need M > 0 and |a| <= M (M=%g a=%g)", M, a);
    if (p->metric_kind == GB200_METRIC_BUMBLEBEE && (!(p->metric_params[2] > -1.0) || std::fabs(a) > 0.3))
        return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "Bumblebee metric needs l > -1 and |a| <= 0.3 (bumblebee-ad.jl:33-40)");
    if (p->metric_kind == GB200_METRIC_DILATON_AXION && p->metric_params[2] != 0.0 && (a == 0.0 || p->metric_params[3] == 0.0))
        return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "Dilaton-Axion metric with beta != 0 needs a != 0 and b != 0 (beta / a, beta / b; dilaton-axion-ad.jl:24-26)");
    if (p->metric_kind == GB200_METRIC_KERR_NEWMAN && a * a + p->metric_params[2] * p->metric_params[2] > M * M)
        return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "Kerr-Newman metric needs a^2 + Q^2 <= M^2 (kerr-newman-ad.jl:50-52)");
    if (p->geometry_kind < GB200_GEOMETRY_NONE || p->geometry_kind > GB200_GEOMETRY_THICK_TABLE)
        return fail(ctx, GB200_ERR_UNSUPPORTED, "geometry kind %d is outside the hot-path scope", p->geometry_kind);
    if (p->geometry_kind == GB200_GEOMETRY_THIN_DISC && !(p->geometry_params[0] <= p->geometry_params[1]))
        return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "ThinDisc needs inner_radius <= outer_radius");
    if (p->callback_kind != GB200_CALLBACK_NONE && p->callback_kind != GB200_CALLBACK_UPPER_HEMISPHERE)
        return fail(ctx, GB200_ERR_UNSUPPORTED, "callback kind %d cannot run on the device", p->callback_kind);
    if (p->pow_mode != GB200_POW_EXACT && p->pow_mode != GB200_POW_FAST32) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad pow_mode");
    if (!(p->mu == p->mu) && ic->kind != GB200_IC_EXPLICIT) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "mu = NaN (keep v^t as given) needs explicit initial conditions");
    if (!(p->lambda_max > p->lambda_min)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "need lambda_max > lambda_min");
    if (!(p->abstol > 0) || !(p->reltol > 0)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "abstol and reltol must be positive");
    if (!(p->chart_outer > p->chart_inner)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "chart outer radius must exceed inner radius");
    if (ic->kind == GB200_IC_RENDER_GRID) {
        if (ic->width < 1 || ic->height < 1) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "image size must be positive");
        if (!(ic->lo0 <= ic->hi0) || !(ic->lo1 <= ic->hi1)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "alpha / beta limits must be sorted");
        if (ic->n != ic->width * ic->height) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "ic.n != width*height");
    } else if (ic->kind == GB200_IC_POLAR_PLANE) {
        if (ic->width < 2 || ic->height < 1) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "PolarPlane needs Nr >= 2, Ntheta >= 1");
        if (ic->grid_kind < GB200_GRID_LINEAR || ic->grid_kind > GB200_GRID_INVERSE) return fail(ctx, GB200_ERR_UNSUPPORTED, "grid kind %d", ic->grid_kind);
        if (!(ic->lo0 > 0) || !(ic->hi0 > ic->lo0)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "PolarPlane needs 0 < r_min < r_max");
        if (ic->n != ic->width * ic->height) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "ic.n != Nr*Ntheta");
    } else if (ic->kind == GB200_IC_CARTESIAN_PLANE) {
        if (ic->width < 4 || ic->height < 4) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "CartesianPlane needs Nx, Ny >= 4");
        if (ic->grid_kind < GB200_GRID_LINEAR || ic->grid_kind > GB200_GRID_INVERSE) return fail(ctx, GB200_ERR_UNSUPPORTED, "grid kind %d", ic->grid_kind);
        if (!(ic->hi0 > ic->lo0) || !(ic->hi1 > ic->lo1)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "CartesianPlane needs x_min < x_max, y_min < y_max");
        if (ic->grid_kind != GB200_GRID_LINEAR && (!(ic->lo0 > 0) || !(ic->lo1 > 0))) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "geometric / inverse grids need positive minima");
        if (ic->n != (2 * (ic->height / 2) - 1) * (2 * (ic->width / 2) - 1)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "ic.n != (2(Ny/2)-1)(2(Nx/2)-1)");
    } else if (ic->kind == GB200_IC_IMPACT_PARAMETERS) {
        if (ic->n < 1 || !ic->x[0] || !ic->x[1]) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "impact-parameter IC needs n >= 1 and alpha / beta arrays");
    } else if (ic->kind == GB200_IC_EXPLICIT) {
        if (ic->n < 1) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "explicit IC needs n >= 1");
        for (int k = 0; k < 4; ++k)
            if (!ic->x[k] || !ic->v[k]) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "explicit IC pointers must be non-null");
    } else return fail(ctx, GB200_ERR_UNSUPPORTED, "ic kind %d", ic->kind);
    return GB200_OK;
}

static int validate_range(gb200_ctx* ctx, const gb200_ic* ic, const gb200_range* rg) {
    if (!rg) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null range");
    if (rg->count < 0 || rg->first < 0 || rg->stride < 1 || rg->block < 0) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad range");
    const int64_t blk = rg->block > 0 ? rg->block : 1;
    if (rg->count > 0) {
        const int64_t nlast = rg->count - 1;
        const int64_t last = rg->first + (nlast / blk) * (rg->stride * blk) + nlast % blk;
        if (last >= ic->n) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "range exceeds the %lld rays of the IC", (long long)ic->n);
    }
    return GB200_OK;
}

// ---------------------------------------------------------------- problem -> kernel parameters (host arithmetic)
static void dd_step(double lo, double hi, int64_t n, double* shi, double* slo) { // (hi-lo)/(n-1) as double-double
    if (n <= 1) { *shi = 0; *slo = 0; return; }
    const long double s = ((long double)hi - (long double)lo) / (long double)(n - 1);
    *shi = (double)s;
    *slo = (double)(s - (long double)*shi);
}

static void fill_params(const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, GbParams& P) {
    memset(&P, 0, sizeof P);
    P.metric_kind = p->metric_kind;
    P.M = p->metric_params[0]; P.a = p->metric_params[1]; P.eps3 = p->metric_params[2];
    P.a2 = P.a * P.a; P.twoM = 2.0 * P.M; P.jp_e = P.eps3 * P.M * P.M * P.M;
    for (int k = 0; k < 8; ++k) P.mp[k] = p->metric_params[k];
    P.lam0 = p->lambda_min; P.lam1 = p->lambda_max; P.abstol = p->abstol; P.reltol = p->reltol;
    P.dtmax = p->dtmax > 0 ? p->dtmax : (p->lambda_max - p->lambda_min);
    P.mu = p->mu;
    P.maxiters = p->maxiters > 0 ? p->maxiters : 1000000;
    P.pow_mode = p->pow_mode;
    P.geometry_kind = p->geometry_kind;
    P.gp0 = p->geometry_params[0]; P.gp1 = p->geometry_params[1]; P.gp2 = p->geometry_params[2];
    P.gtol = p->gtol;
    P.chart_inner = p->chart_inner; P.chart_outer = p->chart_outer;
    P.callback_kind = p->callback_kind; P.callback_delta = p->callback_delta;
    P.ic_kind = ic->kind; P.grid_kind = ic->grid_kind;
    P.width = ic->width; P.height = ic->height;
    for (int k = 0; k < 4; ++k) P.xo[k] = p->observer[k];
    if (ic->kind == GB200_IC_RENDER_GRID) {
        P.lo0 = ic->lo0; dd_step(ic->lo0, ic->hi0, ic->width, &P.step0_hi, &P.step0_lo);
        P.lo1 = ic->lo1; dd_step(ic->lo1, ic->hi1, ic->height, &P.step1_hi, &P.step1_lo);
    } else if (ic->kind == GB200_IC_POLAR_PLANE) {
        P.lo0 = ic->lo0; dd_step(ic->lo0, ic->hi0, ic->width, &P.step0_hi, &P.step0_lo);
        P.geoK = std::pow(ic->hi0 / ic->lo0, 1.0 / (double)(ic->width - 1));
        P.inv_lo_hi = 1.0 / ic->hi0; dd_step(1.0 / ic->hi0, 1.0 / ic->lo0, ic->width, &P.inv_step_hi, &P.inv_step_lo);
        const double dth = (ic->hi1 - ic->lo1) / (double)ic->height; // planes.jl:95-96
        P.lo1 = ic->lo1; dd_step(ic->lo1, ic->hi1 - dth, ic->height, &P.step1_hi, &P.step1_lo);
    }
    if (ic->kind == GB200_IC_CARTESIAN_PLANE) {
        P.n0 = ic->width / 2; P.n1 = ic->height / 2;
        P.lo0 = ic->lo0; dd_step(ic->lo0, ic->hi0, P.n0, &P.step0_hi, &P.step0_lo);
        P.lo1 = ic->lo1; dd_step(ic->lo1, ic->hi1, P.n1, &P.step1_hi, &P.step1_lo);
        if (ic->grid_kind != GB200_GRID_LINEAR) {
            P.geoK = std::pow(ic->hi0 / ic->lo0, 1.0 / (double)(P.n0 - 1));
            P.geoK1 = std::pow(ic->hi1 / ic->lo1, 1.0 / (double)(P.n1 - 1));
            P.inv_lo_hi = 1.0 / ic->hi0; dd_step(1.0 / ic->hi0, 1.0 / ic->lo0, P.n0, &P.inv_step_hi, &P.inv_step_lo);
            P.inv1_lo_hi = 1.0 / ic->hi1; dd_step(1.0 / ic->hi1, 1.0 / ic->lo1, P.n1, &P.inv1_step_hi, &P.inv1_step_lo);
        }
    }
    if (ic->kind != GB200_IC_EXPLICIT) {
        // closed-form LNRF (ZAMO) co-basis at the observer, generic for static axisymmetric g:
        //   e^(t) = N dt, e^(r) = sqrt(g_rr) dr, e^(th) = sqrt(g_thth) dth, e^(ph) = sqrt(g_phph)(dph - w dt)
        double g[5];
        metric_components_rt(P, P.xo[1], P.xo[2], g);
        for (int k = 0; k < 5; ++k) P.go[k] = g[k];
        const double D = g[0] * g[3] - g[4] * g[4];
        const double gitt = g[3] / D, giphph = g[0] / D, gitph = -g[4] / D;
        const double N = std::sqrt(-1.0 / gitt);
        const double om = -g[4] / g[3];
        const double sph = std::sqrt(g[3]);
        P.c_r = 1.0 / std::sqrt(g[1]);
        P.c_th = 1.0 / std::sqrt(g[2]);
        P.c_ph0 = gitph * N;
        P.c_ph1 = (giphph - gitph * om) * sph;
    }
    P.first = rg->first; P.count = rg->count; P.stride = rg->stride;
    P.block = rg->block > 0 ? rg->block : 1;
    // 2-D tiled work order when the range is made of whole strips of GB_TILE_C image columns (or theta-rows of the plane)
    P.tile_h = 0;
    if (ic->kind != GB200_IC_EXPLICIT && ic->kind != GB200_IC_IMPACT_PARAMETERS && !getenv("GB200_NO_TILING")) {
        const int64_t h = (ic->kind == GB200_IC_RENDER_GRID) ? ic->height
                        : (ic->kind == GB200_IC_CARTESIAN_PLANE) ? (2 * (ic->height / 2) - 1) : ic->width; // fastest-varying extent of the ray index
        const int64_t strip = GB_TILE_C * h;
        const bool contiguous = (P.stride == 1) || (P.block % strip == 0);
        if (h % GB_TILE_R == 0 && contiguous && P.first % strip == 0 && P.count % strip == 0) P.tile_h = h;
    }
}

// ---------------------------------------------------------------- device memory
static int pool_get(gb200_ctx* ctx, size_t slot, size_t bytes, void** out) {
    if (ctx->pool.size() <= slot) ctx->pool.resize(slot + 1);
    DevBuf& b = ctx->pool[slot];
    if (b.cap < bytes) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr; b.cap = 0;
        cudaError_t e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) return fail(ctx, GB200_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        b.cap = bytes;
    }
    *out = b.p;
    return GB200_OK;
}
// pool slots
enum { SL_STATUS = 0, SL_LAMBDA, SL_X0, SL_V0 = SL_X0 + 4, SL_XI0 = SL_V0 + 4, SL_VI0 = SL_XI0 + 4, SL_NACC = SL_VI0 + 4, SL_NREJ, SL_FLAGS,
       SL_IMG0, SL_G = SL_IMG0 + GB_MAX_PF, SL_F, SL_EX0, SL_EV0 = SL_EX0 + 4, SL_BINS = SL_EV0 + 4, SL_PARTIAL, SL_FLUX, SL_EMR, SL_EME,
       SL_PL0, SL_SCRATCH = SL_PL0 + 4, SL_BATCH, SL_BATCH_QUEUE, SL_COUNT };

static int upload(gb200_ctx* ctx, size_t slot, const void* host, size_t bytes, const void** dev) {
    void* d = nullptr;
    int rc = pool_get(ctx, slot, bytes, &d);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->cur));
    *dev = d;
    return GB200_OK;
}

static int upload_ic_and_tables(gb200_ctx* ctx, const gb200_ic* ic, const gb200_plunging_table* pl, const gb200_emissivity* em, GbParams& P) {
    if (ic->kind == GB200_IC_IMPACT_PARAMETERS) {
        for (int k = 0; k < 3; ++k) {
            P.ex[k] = nullptr;
            if (k == 2 && !ic->x[2]) continue; // optional per-ray datum-plane heights
            const void* d;
            int rc = upload(ctx, SL_EX0 + k, ic->x[k], sizeof(double) * ic->n, &d); if (rc) return rc;
            P.ex[k] = (const double*)d;
        }
    }
    if (ic->kind == GB200_IC_EXPLICIT) {
        for (int k = 0; k < 4; ++k) {
            const void* d;
            int rc = upload(ctx, SL_EX0 + k, ic->x[k], sizeof(double) * ic->n, &d); if (rc) return rc;
            P.ex[k] = (const double*)d;
            rc = upload(ctx, SL_EV0 + k, ic->v[k], sizeof(double) * ic->n, &d); if (rc) return rc;
            P.ev[k] = (const double*)d;
        }
    }
    if (pl && pl->n >= 2) {
        const double* src[4] = {pl->r, pl->ut, pl->ur, pl->uphi};
        const double* dst[4];
        for (int k = 0; k < 4; ++k) {
            const void* d;
            int rc = upload(ctx, SL_PL0 + k, src[k], sizeof(double) * pl->n, &d); if (rc) return rc;
            dst[k] = (const double*)d;
        }
        P.pl_n = pl->n; P.pl_r = dst[0]; P.pl_ut = dst[1]; P.pl_ur = dst[2]; P.pl_uphi = dst[3];
    }
    if (em) {
        P.emis_kind = em->kind; P.emis_index = em->index; P.emis_n = em->n;
        if (em->kind == GB200_EMISSIVITY_TABLE) {
            if (em->n < 2 || !em->r || !em->eps) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "emissivity table needs n >= 2");
            const void* d;
            int rc = upload(ctx, SL_EMR, em->r, sizeof(double) * em->n, &d); if (rc) return rc;
            P.emis_r = (const double*)d;
            rc = upload(ctx, SL_EME, em->eps, sizeof(double) * em->n, &d); if (rc) return rc;
            P.emis_eps = (const double*)d;
        } else if (em->kind != GB200_EMISSIVITY_POWERLAW) return fail(ctx, GB200_ERR_UNSUPPORTED, "emissivity kind %d", em->kind);
    }
    return GB200_OK;
}

// GB200_GEOMETRY_THICK_TABLE: point the launch at the context's cross-section table
static int bind_geometry(gb200_ctx* ctx, GbParams& P) {
    if (P.geometry_kind != GB200_GEOMETRY_THICK_TABLE) return GB200_OK;
    if (ctx->cs_n < 2) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "GB200_GEOMETRY_THICK_TABLE needs a table: call gb200_set_cross_section first");
    P.cs_n = ctx->cs_n; P.cs_rho = ctx->d_cs; P.cs_h = ctx->d_cs + ctx->cs_n;
    P.gp0 = ctx->cs_max;
    return GB200_OK;
}

static bool needs_isco(const int32_t* pfs, int npf, bool lineprofile) {
    if (lineprofile) return true;
    for (int k = 0; k < npf; ++k) if (pfs[k] == GB200_PF_REDSHIFT) return true;
    return false;
}

static int set_isco(gb200_ctx* ctx, const gb200_problem* p, GbParams& P) {
    double r = 0;
    int rc = gb200_isco(p->metric_kind, p->metric_params, &r);
    if (rc) return fail(ctx, rc, "no ISCO found for this metric (is it a naked singularity?)");
    P.r_isco = r;
    return GB200_OK;
}

// run the trace kernel on `stream` (queue reset + launch), timing it with events when `timed`
static int run_trace(gb200_ctx* ctx, GbParams& P, cudaStream_t stream, bool timed) {
    P.queue = ctx->d_queue;
    P.counters = ctx->d_queue + 1;
    CU(ctx, cudaMemsetAsync(ctx->d_queue, 0, 4 * sizeof(unsigned long long), stream));
    if (timed) CU(ctx, cudaEventRecord(ctx->ev1, stream));
    int blocks = 0;
    if (P.count > 0) CU(ctx, gb200_launch_trace(P, ctx->sm_count, stream, &blocks));
    if (timed) CU(ctx, cudaEventRecord(ctx->ev2, stream));
    ctx->stats.launches += (P.count > 0) ? 1 : 0;
    return GB200_OK;
}

static int finish_stats(gb200_ctx* ctx, int64_t rays) {
    unsigned long long c[4];
    CU(ctx, cudaMemcpyAsync(c, ctx->d_queue, sizeof c, cudaMemcpyDeviceToHost, ctx->cur));
    CU(ctx, cudaEventRecord(ctx->ev3, ctx->cur));
    CU(ctx, cudaStreamSynchronize(ctx->cur));
    float k = 0, tot = 0;
    cudaEventElapsedTime(&k, ctx->ev1, ctx->ev2);
    cudaEventElapsedTime(&tot, ctx->ev0, ctx->ev3);
    ctx->stats.kernel_ms = k; ctx->stats.total_ms = tot; ctx->stats.rays = rays;
    ctx->stats.steps_accepted = (int64_t)c[1]; ctx->stats.steps_rejected = (int64_t)c[2]; ctx->stats.flagged = (int64_t)c[3];
    return GB200_OK;
}

// ---------------------------------------------------------------- pipelined host-output launches
// Large host-output calls are pipelined: the range is cut into chunks of whole tile strips, the chunk kernels alternate
// between two streams (the next chunk's CTAs take over the SMs as the previous chunk drains, so there is no idle tail
// between them) and every chunk's results travel to the host while the following chunks compute.
static int64_t pipeline_unit(const GbParams& P) {
    const int64_t strip = P.tile_h ? (int64_t)GB_TILE_C * P.tile_h : 1;
    return (P.block % strip == 0) ? P.block : ((strip % P.block == 0) ? strip : 0);
}
static bool pipeline_eligible(const GbParams& P, const gb200_range* rg) {
    const int64_t unit = pipeline_unit(P);
    return unit > 0 && rg->count >= (1 << 20) && rg->count / unit >= 8 && !getenv("GB200_NO_PIPELINE");
}
// bind(Pc, slot0): point the chunk's outputs at slot0 of the full-range device buffers;
// copy_back(slot0, cnt, stream): enqueue the device-to-host copies of that chunk.
static int run_pipelined(gb200_ctx* ctx, const gb200_range* rg, const GbParams& P,
                         const std::function<void(GbParams&, int64_t)>& bind,
                         const std::function<cudaError_t(int64_t, int64_t, cudaStream_t)>& copy_back) {
    const int K = 4;
    cudaStream_t stream = ctx->cur;
    const int64_t unit = pipeline_unit(P);
    const int64_t per = ((rg->count / unit + K - 1) / K) * unit;
    while (ctx->pool_streams.size() < 2) {
        cudaStream_t st;
        CU(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->pool_streams.push_back(st);
    }
    while (ctx->chunk_events.size() < (size_t)K + 2) {
        cudaEvent_t ev;
        CU(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->chunk_events.push_back(ev);
    }
    cudaStream_t alt = ctx->pool_streams[0], copy = ctx->pool_streams[1];
    void* qv = nullptr;
    int rc = pool_get(ctx, SL_BATCH_QUEUE, sizeof(unsigned long long) * 4 * K, &qv); if (rc) return rc;
    unsigned long long* queues = (unsigned long long*)qv;
    CU(ctx, cudaMemsetAsync(queues, 0, sizeof(unsigned long long) * 4 * K, stream));
    CU(ctx, cudaEventRecord(ctx->ev1, stream)); // uploads and queue reset done: the other stream may start
    CU(ctx, cudaStreamWaitEvent(alt, ctx->ev1, 0));
    int nchunks = 0;
    for (int c = 0; c < K; ++c) {
        const int64_t slot0 = (int64_t)c * per;
        if (slot0 >= rg->count) break;
        GbParams Pc = P; // same problem, same tables; only the slot range and the output pointers move
        Pc.first = rg->first + (slot0 / P.block) * (P.stride * P.block);
        Pc.count = (rg->count - slot0 < per) ? rg->count - slot0 : per;
        bind(Pc, slot0);
        Pc.queue = queues + 4 * c;
        Pc.counters = Pc.queue + 1;
        cudaStream_t st = (c % 2 == 0) ? stream : alt;
        int blocks = 0;
        CU(ctx, gb200_launch_trace(Pc, ctx->sm_count, st, &blocks));
        ctx->stats.launches += 1;
        CU(ctx, cudaEventRecord(ctx->chunk_events[(size_t)c], st));
        ++nchunks;
    }
    // join the kernels on the context stream (end of the kernel span) before any copy is enqueued: a device-to-host
    // copy into pageable memory blocks the calling thread, so the copies come after all launches and event records
    for (int c = 1; c < nchunks; c += 2) CU(ctx, cudaStreamWaitEvent(stream, ctx->chunk_events[(size_t)c], 0));
    CU(ctx, cudaEventRecord(ctx->ev2, stream));
    for (int c = 0; c < nchunks; ++c) {
        const int64_t slot0 = (int64_t)c * per;
        const int64_t cnt = (rg->count - slot0 < per) ? rg->count - slot0 : per;
        CU(ctx, cudaStreamWaitEvent(copy, ctx->chunk_events[(size_t)c], 0));
        CU(ctx, copy_back(slot0, cnt, copy));
    }
    CU(ctx, cudaEventRecord(ctx->chunk_events[(size_t)K], copy));
    CU(ctx, cudaStreamWaitEvent(stream, ctx->chunk_events[(size_t)K], 0));
    std::vector<unsigned long long> cq((size_t)4 * K);
    CU(ctx, cudaMemcpyAsync(cq.data(), queues, sizeof(unsigned long long) * cq.size(), cudaMemcpyDeviceToHost, stream));
    CU(ctx, cudaEventRecord(ctx->ev3, stream));
    CU(ctx, cudaStreamSynchronize(stream));
    float kms = 0, tot = 0;
    cudaEventElapsedTime(&kms, ctx->ev1, ctx->ev2);
    cudaEventElapsedTime(&tot, ctx->ev0, ctx->ev3);
    ctx->stats.kernel_ms = kms; ctx->stats.total_ms = tot; ctx->stats.rays = rg->count;
    for (int c = 0; c < nchunks; ++c) {
        ctx->stats.steps_accepted += (int64_t)cq[(size_t)c * 4 + 1]; ctx->stats.steps_rejected += (int64_t)cq[(size_t)c * 4 + 2];
        ctx->stats.flagged += (int64_t)cq[(size_t)c * 4 + 3];
    }
    return GB200_OK;
}

// ---------------------------------------------------------------- C ABI
extern "C" {

int gb200_version(void) { return GB200_VERSION; }

const char* gb200_last_error(gb200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int gb200_init(int device, gb200_ctx** out) {
    if (!out) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, GB200_ERR_NO_DEVICE, "no CUDA device available (%s); libgradus_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= ndev) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "device %d out of range [0, %d)", device, ndev);
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(nullptr, GB200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    gb200_ctx* ctx = new gb200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    CU(nullptr, cudaSetDevice(device));
    CU(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(nullptr, cudaEventCreate(&ctx->ev0)); CU(nullptr, cudaEventCreate(&ctx->ev1));
    CU(nullptr, cudaEventCreate(&ctx->ev2)); CU(nullptr, cudaEventCreate(&ctx->ev3));
    CU(nullptr, cudaEventCreateWithFlags(&ctx->inflight, cudaEventDisableTiming));
    CU(nullptr, cudaMalloc(&ctx->d_queue, 4 * sizeof(unsigned long long)));
    *out = ctx;
    return GB200_OK;
}

int gb200_set_cross_section(gb200_ctx* ctx, const double* rho, const double* height, int32_t n) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (n != 0 && (n < 2 || !rho || !height)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "cross-section table needs n >= 2 points");
    for (int i = 1; i < n; ++i) if (!(rho[i] > rho[i - 1])) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "cross-section radii must be strictly increasing");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    CU(ctx, cudaStreamSynchronize(ctx->stream)); // nothing in flight may still read the old table
    if (ctx->d_cs) { cudaFree(ctx->d_cs); ctx->d_cs = nullptr; }
    ctx->cs_n = 0; ctx->cs_max = 0.0;
    if (n == 0) return GB200_OK;
    CU(ctx, cudaMalloc(&ctx->d_cs, sizeof(double) * 2 * (size_t)n));
    CU(ctx, cudaMemcpy(ctx->d_cs, rho, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(ctx->d_cs + n, height, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    ctx->cs_n = n;
    for (int i = 0; i < n; ++i) ctx->cs_max = std::max(ctx->cs_max, height[i]);
    return GB200_OK;
}

void gb200_destroy(gb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->d_cs) cudaFree(ctx->d_cs);
    for (auto& b : ctx->pool) if (b.p) cudaFree(b.p);
    if (ctx->d_queue) cudaFree(ctx->d_queue);
    for (auto st : ctx->pool_streams) cudaStreamDestroy(st);
    for (auto ev : ctx->chunk_events) cudaEventDestroy(ev);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    if (ctx->inflight) cudaEventDestroy(ctx->inflight);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int gb200_get_stats(gb200_ctx* ctx, gb200_stats* out) {
    if (!ctx || !out) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null argument");
    *out = ctx->stats;
    return GB200_OK;
}

int gb200_validate(const gb200_problem* p, const gb200_ic* ic) { return validate(nullptr, p, ic); }

int gb200_isco(int32_t metric_kind, const double* mp, double* out) {
    if (!mp || !out) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null argument");
    if (metric_kind == GB200_METRIC_KERR) { *out = kerr_isco_host(mp[0], mp[1]); return GB200_OK; }
    if (metric_kind == GB200_METRIC_MORRIS_THORNE)
        return fail(nullptr, GB200_ERR_UNSUPPORTED, "the Morris-Thorne wormhole has no circular-orbit energy minimum: no ISCO");
    if (metric_kind > GB200_METRIC_KERR && metric_kind < GB200_METRIC_COUNT) {
        int rc = generic_isco_host(metric_kind, mp, out);
        if (rc) return fail(nullptr, rc, "No boundaries for minimization could be determined. It is likely this configuration does not have an ISCO solution.");
        return GB200_OK;
    }
    return fail(nullptr, GB200_ERR_UNSUPPORTED, "metric kind %d", metric_kind);
}

int gb200_radiative_efficiency(int32_t metric_kind, const double* mp, double* out) {
    double r = 0;
    int rc = gb200_isco(metric_kind, mp, &r);
    if (rc) return rc;
    *out = 1.0 - circular_energy_host<double>(metric_kind, mp, r);
    return GB200_OK;
}

// gb200_trace and gb200_trace_target: `target` (r, theta, phi) non-null selects the closest-approach objective (the generic
// integrator with the distance condition in place of a geometry); `closest` then receives one distance per ray.
static int trace_common(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, gb200_endpoints* out,
                        const double* target, double d_tol, double* closest) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    int rc = validate(ctx, p, ic); if (rc) return rc;
    rc = validate_range(ctx, ic, rg); if (rc) return rc;
    static const gb200_endpoints no_endpoints{};
    if (target) {
        if (p->geometry_kind != GB200_GEOMETRY_NONE) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "gb200_trace_target: the distance callback takes the geometry's place (geometry_kind must be NONE)");
        if (!(d_tol > 0) || !closest) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "gb200_trace_target needs d_tol > 0 and a closest[] array");
        if (!out) out = const_cast<gb200_endpoints*>(&no_endpoints);
    } else if (!out) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null endpoints");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    GbParams P;
    fill_params(p, ic, rg, P);
    { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
    rc = upload_ic_and_tables(ctx, ic, nullptr, nullptr, P); if (rc) return rc;
    const size_t n = (size_t)rg->count;
    void* d;
    if (target) { // to_cartesian(target), src/geometry/geometry.jl:13-16
        P.geometry_kind = GB200_GEOMETRY_TARGET_POINT;
        P.gp0 = target[0] * std::sin(target[1]) * std::cos(target[2]);
        P.gp1 = target[0] * std::sin(target[1]) * std::sin(target[2]);
        P.gp2 = target[0] * std::cos(target[1]);
        P.gtol = d_tol;
        rc = pool_get(ctx, SL_G, n * sizeof(double) + 8, &d); if (rc) return rc;
        P.o_closest = (double*)d;
    }
#define WANT(ptr, slot, type, field)                                               \
    if (ptr) { rc = pool_get(ctx, slot, n * sizeof(type) + 8, &d); if (rc) return rc; field = (type*)d; }
    WANT(out->status, SL_STATUS, int32_t, P.o_status)
    WANT(out->lambda_max, SL_LAMBDA, double, P.o_lambda)
    for (int k = 0; k < 4; ++k) {
        WANT(out->x[k], SL_X0 + k, double, P.o_x[k])
        WANT(out->v[k], SL_V0 + k, double, P.o_v[k])
        WANT(out->x_init[k], SL_XI0 + k, double, P.o_x0[k])
        WANT(out->v_init[k], SL_VI0 + k, double, P.o_v0[k])
    }
    WANT(out->naccept, SL_NACC, int32_t, P.o_naccept)
    WANT(out->nreject, SL_NREJ, int32_t, P.o_nreject)
    WANT(out->flags, SL_FLAGS, int32_t, P.o_flags)
#undef WANT
    if (!target && pipeline_eligible(P, rg)) {
        const GbParams Pf = P;
        return run_pipelined(ctx, rg, P,
            [&](GbParams& Pc, int64_t s0) {
                if (Pf.o_status) Pc.o_status = Pf.o_status + s0;
                if (Pf.o_lambda) Pc.o_lambda = Pf.o_lambda + s0;
                for (int k = 0; k < 4; ++k) {
                    if (Pf.o_x[k]) Pc.o_x[k] = Pf.o_x[k] + s0;
                    if (Pf.o_v[k]) Pc.o_v[k] = Pf.o_v[k] + s0;
                    if (Pf.o_x0[k]) Pc.o_x0[k] = Pf.o_x0[k] + s0;
                    if (Pf.o_v0[k]) Pc.o_v0[k] = Pf.o_v0[k] + s0;
                }
                if (Pf.o_naccept) Pc.o_naccept = Pf.o_naccept + s0;
                if (Pf.o_nreject) Pc.o_nreject = Pf.o_nreject + s0;
                if (Pf.o_flags) Pc.o_flags = Pf.o_flags + s0;
            },
            [&](int64_t s0, int64_t cnt, cudaStream_t copy) -> cudaError_t {
                cudaError_t e = cudaSuccess;
#define BACKC(ptr, field, type) \
    if (e == cudaSuccess && ptr) e = cudaMemcpyAsync(ptr + s0, field + s0, (size_t)cnt * sizeof(type), cudaMemcpyDeviceToHost, copy);
                BACKC(out->status, Pf.o_status, int32_t)
                BACKC(out->lambda_max, Pf.o_lambda, double)
                for (int k = 0; k < 4; ++k) {
                    BACKC(out->x[k], Pf.o_x[k], double)
                    BACKC(out->v[k], Pf.o_v[k], double)
                    BACKC(out->x_init[k], Pf.o_x0[k], double)
                    BACKC(out->v_init[k], Pf.o_v0[k], double)
                }
                BACKC(out->naccept, Pf.o_naccept, int32_t)
                BACKC(out->nreject, Pf.o_nreject, int32_t)
                BACKC(out->flags, Pf.o_flags, int32_t)
#undef BACKC
                return e;
            });
    }
    if (target) { // one ray per thread through the generic integrator; no ticket queue, no step counters
        CU(ctx, cudaMemsetAsync(ctx->d_queue, 0, 4 * sizeof(unsigned long long), ctx->stream));
        CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU(ctx, gb200_launch_target(P, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
        ctx->stats.launches += n ? 1 : 0;
        if (n) CU(ctx, cudaMemcpyAsync(closest, P.o_closest, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    } else { rc = run_trace(ctx, P, ctx->stream, true); if (rc) return rc; }
#define BACK(ptr, field, type) \
    if (ptr && n) CU(ctx, cudaMemcpyAsync(ptr, field, n * sizeof(type), cudaMemcpyDeviceToHost, ctx->stream));
    BACK(out->status, P.o_status, int32_t)
    BACK(out->lambda_max, P.o_lambda, double)
    for (int k = 0; k < 4; ++k) {
        BACK(out->x[k], P.o_x[k], double)
        BACK(out->v[k], P.o_v[k], double)
        BACK(out->x_init[k], P.o_x0[k], double)
        BACK(out->v_init[k], P.o_v0[k], double)
    }
    BACK(out->naccept, P.o_naccept, int32_t)
    BACK(out->nreject, P.o_nreject, int32_t)
    BACK(out->flags, P.o_flags, int32_t)
#undef BACK
    return finish_stats(ctx, rg->count);
}

int gb200_trace(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, gb200_endpoints* out) {
    return trace_common(ctx, p, ic, rg, out, nullptr, 0.0, nullptr);
}

int gb200_trace_target(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const double* target, double d_tol,
                       gb200_endpoints* out, double* closest) {
    if (!target) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null target");
    return trace_common(ctx, p, ic, rg, out, target, d_tol, closest);
}

int gb200_trace_batch(gb200_ctx* ctx, int32_t nbatch, const gb200_problem* problems, const gb200_ic* ics,
                      const gb200_range* ranges, gb200_endpoints* outs) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (nbatch < 1 || !problems || !ics || !ranges || !outs) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad batch arguments");
    for (int b = 0; b < nbatch; ++b) {
        int rc = validate(ctx, &problems[b], &ics[b]); if (rc) return rc;
        rc = validate_range(ctx, &ics[b], &ranges[b]); if (rc) return rc;
        if (ics[b].kind == GB200_IC_IMPACT_PARAMETERS) return fail(ctx, GB200_ERR_UNSUPPORTED, "impact-parameter lists are not batched; use gb200_trace / gb200_render");
    }
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    const int nstreams = nbatch < 32 ? nbatch : 32;
    while ((int)ctx->pool_streams.size() < nstreams) {
        cudaStream_t st;
        CU(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->pool_streams.push_back(st);
    }
    // One device arena: [explicit ICs of every ensemble | outputs of every ensemble], mirrored by one pinned host
    // staging buffer, so the whole batch costs one H2D and one D2H copy instead of ~20 small ones per ensemble.
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    struct Off { size_t status, lambda, x[4], v[4], x0[4], v0[4], nacc, nrej, flags, ex[4], ev[4]; };
    std::vector<Off> off((size_t)nbatch);
    size_t total = 0;
    for (int b = 0; b < nbatch; ++b)
        if (ics[b].kind == GB200_IC_EXPLICIT)
            for (int k = 0; k < 4; ++k) { off[(size_t)b].ex[k] = total; total += al(8 * (size_t)ics[b].n); off[(size_t)b].ev[k] = total; total += al(8 * (size_t)ics[b].n); }
    const size_t in_bytes = total;
    for (int b = 0; b < nbatch; ++b) {
        const size_t n = (size_t)ranges[b].count;
        Off& o = off[(size_t)b];
        o.status = total; total += al(4 * n);
        o.lambda = total; total += al(8 * n);
        for (int k = 0; k < 4; ++k) { o.x[k] = total; total += al(8 * n); o.v[k] = total; total += al(8 * n);
                                      o.x0[k] = total; total += al(8 * n); o.v0[k] = total; total += al(8 * n); }
        o.nacc = total; total += al(4 * n); o.nrej = total; total += al(4 * n); o.flags = total; total += al(4 * n);
    }
    void* arena_v = nullptr;
    int rc = pool_get(ctx, SL_BATCH, total + 256, &arena_v); if (rc) return rc;
    char* arena = (char*)arena_v;
    if (ctx->stage_cap < total) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr; ctx->stage_cap = 0;
        CU(ctx, cudaMallocHost(&ctx->stage, total + 256));
        ctx->stage_cap = total;
    }
    char* stage = (char*)ctx->stage;
    void* qv = nullptr;
    rc = pool_get(ctx, SL_BATCH_QUEUE, sizeof(unsigned long long) * 4 * (size_t)nbatch, &qv); if (rc) return rc;
    unsigned long long* queues = (unsigned long long*)qv;
    for (int b = 0; b < nbatch; ++b)
        if (ics[b].kind == GB200_IC_EXPLICIT)
            for (int k = 0; k < 4; ++k) {
                memcpy(stage + off[(size_t)b].ex[k], ics[b].x[k], 8 * (size_t)ics[b].n);
                memcpy(stage + off[(size_t)b].ev[k], ics[b].v[k], 8 * (size_t)ics[b].n);
            }
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CU(ctx, cudaMemsetAsync(queues, 0, sizeof(unsigned long long) * 4 * (size_t)nbatch, ctx->stream));
    if (in_bytes) CU(ctx, cudaMemcpyAsync(arena, stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    for (int sidx = 0; sidx < nstreams; ++sidx) CU(ctx, cudaStreamWaitEvent(ctx->pool_streams[(size_t)sidx], ctx->ev1, 0));
    for (int b = 0; b < nbatch; ++b) {
        cudaStream_t st = ctx->pool_streams[(size_t)(b % nstreams)];
        const gb200_endpoints& out = outs[b];
        const Off& o = off[(size_t)b];
        GbParams P;
        fill_params(&problems[b], &ics[b], &ranges[b], P);
        { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
        if (ics[b].kind == GB200_IC_EXPLICIT)
            for (int k = 0; k < 4; ++k) { P.ex[k] = (const double*)(arena + o.ex[k]); P.ev[k] = (const double*)(arena + o.ev[k]); }
        if (out.status) P.o_status = (int32_t*)(arena + o.status);
        if (out.lambda_max) P.o_lambda = (double*)(arena + o.lambda);
        for (int k = 0; k < 4; ++k) {
            if (out.x[k]) P.o_x[k] = (double*)(arena + o.x[k]);
            if (out.v[k]) P.o_v[k] = (double*)(arena + o.v[k]);
            if (out.x_init[k]) P.o_x0[k] = (double*)(arena + o.x0[k]);
            if (out.v_init[k]) P.o_v0[k] = (double*)(arena + o.v0[k]);
        }
        if (out.naccept) P.o_naccept = (int32_t*)(arena + o.nacc);
        if (out.nreject) P.o_nreject = (int32_t*)(arena + o.nrej);
        if (out.flags) P.o_flags = (int32_t*)(arena + o.flags);
        P.queue = queues + 4 * (size_t)b;
        P.counters = P.queue + 1;
        int blocks = 0;
        if (ranges[b].count) { CU(ctx, gb200_launch_trace(P, ctx->sm_count, st, &blocks)); ctx->stats.launches += 1; }
    }
    for (int sidx = 0; sidx < nstreams; ++sidx) { // join the pool back into the context stream
        CU(ctx, cudaEventRecord(ctx->ev2, ctx->pool_streams[(size_t)sidx]));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev2, 0));
    }
    std::vector<unsigned long long> c((size_t)nbatch * 4);
    if (total > in_bytes) CU(ctx, cudaMemcpyAsync(stage + in_bytes, arena + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(c.data(), queues, sizeof(unsigned long long) * c.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < nbatch; ++b) { // scatter the staging buffer into the caller's SoA
        const gb200_endpoints& out = outs[b];
        const Off& o = off[(size_t)b];
        const size_t n = (size_t)ranges[b].count;
        if (out.status) memcpy(out.status, stage + o.status, 4 * n);
        if (out.lambda_max) memcpy(out.lambda_max, stage + o.lambda, 8 * n);
        for (int k = 0; k < 4; ++k) {
            if (out.x[k]) memcpy(out.x[k], stage + o.x[k], 8 * n);
            if (out.v[k]) memcpy(out.v[k], stage + o.v[k], 8 * n);
            if (out.x_init[k]) memcpy(out.x_init[k], stage + o.x0[k], 8 * n);
            if (out.v_init[k]) memcpy(out.v_init[k], stage + o.v0[k], 8 * n);
        }
        if (out.naccept) memcpy(out.naccept, stage + o.nacc, 4 * n);
        if (out.nreject) memcpy(out.nreject, stage + o.nrej, 4 * n);
        if (out.flags) memcpy(out.flags, stage + o.flags, 4 * n);
    }
    float tot = 0, dev = 0;
    cudaEventElapsedTime(&tot, ctx->ev0, ctx->ev3);
    cudaEventElapsedTime(&dev, ctx->ev1, ctx->ev2); // launches only (ev2: the end of the last pool stream), without the two staged copies
    ctx->stats.kernel_ms = dev; ctx->stats.total_ms = tot;
    for (int b = 0; b < nbatch; ++b) {
        ctx->stats.rays += ranges[b].count;
        ctx->stats.steps_accepted += (int64_t)c[(size_t)b * 4 + 1]; ctx->stats.steps_rejected += (int64_t)c[(size_t)b * 4 + 2];
        ctx->stats.flagged += (int64_t)c[(size_t)b * 4 + 3];
    }
    return GB200_OK;
}

int gb200_trace_path(gb200_ctx* ctx, const gb200_problem* p, const double* u0, int32_t cap, double* lambda, double* u,
                     int32_t* nrows, int32_t* status) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (!p || !u0 || cap < 1 || !lambda || !u || !nrows || !status) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad path arguments");
    gb200_ic ic{};
    ic.kind = GB200_IC_EXPLICIT; ic.n = 1;
    for (int k = 0; k < 4; ++k) { ic.x[k] = u0 + k; ic.v[k] = u0 + 4 + k; }
    int rc = validate(ctx, p, &ic); if (rc) return rc;
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    gb200_range rg{0, 1, 1};
    GbParams P;
    fill_params(p, &ic, &rg, P);
    { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
    P.geometry_kind = GB200_GEOMETRY_NONE; // a bare geodesic: chart boundaries and lambda_max only
    P.callback_kind = GB200_CALLBACK_NONE;
    void *d_u0, *d_lam, *d_u, *d_meta;
    rc = pool_get(ctx, SL_EX0, 64, &d_u0); if (rc) return rc;
    rc = pool_get(ctx, SL_LAMBDA, sizeof(double) * (size_t)cap, &d_lam); if (rc) return rc;
    rc = pool_get(ctx, SL_X0, sizeof(double) * 8 * (size_t)cap, &d_u); if (rc) return rc;
    rc = pool_get(ctx, SL_SCRATCH, 64, &d_meta); if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d_u0, u0, 64, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, gb200_launch_path(P, (const double*)d_u0, cap, (double*)d_lam, (double*)d_u, (int*)d_meta, ctx->stream));
    ctx->stats.launches = 1;
    int meta[2] = {0, 0};
    CU(ctx, cudaMemcpyAsync(meta, d_meta, sizeof meta, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    const int nw = meta[0] < cap ? meta[0] : cap;
    CU(ctx, cudaMemcpy(lambda, d_lam, sizeof(double) * (size_t)nw, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(u, d_u, sizeof(double) * 8 * (size_t)nw, cudaMemcpyDeviceToHost));
    *nrows = meta[0];
    *status = meta[1];
    return GB200_OK;
}

int gb200_build_plunging_table(gb200_ctx* ctx, int32_t metric_kind, const double* mp, int32_t cap, double* r, double* ut, double* ur,
                         double* uphi, int32_t* n) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (!mp || cap < 2 || !r || !ut || !ur || !uphi || !n) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad plunging-table arguments");
    double risco = 0;
    int rc = gb200_isco(metric_kind, mp, &risco);
    if (rc) return fail(ctx, rc, "no ISCO for this metric");
    // CircularOrbits.plunging_fourvelocity at the ISCO (circular-orbits.jl:129-150)
    double g[5], dr[5], dth[5];
    metric_jacobian_kind<double>(metric_kind, mp, risco, 1.0, 0.0, g, dr, dth);
    const double D = g[0] * g[3] - g[4] * g[4];
    const double gitt = g[3] / D, giphph = g[0] / D, gitph = -g[4] / D;
    const double Om = -(dr[4] - std::sqrt(dr[4] * dr[4] - dr[0] * dr[3])) / dr[3];
    const double A = -(Om * gitt - gitph), B = (Om * gitph - giphph);
    const double denom = B * B * gitt + 2.0 * A * B * gitph + A * A * giphph;
    const double dd = -(denom > 0 ? 1.0 : -1.0) * std::sqrt(1.0 / std::fabs(denom));
    const double ut_ = B * dd, uph_ = A * dd; // covariant u_t, u_phi
    const double E = -ut_, L = uph_;
    const double vt = gitt * ut_ + gitph * uph_, vph = gitph * ut_ + giphph * uph_;
    const double nom = gitt * E * E - 2.0 * gitph * E * L + giphph * L * L + 1.0;
    const double vr = -std::sqrt(std::fabs(nom / (-g[1])));
    gb200_problem p{};
    p.metric_kind = metric_kind;
    for (int k = 0; k < 8; ++k) p.metric_params[k] = mp[k];
    p.mu = 1.0; p.abstol = 1e-9; p.reltol = 1e-9; p.lambda_min = 0.0; p.lambda_max = 50000.0; p.gtol = 1e-2;
    const double q2 = metric_kind == GB200_METRIC_KERR_NEWMAN ? mp[2] * mp[2] : 0.0;
    const double rh = mp[0] + std::sqrt(mp[0] * mp[0] - mp[1] * mp[1] - q2);
    p.chart_inner = rh * 1.000001; p.chart_outer = 12000.0;
    const double u0[8] = {0.0, risco - 1e-8, M_PI / 2, 0.0, vt, vr, 0.0, vph};
    const int pathcap = 1 << 20;
    std::vector<double> lam((size_t)pathcap), u((size_t)pathcap * 8);
    int32_t rows = 0, status = 0;
    rc = gb200_trace_path(ctx, &p, u0, pathcap, lam.data(), u.data(), &rows, &status);
    if (rc) return rc;
    if (rows > pathcap) return fail(ctx, GB200_ERR_NOMEM, "plunging geodesic needs %d rows (> %d)", rows, pathcap);
    // PlungingInterpolation (orbit-solving.jl:99-135): sort by r, drop the innermost sample
    std::vector<int> idx((size_t)rows);
    for (int i = 0; i < rows; ++i) idx[(size_t)i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a_, int b_) { return u[(size_t)a_ * 8 + 1] < u[(size_t)b_ * 8 + 1]; });
    const int nt = rows - 1;
    if (nt > cap) return fail(ctx, GB200_ERR_NOMEM, "plunging table needs %d entries (cap %d)", nt, cap);
    for (int i = 0; i < nt; ++i) {
        const size_t s = (size_t)idx[(size_t)i + 1] * 8;
        r[i] = u[s + 1]; ut[i] = u[s + 4]; ur[i] = u[s + 5]; uphi[i] = u[s + 7];
    }
    *n = nt;
    return GB200_OK;
}

static int render_common(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const int32_t* pfs, int32_t npf,
                         const gb200_plunging_table* pl, double* const* images, bool device_out, cudaStream_t stream, int async) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    int rc = validate(ctx, p, ic); if (rc) return rc;
    rc = validate_range(ctx, ic, rg); if (rc) return rc;
    if (npf < 1 || npf > GB_MAX_PF || !pfs || !images) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "need 1..%d point functions", GB_MAX_PF);
    for (int k = 0; k < npf; ++k) {
        if (pfs[k] < GB200_PF_SHADOW || pfs[k] > GB200_PF_RADIUS) return fail(ctx, GB200_ERR_UNSUPPORTED, "point function %d", pfs[k]);
        if (!images[k]) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null image pointer");
    }
    { int rc_ = begin_call(ctx, stream); if (rc_) return rc_; }
    CU(ctx, cudaEventRecord(ctx->ev0, stream));
    GbParams P;
    fill_params(p, ic, rg, P);
    { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
    if (needs_isco(pfs, npf, false)) { rc = set_isco(ctx, p, P); if (rc) return rc; }
    rc = upload_ic_and_tables(ctx, ic, pl, nullptr, P); if (rc) return rc;
    const size_t n = (size_t)rg->count;
    P.npf = npf;
    for (int k = 0; k < npf; ++k) {
        P.pf[k] = pfs[k];
        if (device_out) P.o_img[k] = images[k];
        else { void* d; rc = pool_get(ctx, SL_IMG0 + k, n * sizeof(double) + 8, &d); if (rc) return rc; P.o_img[k] = (double*)d; }
    }
    if (!device_out && !async && pipeline_eligible(P, rg)) {
        const GbParams Pfull = P;
        return run_pipelined(ctx, rg, P,
            [&](GbParams& Pc, int64_t slot0) { for (int k = 0; k < npf; ++k) Pc.o_img[k] = Pfull.o_img[k] + slot0; },
            [&](int64_t slot0, int64_t cnt, cudaStream_t copy) -> cudaError_t {
                for (int k = 0; k < npf; ++k) {
                    cudaError_t e = cudaMemcpyAsync(images[k] + slot0, Pfull.o_img[k] + slot0, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, copy);
                    if (e != cudaSuccess) return e;
                }
                return cudaSuccess;
            });
    }
    rc = run_trace(ctx, P, stream, true); if (rc) return rc;
    if (!device_out)
        for (int k = 0; k < npf; ++k)
            if (n) CU(ctx, cudaMemcpyAsync(images[k], P.o_img[k], n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (async) { ctx->stats.rays = rg->count; return end_async_call(ctx, stream); }
    return finish_stats(ctx, rg->count);
}

int gb200_render(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const int32_t* pfs, int32_t npf,
                 const gb200_plunging_table* pl, double* const* images) {
    return render_common(ctx, p, ic, rg, pfs, npf, pl, images, false, ctx ? ctx->stream : nullptr, 0);
}

int gb200_render_device(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const int32_t* pfs, int32_t npf,
                        const gb200_plunging_table* pl, double* const* d_images, void* cuda_stream, int async) {
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : (ctx ? ctx->stream : nullptr);
    return render_common(ctx, p, ic, rg, pfs, npf, pl, d_images, true, s, async);
}

int gb200_render_batch(gb200_ctx* ctx, int32_t nbatch, const gb200_problem* problems, const gb200_ic* ics, const gb200_range* ranges,
                       const int32_t* pfs, int32_t npf, const gb200_plunging_table* const* pls, double* const* images) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (nbatch < 1 || !problems || !ics || !ranges || !pfs || !images) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad batch arguments");
    if (npf < 1 || npf > GB_MAX_PF) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "need 1..%d point functions", GB_MAX_PF);
    for (int k = 0; k < npf; ++k)
        if (pfs[k] < GB200_PF_SHADOW || pfs[k] > GB200_PF_RADIUS) return fail(ctx, GB200_ERR_UNSUPPORTED, "point function %d", pfs[k]);
    for (int b = 0; b < nbatch; ++b) {
        int rc = validate(ctx, &problems[b], &ics[b]); if (rc) return rc;
        rc = validate_range(ctx, &ics[b], &ranges[b]); if (rc) return rc;
        if (ics[b].kind == GB200_IC_EXPLICIT) return fail(ctx, GB200_ERR_UNSUPPORTED, "explicit initial conditions are batched by gb200_trace_batch");
        for (int k = 0; k < npf; ++k) if (ranges[b].count && !images[(size_t)b * npf + k]) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null image pointer");
    }
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    const int nstreams = nbatch < 32 ? nbatch : 32;
    while ((int)ctx->pool_streams.size() < nstreams) {
        cudaStream_t st;
        CU(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->pool_streams.push_back(st);
    }
    // One arena: [impact-parameter lists (+ heights) and plunging tables of every problem | images of every problem],
    // mirrored by one pinned staging buffer: one H2D and one D2H copy for the whole batch.
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    struct Off { size_t ex[3], pl[4], img[GB_MAX_PF]; };
    std::vector<Off> off((size_t)nbatch);
    size_t total = 0;
    for (int b = 0; b < nbatch; ++b) {
        Off& o = off[(size_t)b];
        if (ics[b].kind == GB200_IC_IMPACT_PARAMETERS)
            for (int k = 0; k < 3; ++k) if (ics[b].x[k]) { o.ex[k] = total; total += al(8 * (size_t)ics[b].n); }
        if (pls && pls[b] && pls[b]->n >= 2) {
            if (b > 0 && pls[b] == pls[b - 1]) { for (int k = 0; k < 4; ++k) o.pl[k] = off[(size_t)b - 1].pl[k]; }
            else for (int k = 0; k < 4; ++k) { o.pl[k] = total; total += al(8 * (size_t)pls[b]->n); }
        }
    }
    const size_t in_bytes = total;
    for (int b = 0; b < nbatch; ++b)
        for (int k = 0; k < npf; ++k) { off[(size_t)b].img[k] = total; total += al(8 * (size_t)ranges[b].count); }
    void* arena_v = nullptr;
    int rc = pool_get(ctx, SL_BATCH, total + 256, &arena_v); if (rc) return rc;
    char* arena = (char*)arena_v;
    if (ctx->stage_cap < total) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr; ctx->stage_cap = 0;
        CU(ctx, cudaMallocHost(&ctx->stage, total + 256));
        ctx->stage_cap = total;
    }
    char* stage = (char*)ctx->stage;
    void* qv = nullptr;
    rc = pool_get(ctx, SL_BATCH_QUEUE, sizeof(unsigned long long) * 4 * (size_t)nbatch, &qv); if (rc) return rc;
    unsigned long long* queues = (unsigned long long*)qv;
    for (int b = 0; b < nbatch; ++b) {
        const Off& o = off[(size_t)b];
        if (ics[b].kind == GB200_IC_IMPACT_PARAMETERS)
            for (int k = 0; k < 3; ++k) if (ics[b].x[k]) memcpy(stage + o.ex[k], ics[b].x[k], 8 * (size_t)ics[b].n);
        if (pls && pls[b] && pls[b]->n >= 2 && !(b > 0 && pls[b] == pls[b - 1])) {
            const double* src[4] = {pls[b]->r, pls[b]->ut, pls[b]->ur, pls[b]->uphi};
            for (int k = 0; k < 4; ++k) memcpy(stage + o.pl[k], src[k], 8 * (size_t)pls[b]->n);
        }
    }
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CU(ctx, cudaMemsetAsync(queues, 0, sizeof(unsigned long long) * 4 * (size_t)nbatch, ctx->stream));
    if (in_bytes) CU(ctx, cudaMemcpyAsync(arena, stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    for (int sidx = 0; sidx < nstreams; ++sidx) CU(ctx, cudaStreamWaitEvent(ctx->pool_streams[(size_t)sidx], ctx->ev1, 0));
    const bool want_isco = needs_isco(pfs, npf, false);
    for (int b = 0; b < nbatch; ++b) {
        if (!ranges[b].count) continue;
        cudaStream_t st = ctx->pool_streams[(size_t)(b % nstreams)];
        const Off& o = off[(size_t)b];
        GbParams P;
        fill_params(&problems[b], &ics[b], &ranges[b], P);
        { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
        if (want_isco) { rc = set_isco(ctx, &problems[b], P); if (rc) return rc; }
        if (ics[b].kind == GB200_IC_IMPACT_PARAMETERS)
            for (int k = 0; k < 3; ++k) P.ex[k] = ics[b].x[k] ? (const double*)(arena + o.ex[k]) : nullptr;
        if (pls && pls[b] && pls[b]->n >= 2) {
            P.pl_n = pls[b]->n;
            P.pl_r = (const double*)(arena + o.pl[0]); P.pl_ut = (const double*)(arena + o.pl[1]);
            P.pl_ur = (const double*)(arena + o.pl[2]); P.pl_uphi = (const double*)(arena + o.pl[3]);
        }
        P.npf = npf;
        for (int k = 0; k < npf; ++k) { P.pf[k] = pfs[k]; P.o_img[k] = (double*)(arena + o.img[k]); }
        P.queue = queues + 4 * (size_t)b;
        P.counters = P.queue + 1;
        int blocks = 0;
        CU(ctx, gb200_launch_trace(P, ctx->sm_count, st, &blocks));
        ctx->stats.launches += 1;
    }
    for (int sidx = 0; sidx < nstreams; ++sidx) { // join the pool back into the context stream
        CU(ctx, cudaEventRecord(ctx->ev2, ctx->pool_streams[(size_t)sidx]));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev2, 0));
    }
    std::vector<unsigned long long> c((size_t)nbatch * 4);
    if (total > in_bytes) CU(ctx, cudaMemcpyAsync(stage + in_bytes, arena + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(c.data(), queues, sizeof(unsigned long long) * c.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < nbatch; ++b)
        for (int k = 0; k < npf; ++k)
            if (ranges[b].count) memcpy(images[(size_t)b * npf + k], stage + off[(size_t)b].img[k], 8 * (size_t)ranges[b].count);
    float tot = 0;
    cudaEventElapsedTime(&tot, ctx->ev0, ctx->ev3);
    ctx->stats.kernel_ms = tot; ctx->stats.total_ms = tot;
    for (int b = 0; b < nbatch; ++b) {
        ctx->stats.rays += ranges[b].count;
        ctx->stats.steps_accepted += (int64_t)c[(size_t)b * 4 + 1]; ctx->stats.steps_rejected += (int64_t)c[(size_t)b * 4 + 2];
        ctx->stats.flagged += (int64_t)c[(size_t)b * 4 + 3];
    }
    return GB200_OK;
}

// ---------------------------------------------------------------- forward-mode traces (gb200_dual.cu)
int gb200_trace_dual_batch(gb200_ctx* ctx, int32_t nbatch, const gb200_problem* problems, const gb200_dual_ic* ics, int32_t norm_mode,
                           const gb200_plunging_table* const* pls, gb200_dual_out* outs) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (nbatch < 1 || !problems || !ics || !outs) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad batch arguments");
    if (norm_mode != GB200_DUAL_NORM_WITH_PARTIALS && norm_mode != GB200_DUAL_NORM_VALUES_ONLY) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad norm_mode");
    std::vector<gb200_ic> fake((size_t)nbatch);
    for (int b = 0; b < nbatch; ++b) {
        const gb200_dual_ic& d = ics[b];
        if (d.n < 0 || (d.npartials != 1 && d.npartials != 2)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "dual IC needs n >= 0 and 1 or 2 partials");
        if (d.n > 0 && (!d.alpha || !d.beta || !d.dalpha || !d.dbeta)) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "dual IC needs alpha, beta and their partials");
        gb200_ic& f = fake[(size_t)b];
        memset(&f, 0, sizeof f);
        f.kind = GB200_IC_IMPACT_PARAMETERS; f.n = d.n > 0 ? d.n : 1;
        f.x[0] = d.alpha ? d.alpha : (const double*)&f; f.x[1] = d.beta ? d.beta : (const double*)&f;
        int rc = validate(ctx, &problems[b], &f); if (rc) return rc;
        if (d.height && problems[b].geometry_kind != GB200_GEOMETRY_DATUM_PLANE) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "per-ray heights need GB200_GEOMETRY_DATUM_PLANE");
    }
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    const int nstreams = nbatch < 32 ? nbatch : 32;
    while ((int)ctx->pool_streams.size() < nstreams) {
        cudaStream_t st;
        CU(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->pool_streams.push_back(st);
    }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    struct Off { size_t alpha, beta, dalpha, dbeta, height, pl[4], status, lambda, x[4], v[4], g, dg, rho, drho, nacc, nrej, flags; };
    std::vector<Off> off((size_t)nbatch);
    size_t total = 0;
    for (int b = 0; b < nbatch; ++b) {
        Off& o = off[(size_t)b];
        const size_t n = (size_t)ics[b].n, nd = (size_t)ics[b].npartials;
        o.alpha = total; total += al(8 * n); o.beta = total; total += al(8 * n);
        o.dalpha = total; total += al(8 * n * nd); o.dbeta = total; total += al(8 * n * nd);
        o.height = total; if (ics[b].height) total += al(8 * n);
        if (pls && pls[b] && pls[b]->n >= 2) {
            if (b > 0 && pls[b] == pls[b - 1]) { for (int k = 0; k < 4; ++k) o.pl[k] = off[(size_t)b - 1].pl[k]; }
            else for (int k = 0; k < 4; ++k) { o.pl[k] = total; total += al(8 * (size_t)pls[b]->n); }
        }
    }
    const size_t in_bytes = total;
    for (int b = 0; b < nbatch; ++b) {
        Off& o = off[(size_t)b];
        const size_t n = (size_t)ics[b].n, nd = (size_t)ics[b].npartials;
        o.status = total; total += al(4 * n); o.lambda = total; total += al(8 * n);
        for (int k = 0; k < 4; ++k) { o.x[k] = total; total += al(8 * n); o.v[k] = total; total += al(8 * n); }
        o.g = total; total += al(8 * n); o.dg = total; total += al(8 * n * nd);
        o.rho = total; total += al(8 * n); o.drho = total; total += al(8 * n * nd);
        o.nacc = total; total += al(4 * n); o.nrej = total; total += al(4 * n); o.flags = total; total += al(4 * n);
    }
    void* arena_v = nullptr;
    int rc = pool_get(ctx, SL_BATCH, total + 256, &arena_v); if (rc) return rc;
    char* arena = (char*)arena_v;
    if (ctx->stage_cap < total) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr; ctx->stage_cap = 0;
        CU(ctx, cudaMallocHost(&ctx->stage, total + 256));
        ctx->stage_cap = total;
    }
    char* stage = (char*)ctx->stage;
    for (int b = 0; b < nbatch; ++b) {
        const Off& o = off[(size_t)b];
        const gb200_dual_ic& d = ics[b];
        const size_t n = (size_t)d.n, nd = (size_t)d.npartials;
        if (n) {
            memcpy(stage + o.alpha, d.alpha, 8 * n); memcpy(stage + o.beta, d.beta, 8 * n);
            memcpy(stage + o.dalpha, d.dalpha, 8 * n * nd); memcpy(stage + o.dbeta, d.dbeta, 8 * n * nd);
            if (d.height) memcpy(stage + o.height, d.height, 8 * n);
        }
        if (pls && pls[b] && pls[b]->n >= 2 && !(b > 0 && pls[b] == pls[b - 1])) {
            const double* src[4] = {pls[b]->r, pls[b]->ut, pls[b]->ur, pls[b]->uphi};
            for (int k = 0; k < 4; ++k) memcpy(stage + o.pl[k], src[k], 8 * (size_t)pls[b]->n);
        }
    }
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (in_bytes) CU(ctx, cudaMemcpyAsync(arena, stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    for (int sidx = 0; sidx < nstreams; ++sidx) CU(ctx, cudaStreamWaitEvent(ctx->pool_streams[(size_t)sidx], ctx->ev1, 0));
    for (int b = 0; b < nbatch; ++b) {
        const gb200_dual_ic& d = ics[b];
        if (d.n == 0) continue;
        const gb200_dual_out& out = outs[b];
        const Off& o = off[(size_t)b];
        gb200_range rg{0, d.n, 1, 1};
        GbParams P;
        fill_params(&problems[b], &fake[(size_t)b], &rg, P);
        { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
        if (out.g) { rc = set_isco(ctx, &problems[b], P); if (rc) return rc; }
        if (pls && pls[b] && pls[b]->n >= 2) {
            P.pl_n = pls[b]->n;
            P.pl_r = (const double*)(arena + o.pl[0]); P.pl_ut = (const double*)(arena + o.pl[1]);
            P.pl_ur = (const double*)(arena + o.pl[2]); P.pl_uphi = (const double*)(arena + o.pl[3]);
        }
        GbDualIO io;
        memset(&io, 0, sizeof io);
        io.n = d.n; io.npartials = d.npartials; io.norm_partials = (norm_mode == GB200_DUAL_NORM_WITH_PARTIALS) ? 1 : 0;
        io.alpha = (const double*)(arena + o.alpha); io.beta = (const double*)(arena + o.beta);
        io.dalpha = (const double*)(arena + o.dalpha); io.dbeta = (const double*)(arena + o.dbeta);
        io.height = d.height ? (const double*)(arena + o.height) : nullptr;
        if (out.status) io.status = (int32_t*)(arena + o.status);
        if (out.lambda_max) io.lambda = (double*)(arena + o.lambda);
        for (int k = 0; k < 4; ++k) {
            if (out.x[k]) io.x[k] = (double*)(arena + o.x[k]);
            if (out.v[k]) io.v[k] = (double*)(arena + o.v[k]);
        }
        if (out.g) io.g = (double*)(arena + o.g);
        if (out.dg) { if (!out.g) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "dg needs g"); io.dg = (double*)(arena + o.dg); }
        if (out.rho) io.rho = (double*)(arena + o.rho);
        if (out.drho) io.drho = (double*)(arena + o.drho);
        if (out.naccept) io.naccept = (int32_t*)(arena + o.nacc);
        if (out.nreject) io.nreject = (int32_t*)(arena + o.nrej);
        if (out.flags) io.flags = (int32_t*)(arena + o.flags);
        CU(ctx, gb200_launch_dual(P, io, ctx->pool_streams[(size_t)(b % nstreams)]));
        ctx->stats.launches += 1;
        ctx->stats.rays += d.n;
    }
    for (int sidx = 0; sidx < nstreams; ++sidx) {
        CU(ctx, cudaEventRecord(ctx->ev2, ctx->pool_streams[(size_t)sidx]));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev2, 0));
    }
    if (total > in_bytes) CU(ctx, cudaMemcpyAsync(stage + in_bytes, arena + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < nbatch; ++b) {
        const gb200_dual_out& out = outs[b];
        const Off& o = off[(size_t)b];
        const size_t n = (size_t)ics[b].n, nd = (size_t)ics[b].npartials;
        if (!n) continue;
        if (out.status) memcpy(out.status, stage + o.status, 4 * n);
        if (out.lambda_max) memcpy(out.lambda_max, stage + o.lambda, 8 * n);
        for (int k = 0; k < 4; ++k) {
            if (out.x[k]) memcpy(out.x[k], stage + o.x[k], 8 * n);
            if (out.v[k]) memcpy(out.v[k], stage + o.v[k], 8 * n);
        }
        if (out.g) memcpy(out.g, stage + o.g, 8 * n);
        if (out.dg) memcpy(out.dg, stage + o.dg, 8 * n * nd);
        if (out.rho) memcpy(out.rho, stage + o.rho, 8 * n);
        if (out.drho) memcpy(out.drho, stage + o.drho, 8 * n * nd);
        if (out.naccept) memcpy(out.naccept, stage + o.nacc, 4 * n);
        if (out.nreject) memcpy(out.nreject, stage + o.nrej, 4 * n);
        if (out.flags) memcpy(out.flags, stage + o.flags, 4 * n);
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, ctx->ev0, ctx->ev3);
    ctx->stats.kernel_ms = tot; ctx->stats.total_ms = tot;
    return GB200_OK;
}

int gb200_trace_dual(gb200_ctx* ctx, const gb200_problem* p, const gb200_dual_ic* ic, int32_t norm_mode,
                     const gb200_plunging_table* plunging, gb200_dual_out* out) {
    if (!p || !ic || !out) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null argument");
    const gb200_plunging_table* pls[1] = {plunging};
    return gb200_trace_dual_batch(ctx, 1, p, ic, norm_mode, pls, out);
}

static int lineprofile_common(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const gb200_emissivity* em,
                              const gb200_plunging_table* pl, const double* bins, int32_t nbins, const gb200_lineprofile_opts* opts,
                              double* flux, bool device_out, cudaStream_t stream, int async) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    int rc = validate(ctx, p, ic); if (rc) return rc;
    rc = validate_range(ctx, ic, rg); if (rc) return rc;
    if (!em || !bins || !opts || !flux || nbins < 1 || nbins > 8192) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad line-profile arguments (1 <= nbins <= 8192)");
    for (int b = 1; b < nbins; ++b) if (!(bins[b] > bins[b - 1])) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bins must be strictly increasing");
    { int rc_ = begin_call(ctx, stream); if (rc_) return rc_; }
    CU(ctx, cudaEventRecord(ctx->ev0, stream));
    GbParams P;
    fill_params(p, ic, rg, P);
    { int rcg_ = bind_geometry(ctx, P); if (rcg_) return rcg_; }
    rc = set_isco(ctx, p, P); if (rc) return rc;
    rc = upload_ic_and_tables(ctx, ic, pl, em, P); if (rc) return rc;
    P.min_re = opts->min_re; P.max_re = opts->max_re;
    const size_t n = (size_t)rg->count;
    void* d;
    const void* dbins;
    rc = upload(ctx, SL_BINS, bins, sizeof(double) * nbins, &dbins); if (rc) return rc;
    double* dflux = flux;
    if (!device_out) { rc = pool_get(ctx, SL_FLUX, sizeof(double) * nbins, &d); if (rc) return rc; dflux = (double*)d; }
    // Fused path: the histogram is accumulated inside the trace kernel, per CTA in shared memory, as 128-bit fixed point
    // (gb_hist_add).  The scale is a power of two chosen from a generous bound on sum f = sum eps(rho) g^3 area
    // (g <= 1e3), so f * scale is exact and the bins hold the exact sum of the contributions, whatever the order.
    double eps_max = 0.0;
    if (em->kind == GB200_EMISSIVITY_POWERLAW) eps_max = std::max(std::pow(opts->min_re, -em->index), std::pow(opts->max_re, -em->index));
    else for (int k = 0; k < em->n; ++k) eps_max = std::max(eps_max, std::fabs(em->eps[k]));
    const double area_max = (ic->kind == GB200_IC_POLAR_PLANE) ? ic->hi0 * ic->hi0 : 1.0;
    const double bound = (double)(n > 0 ? n : 1) * eps_max * 1e9 * area_max;
    const bool fused = nbins <= 512 && std::isfinite(bound) && bound > 0.0 && !getenv("GB200_HIST_TWO_PASS");
    if (fused) {
        rc = pool_get(ctx, SL_PARTIAL, sizeof(unsigned long long) * 2 * (size_t)nbins, &d); if (rc) return rc;
        P.lp_acc = (unsigned long long*)d;
        P.lp_bins = (const double*)dbins; P.lp_nbins = nbins; P.lp_right_closed = opts->bin_right_closed;
        P.lp_scale = std::exp2(std::floor(120.0 - std::log2(bound)));
        CU(ctx, cudaMemsetAsync(P.lp_acc, 0, sizeof(unsigned long long) * 2 * (size_t)nbins, stream));
        rc = run_trace(ctx, P, stream, true); if (rc) return rc;
        CU(ctx, gb200_launch_hist128_finish(P.lp_acc, nbins, P.lp_scale, dflux, stream));
        ctx->stats.launches += 1;
    } else { // more bins than fit beside the stage data in shared memory: (g, f) per ray through HBM, then the binning kernels
        rc = pool_get(ctx, SL_G, n * sizeof(double) + 8, &d); if (rc) return rc; P.o_g = (double*)d;
        rc = pool_get(ctx, SL_F, n * sizeof(double) + 8, &d); if (rc) return rc; P.o_f = (double*)d;
        const int nblocks = ctx->sm_count * 4;
        rc = pool_get(ctx, SL_PARTIAL, sizeof(double) * (size_t)nblocks * nbins, &d); if (rc) return rc;
        double* dpartial = (double*)d;
        rc = run_trace(ctx, P, stream, true); if (rc) return rc;
        CU(ctx, gb200_launch_hist(P.o_g, P.o_f, rg->count, (const double*)dbins, nbins, opts->bin_right_closed, dpartial, nblocks, dflux, stream));
        ctx->stats.launches += 2;
    }
    if (!device_out) {
        std::vector<double> h((size_t)nbins);
        CU(ctx, cudaMemcpyAsync(h.data(), dflux, sizeof(double) * nbins, cudaMemcpyDeviceToHost, stream));
        CU(ctx, cudaStreamSynchronize(stream));
        double tot = 0;
        for (int b = 0; b < nbins; ++b) tot += h[(size_t)b];
        for (int b = 0; b < nbins; ++b) flux[b] = opts->normalise ? h[(size_t)b] / tot : h[(size_t)b]; // flux ./ sum(flux), line-profiles.jl:197
    }
    if (async) { ctx->stats.rays = rg->count; return end_async_call(ctx, stream); }
    return finish_stats(ctx, rg->count);
}

int gb200_lineprofile(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const gb200_emissivity* em,
                      const gb200_plunging_table* pl, const double* bins, int32_t nbins, const gb200_lineprofile_opts* opts, double* flux_out) {
    return lineprofile_common(ctx, p, ic, rg, em, pl, bins, nbins, opts, flux_out, false, ctx ? ctx->stream : nullptr, 0);
}

int gb200_lineprofile_device(gb200_ctx* ctx, const gb200_problem* p, const gb200_ic* ic, const gb200_range* rg, const gb200_emissivity* em,
                             const gb200_plunging_table* pl, const double* bins, int32_t nbins, const gb200_lineprofile_opts* opts,
                             double* d_flux, void* cuda_stream, int async) {
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : (ctx ? ctx->stream : nullptr);
    return lineprofile_common(ctx, p, ic, rg, em, pl, bins, nbins, opts, d_flux, true, s, async);
}

// ---------------------------------------------------------------- single-process multi-GPU (what a Julia caller uses)
// NCCL is opened at run time (dlopen("libnccl.so.2")): a process that already holds an NCCL (PyTorch bundles one) shares it,
// and a single-GPU caller needs none at all.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string* why) {
        if (handle) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { *why = dlerror() ? dlerror() : "libnccl.so.2 not found"; return false; }
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) { *why = "libnccl lacks a required symbol"; return false; }
        return true;
    }
};
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

struct gb200_comm {
    std::vector<gb200_ctx*> ctx;
    std::vector<ncclComm_t> nccl; // empty when ndev == 1 and NCCL is unavailable
    std::vector<double*> d_flux;  // per device: nbins doubles
    int flux_cap = 0;
};
#define NC(call)                                                                                                         \
    do {                                                                                                                 \
        ncclResult_t r__ = (call);                                                                                       \
        if (r__ != ncclSuccess) return fail(nullptr, GB200_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString(r__));       \
    } while (0)

// rank d's share of the rays: whole strips of four image columns / theta-rows, strip d, d + n, ...
static gb200_range shard_range(const gb200_ic* ic, int d, int n) {
    int64_t h = 0;
    if (ic->kind == GB200_IC_RENDER_GRID) h = ic->height;
    else if (ic->kind == GB200_IC_POLAR_PLANE) h = ic->width;
    else if (ic->kind == GB200_IC_CARTESIAN_PLANE) h = 2 * (ic->height / 2) - 1;
    const int64_t strip = GB_TILE_C * h;
    if (n == 1) return gb200_range{0, ic->n, 1, 1};
    if (strip > 0 && ic->n % strip == 0) {
        const int64_t nstrips = ic->n / strip;
        const int64_t mine = nstrips > d ? (nstrips - d + n - 1) / n : 0;
        return gb200_range{(int64_t)d * strip, mine * strip, n, strip};
    }
    return gb200_range{d, ic->n > d ? (ic->n - d + n - 1) / n : 0, n, 1};
}

int gb200_comm_init(const int32_t* devices, int32_t ndev, gb200_comm** out) {
    if (!devices || ndev < 1 || !out) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "bad communicator arguments");
    *out = nullptr;
    for (int a = 0; a < ndev; ++a)
        for (int b = a + 1; b < ndev; ++b)
            if (devices[a] == devices[b]) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "device %d listed twice", devices[a]);
    gb200_comm* c = new gb200_comm();
    for (int d = 0; d < ndev; ++d) {
        gb200_ctx* ctx = nullptr;
        int rc = gb200_init(devices[d], &ctx);
        if (rc) { gb200_comm_destroy(c); return rc; }
        c->ctx.push_back(ctx);
    }
    std::string why;
    bool have;
    { std::lock_guard<std::mutex> lock(g_nccl_mutex); have = g_nccl.load(&why); }
    if (!have) {
        if (ndev > 1) { gb200_comm_destroy(c); return fail(nullptr, GB200_ERR_UNSUPPORTED, "several devices need NCCL: %s", why.c_str()); }
    } else {
        c->nccl.resize((size_t)ndev);
        std::vector<int> devs(devices, devices + ndev);
        ncclResult_t r = g_nccl.CommInitAll(c->nccl.data(), ndev, devs.data());
        if (r != ncclSuccess) { c->nccl.clear(); gb200_comm_destroy(c); return fail(nullptr, GB200_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(r)); }
    }
    *out = c;
    return GB200_OK;
}

void gb200_comm_destroy(gb200_comm* c) {
    if (!c) return;
    for (size_t d = 0; d < c->ctx.size(); ++d) {
        if (d < c->d_flux.size() && c->d_flux[d]) { cudaSetDevice(c->ctx[d]->device); cudaFree(c->d_flux[d]); }
        if (d < c->nccl.size() && c->nccl[d]) g_nccl.CommDestroy(c->nccl[d]);
        gb200_destroy(c->ctx[d]);
    }
    delete c;
}

int gb200_comm_size(gb200_comm* c) { return c ? (int)c->ctx.size() : 0; }
gb200_ctx* gb200_comm_context(gb200_comm* c, int32_t i) { return (c && i >= 0 && i < (int)c->ctx.size()) ? c->ctx[(size_t)i] : nullptr; }

int gb200_comm_lineprofile(gb200_comm* c, const gb200_problem* p, const gb200_ic* ic, const gb200_emissivity* em, const gb200_plunging_table* pl,
                           const double* bins, int32_t nbins, const gb200_lineprofile_opts* opts, double* flux_out) {
    if (!c || !ic || !opts || !flux_out) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null argument");
    const int n = (int)c->ctx.size();
    if (c->flux_cap < nbins) {
        c->d_flux.resize((size_t)n, nullptr);
        for (int d = 0; d < n; ++d) {
            CU(c->ctx[(size_t)d], cudaSetDevice(c->ctx[(size_t)d]->device));
            if (c->d_flux[(size_t)d]) cudaFree(c->d_flux[(size_t)d]);
            CU(c->ctx[(size_t)d], cudaMalloc(&c->d_flux[(size_t)d], sizeof(double) * (size_t)nbins));
        }
        c->flux_cap = nbins;
    }
    // every device traces its strips and bins them into its own raw histogram (asynchronous launches, one host thread) ...
    for (int d = 0; d < n; ++d) {
        gb200_ctx* ctx = c->ctx[(size_t)d];
        gb200_range rg = shard_range(ic, d, n);
        int rc = lineprofile_common(ctx, p, ic, &rg, em, pl, bins, nbins, opts, c->d_flux[(size_t)d], true, ctx->stream, 1);
        if (rc) return fail(nullptr, rc, "device %d: %s", ctx->device, ctx->err.c_str());
    }
    // ... one all-reduce over NVLink sums them in place on every device's own stream ...
    if (n > 1) {
        NC(g_nccl.GroupStart());
        for (int d = 0; d < n; ++d) {
            ncclResult_t r = g_nccl.AllReduce(c->d_flux[(size_t)d], c->d_flux[(size_t)d], (size_t)nbins, ncclDouble, ncclSum, c->nccl[(size_t)d], c->ctx[(size_t)d]->stream);
            if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(nullptr, GB200_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(r)); }
        }
        NC(g_nccl.GroupEnd());
    }
    // ... and device 0 hands the sum to the host, which normalises it (flux ./ sum(flux), line-profiles.jl:197)
    std::vector<double> h((size_t)nbins);
    CU(c->ctx[0], cudaSetDevice(c->ctx[0]->device));
    CU(c->ctx[0], cudaMemcpyAsync(h.data(), c->d_flux[0], sizeof(double) * (size_t)nbins, cudaMemcpyDeviceToHost, c->ctx[0]->stream));
    for (int d = 0; d < n; ++d) {
        CU(c->ctx[(size_t)d], cudaSetDevice(c->ctx[(size_t)d]->device));
        CU(c->ctx[(size_t)d], cudaStreamSynchronize(c->ctx[(size_t)d]->stream));
    }
    double tot = 0;
    for (int b = 0; b < nbins; ++b) tot += h[(size_t)b];
    for (int b = 0; b < nbins; ++b) flux_out[b] = opts->normalise ? h[(size_t)b] / tot : h[(size_t)b];
    return GB200_OK;
}

int gb200_comm_render(gb200_comm* c, const gb200_problem* p, const gb200_ic* ic, const int32_t* pfs, int32_t npf,
                      const gb200_plunging_table* pl, double* const* images) {
    if (!c || !ic || !pfs || !images || npf < 1 || npf > GB_MAX_PF) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "bad render arguments");
    const int n = (int)c->ctx.size();
    const int K = 4; // chunks per device: a chunk's images travel to the host while the next chunks compute
    std::vector<gb200_range> rgs((size_t)n);
    std::vector<std::vector<double*>> dimg((size_t)n);
    std::vector<std::vector<cudaEvent_t>> done((size_t)n);
    std::vector<int64_t> per((size_t)n, 0);
    auto cleanup = [&]() { for (auto& v : done) for (auto e : v) cudaEventDestroy(e); };
    // all launches first (asynchronous, one host thread): the devices compute concurrently
    for (int d = 0; d < n; ++d) {
        gb200_ctx* ctx = c->ctx[(size_t)d];
        gb200_range& rg = rgs[(size_t)d];
        rg = shard_range(ic, d, n);
        if (rg.block < 1) rg.block = 1;
        CU(ctx, cudaSetDevice(ctx->device));
        while (ctx->pool_streams.size() < 2) {
            cudaStream_t st;
            CU(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            ctx->pool_streams.push_back(st);
        }
        dimg[(size_t)d].resize((size_t)npf);
        for (int k = 0; k < npf; ++k) {
            void* dv = nullptr;
            int rc = pool_get(ctx, SL_IMG0 + k, sizeof(double) * (size_t)rg.count + 8, &dv);
            if (rc) { cleanup(); return fail(nullptr, rc, "device %d: %s", ctx->device, ctx->err.c_str()); }
            dimg[(size_t)d][(size_t)k] = (double*)dv;
        }
        if (rg.count == 0) continue;
        const int64_t nblk = rg.count / rg.block; // strips (or single rays in the fallback decomposition)
        const int chunks = nblk >= 2 * K ? K : 1;
        per[(size_t)d] = ((nblk + chunks - 1) / chunks) * rg.block;
        for (int64_t s0 = 0; s0 < rg.count; s0 += per[(size_t)d]) {
            gb200_range sub{rg.first + (s0 / rg.block) * rg.stride * rg.block, std::min(per[(size_t)d], rg.count - s0), rg.stride, rg.block};
            double* sub_img[GB_MAX_PF];
            for (int k = 0; k < npf; ++k) sub_img[k] = dimg[(size_t)d][(size_t)k] + s0;
            int rc = render_common(ctx, p, ic, &sub, pfs, npf, pl, sub_img, true, ctx->stream, 1);
            if (rc) { cleanup(); return fail(nullptr, rc, "device %d: %s", ctx->device, ctx->err.c_str()); }
            cudaEvent_t ev;
            CU(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            done[(size_t)d].push_back(ev);
            CU(ctx, cudaEventRecord(ev, ctx->stream));
        }
    }
    // Results travel chunk by chunk: a contiguous copy into the device's pinned staging buffer on its copy stream (truly
    // asynchronous, every device on its own PCIe link), then one host thread per device scatters the chunk's strips into
    // the caller's (pageable) images -- strip b of a device sits at ray first + b * stride * block.  A strided copy
    // straight into pageable memory is staged by the driver on the calling thread and serialises the devices (measured:
    // 8 GPUs 15.1 ms for the 2048 x 2048 render whose kernels take 5.3 ms).
    std::vector<std::vector<cudaEvent_t>> copied((size_t)n);
    auto cleanup2 = [&]() { for (auto& v : copied) for (auto e : v) cudaEventDestroy(e); };
    for (int d = 0; d < n; ++d) {
        gb200_ctx* ctx = c->ctx[(size_t)d];
        const gb200_range& rg = rgs[(size_t)d];
        if (rg.count == 0) continue;
        CU(ctx, cudaSetDevice(ctx->device));
        const size_t total = sizeof(double) * (size_t)rg.count * (size_t)npf;
        if (ctx->stage_cap < total) {
            if (ctx->stage) cudaFreeHost(ctx->stage);
            ctx->stage = nullptr; ctx->stage_cap = 0;
            CU(ctx, cudaMallocHost(&ctx->stage, total + 256));
            ctx->stage_cap = total;
        }
    }
    for (int ch = 0; ch < K; ++ch)
        for (int d = 0; d < n; ++d) {
            if ((size_t)ch >= done[(size_t)d].size()) continue;
            gb200_ctx* ctx = c->ctx[(size_t)d];
            const gb200_range& rg = rgs[(size_t)d];
            const int64_t s0 = (int64_t)ch * per[(size_t)d], cnt = std::min(per[(size_t)d], rg.count - s0);
            CU(ctx, cudaSetDevice(ctx->device));
            cudaStream_t copy = ctx->pool_streams[1];
            CU(ctx, cudaStreamWaitEvent(copy, done[(size_t)d][(size_t)ch], 0));
            for (int k = 0; k < npf; ++k)
                CU(ctx, cudaMemcpyAsync((double*)ctx->stage + (size_t)k * (size_t)rg.count + s0, dimg[(size_t)d][(size_t)k] + s0,
                                        sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, copy));
            cudaEvent_t ev;
            CU(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            copied[(size_t)d].push_back(ev);
            CU(ctx, cudaEventRecord(ev, copy));
        }
    std::vector<cudaError_t> terr((size_t)n, cudaSuccess);
    std::vector<std::thread> workers;
    for (int d = 0; d < n; ++d) {
        if (copied[(size_t)d].empty()) continue;
        workers.emplace_back([&, d]() {
            gb200_ctx* ctx = c->ctx[(size_t)d];
            const gb200_range& rg = rgs[(size_t)d];
            const size_t blk = (size_t)rg.block;
            for (size_t ch = 0; ch < copied[(size_t)d].size(); ++ch) {
                const cudaError_t e = cudaEventSynchronize(copied[(size_t)d][ch]);
                if (e != cudaSuccess) { terr[(size_t)d] = e; return; }
                const int64_t s0 = (int64_t)ch * per[(size_t)d], cnt = std::min(per[(size_t)d], rg.count - s0);
                for (int k = 0; k < npf; ++k) {
                    const double* src = (const double*)ctx->stage + (size_t)k * (size_t)rg.count + s0;
                    double* dst = images[k] + rg.first + (s0 / rg.block) * rg.stride * rg.block;
                    if (rg.stride == 1) memcpy(dst, src, sizeof(double) * (size_t)cnt);
                    else
                        for (int64_t b_ = 0; b_ < cnt / rg.block; ++b_)
                            memcpy(dst + (size_t)b_ * blk * (size_t)rg.stride, src + (size_t)b_ * blk, sizeof(double) * blk);
                }
            }
        });
    }
    for (auto& w : workers) w.join();
    for (int d = 0; d < n; ++d) {
        gb200_ctx* ctx = c->ctx[(size_t)d];
        if (terr[(size_t)d] != cudaSuccess) { cleanup(); cleanup2(); return fail(nullptr, GB200_ERR_CUDA, "device %d: %s", ctx->device, cudaGetErrorString(terr[(size_t)d])); }
        CU(ctx, cudaSetDevice(ctx->device));
        if (ctx->pool_streams.size() >= 2) CU(ctx, cudaStreamSynchronize(ctx->pool_streams[1]));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cleanup2();
    cleanup();
    return GB200_OK;
}

int gb200_bucket2d(gb200_ctx* ctx, int64_t n, const double* x, const double* y, const double* w, const double* xbins, int32_t nx,
                   const double* ybins, int32_t ny, double* out) {
    if (!ctx) return fail(nullptr, GB200_ERR_INVALID_ARGUMENT, "null context");
    if (n < 0 || (n > 0 && (!x || !y || !w)) || !xbins || !ybins || !out || nx < 1 || ny < 1 || (int64_t)nx * ny > (1 << 26))
        return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad bucket arguments");
    for (int b = 1; b < nx; ++b) if (!(xbins[b] > xbins[b - 1])) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bins must be strictly increasing");
    for (int b = 1; b < ny; ++b) if (!(ybins[b] > ybins[b - 1])) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bins must be strictly increasing");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    long double tot = 0.0L;
    for (int64_t i = 0; i < n; ++i) if (w[i] == w[i]) tot += fabsl((long double)w[i]);
    const double scale = tot > 0.0L ? (double)(4611686018427387904.0L / tot) : 1.0; // 2^62 / sum |w|
    void *dx, *dy, *dw, *dxb, *dyb, *dacc, *dout;
    const size_t nb = sizeof(double) * (size_t)(n > 0 ? n : 1), ncell = (size_t)nx * ny;
    int rc = pool_get(ctx, SL_G, nb, &dx); if (rc) return rc;
    rc = pool_get(ctx, SL_F, nb, &dy); if (rc) return rc;
    rc = pool_get(ctx, SL_X0, nb, &dw); if (rc) return rc;
    rc = pool_get(ctx, SL_BINS, sizeof(double) * (size_t)(nx + ny), &dxb); if (rc) return rc;
    dyb = (double*)dxb + nx;
    rc = pool_get(ctx, SL_PARTIAL, sizeof(long long) * ncell, &dacc); if (rc) return rc;
    rc = pool_get(ctx, SL_FLUX, sizeof(double) * ncell, &dout); if (rc) return rc;
    if (n > 0) {
        CU(ctx, cudaMemcpyAsync(dx, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dy, y, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dw, w, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(ctx, cudaMemcpyAsync(dxb, xbins, sizeof(double) * (size_t)nx, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(dyb, ybins, sizeof(double) * (size_t)ny, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, gb200_launch_bucket2d((const double*)dx, (const double*)dy, (const double*)dw, n, (const double*)dxb, nx, (const double*)dyb, ny, scale,
                                  (long long*)dacc, (double*)dout, ctx->sm_count * 8, ctx->stream));
    ctx->stats.launches = 2;
    CU(ctx, cudaMemcpyAsync(out, dout, sizeof(double) * ncell, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return GB200_OK;
}

int gb200_debug_rhs(gb200_ctx* ctx, int32_t metric_kind, const double* mp, int64_t n, const double* u, double* du) {
    if (!ctx || !mp || !u || !du || n < 1) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad arguments");
    if (metric_kind < 0 || metric_kind >= GB200_METRIC_COUNT) return fail(ctx, GB200_ERR_UNSUPPORTED, "metric kind %d", metric_kind);
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    GbParams P;
    memset(&P, 0, sizeof P);
    P.metric_kind = metric_kind; P.M = mp[0]; P.a = mp[1]; P.eps3 = mp[2];
    P.a2 = P.a * P.a; P.twoM = 2.0 * P.M; P.jp_e = P.eps3 * P.M * P.M * P.M;
    for (int k = 0; k < 8; ++k) P.mp[k] = mp[k];
    void *d_u, *d_du;
    int rc = pool_get(ctx, SL_X0, sizeof(double) * 8 * (size_t)n, &d_u); if (rc) return rc;
    rc = pool_get(ctx, SL_V0, sizeof(double) * 8 * (size_t)n, &d_du); if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d_u, u, sizeof(double) * 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, gb200_launch_debug_rhs(P, n, (const double*)d_u, (double*)d_du, ctx->stream));
    CU(ctx, cudaMemcpyAsync(du, d_du, sizeof(double) * 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return GB200_OK;
}

int gb200_debug_math(gb200_ctx* ctx, int64_t n, const double* x, double* out5) {
    if (!ctx || !x || !out5 || n < 1) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad arguments");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    void *d_x, *d_o;
    int rc = pool_get(ctx, SL_X0, sizeof(double) * (size_t)n, &d_x); if (rc) return rc;
    rc = pool_get(ctx, SL_V0, sizeof(double) * 5 * (size_t)n, &d_o); if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, gb200_launch_debug_math(n, (const double*)d_x, (double*)d_o, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out5, d_o, sizeof(double) * 5 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return GB200_OK;
}

int gb200_debug_math_lo(gb200_ctx* ctx, int64_t n, const double* x, double* out2) {
    if (!ctx || !x || !out2 || n < 1) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad arguments");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    void *d_x, *d_o;
    int rc = pool_get(ctx, SL_X0, sizeof(double) * (size_t)n, &d_x); if (rc) return rc;
    rc = pool_get(ctx, SL_V0, sizeof(double) * 2 * (size_t)n, &d_o); if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, gb200_launch_debug_math_lo(n, (const double*)d_x, (double*)d_o, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out2, d_o, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return GB200_OK;
}

int gb200_fp64_peak(gb200_ctx* ctx, double* tflops_out) {
    if (!ctx || !tflops_out) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "null argument");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    void* d;
    int rc = pool_get(ctx, SL_SCRATCH, 64, &d); if (rc) return rc;
    const int blocks = ctx->sm_count * 8, iters = 20000;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU(ctx, gb200_launch_dfma((double*)d, blocks, iters, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2);
        const double flop = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks; // 64 FMAs per iteration per thread
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->fp64_peak = best;
    *tflops_out = best;
    return GB200_OK;
}

int gb200_fp64_issue_probe(gb200_ctx* ctx, int32_t mix, double* tflops_out) {
    if (!ctx || !tflops_out || mix < 0 || mix > 5) return fail(ctx, GB200_ERR_INVALID_ARGUMENT, "bad argument");
    { int rc_ = begin_call(ctx, ctx->stream); if (rc_) return rc_; }
    void* d;
    int rc = pool_get(ctx, SL_SCRATCH, 64, &d); if (rc) return rc;
    const int blocks = ctx->sm_count * 8, iters = 20000;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU(ctx, gb200_launch_dfma_mix((double*)d, blocks, iters, mix, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2);
        const double tf = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops_out = best;
    return GB200_OK;
}

} // extern "C"
