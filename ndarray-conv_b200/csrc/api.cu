// api.cu -- C ABI of libndconv_cuda.so (include/ndconv.h): processor handle, caches, planning, launches.
//
// Built by nvcc for sm_100a (the product).  tests/emul/ compiles the same translation unit with
// -DNDCONV_HOST_EMUL as plain C++ to check the kernel bodies' index logic on a GPU-less box; that
// build reports ndconv_is_emulation() == 1 and is refused by the product loader.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_logic.h"
#include "kernels_direct.h"
#include "kernels_fft.h"
#include "kernels_fft_fast.cuh"
#include "kernels_direct_tile.cuh"

#ifdef NDCONV_CUDA
#include <cuda_runtime.h>
#endif

using namespace ndc;

#define NDCONV_VERSION_STRING "ndconv-b200 0.1.0 (sm_100a)"

// ======================================================================================================
// backend: CUDA runtime, or (tests only) host emulation
// ======================================================================================================
#ifdef NDCONV_CUDA
typedef cudaStream_t stream_t;
#define CU_CHECK(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);              \
            return NDCONV_ERR_CUDA;                                                                      \
        }                                                                                                \
    } while (0)

// cudaFuncSetAttribute configures a kernel on the CURRENT device only: remember, per call site, which devices have been
// configured (a process may hold processors on several GPUs, and ndconv_conv_fft_sharded drives them from one host thread each)
#define NDC_ONCE_PER_DEVICE(...)                                                                         \
    do {                                                                                                 \
        static std::mutex mu_;                                                                           \
        static uint64_t done_ = 0;                                                                       \
        int dev_ = 0;                                                                                    \
        cudaGetDevice(&dev_);                                                                            \
        std::lock_guard<std::mutex> lk_(mu_);                                                            \
        if (!((done_ >> (dev_ & 63)) & 1)) { __VA_ARGS__; done_ |= 1ull << (dev_ & 63); }                \
    } while (0)

template <class Body, class Params> __global__ void __launch_bounds__(512) kentry(const __grid_constant__ Params p)
{
    extern __shared__ __align__(16) unsigned char ndc_smem[];
    BlockCtx c;
    c.tid = threadIdx.x; c.nt = blockDim.x; c.bid = blockIdx.x; c.nb = gridDim.x; c.smem = (char *)ndc_smem;
    Body::run(c, p);
}

struct ProfRec { const char *name; cudaEvent_t a, b; double bytes; };
struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
};
struct LaunchCtx { stream_t st; int64_t *counter; Profiler *prof; };

template <class Body, class Params>
static int launch(const LaunchCtx &lc, const char *name, double alg_bytes, int64_t grid, int block, size_t smem, const Params &p)
{
    NDC_ONCE_PER_DEVICE(CU_CHECK(cudaFuncSetAttribute(kentry<Body, Params>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
    if (grid < 1) grid = 1;
    ProfRec rec;
    const bool prof = lc.prof && lc.prof->on;
    if (prof) { rec.name = name; rec.bytes = alg_bytes; rec.a = lc.prof->get(); rec.b = lc.prof->get(); CU_CHECK(cudaEventRecord(rec.a, lc.st)); }
    kentry<Body, Params><<<(unsigned)grid, block, smem, lc.st>>>(p);
    CU_CHECK(cudaGetLastError());
    if (prof) { CU_CHECK(cudaEventRecord(rec.b, lc.st)); lc.prof->recs.push_back(rec); }
    if (lc.counter) (*lc.counter)++;
    return NDCONV_OK;
}
// launch of a plain __global__ kernel (the sm_100a fast path) with the same counting / profiling
template <class F> static int launch_raw(const LaunchCtx &lc, const char *name, double alg_bytes, F &&do_launch)
{
    ProfRec rec;
    const bool prof = lc.prof && lc.prof->on;
    if (prof) { rec.name = name; rec.bytes = alg_bytes; rec.a = lc.prof->get(); rec.b = lc.prof->get(); CU_CHECK(cudaEventRecord(rec.a, lc.st)); }
    do_launch();
    CU_CHECK(cudaGetLastError());
    if (prof) { CU_CHECK(cudaEventRecord(rec.b, lc.st)); lc.prof->recs.push_back(rec); }
    if (lc.counter) (*lc.counter)++;
    return NDCONV_OK;
}
static int be_malloc(void **p, size_t n) { CU_CHECK(cudaMalloc(p, n ? n : 1)); return NDCONV_OK; }
static void be_free(void *p) { if (p) cudaFree(p); }
static int be_h2d(void *d, const void *s, size_t n, stream_t st) { CU_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st)); return NDCONV_OK; }
static int be_d2h(void *d, const void *s, size_t n, stream_t st) { CU_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st)); return NDCONV_OK; }
static int be_sync(stream_t st) { CU_CHECK(cudaStreamSynchronize(st)); return NDCONV_OK; }
static const int kMaxGridMult = 8;
static int be_num_sms(int dev) { int n = 148; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
#else
typedef void *stream_t;
#define CU_CHECK(expr) do { } while (0)
struct Profiler { bool on = false; };
struct LaunchCtx { stream_t st; int64_t *counter; Profiler *prof; };
template <class Body, class Params>
static int launch(const LaunchCtx &lc, const char *, double, int64_t grid, int, size_t smem, const Params &p)
{
    int64_t *counter = lc.counter;
    std::vector<unsigned char> sm(smem + 64);
    if (grid < 1) grid = 1;
    // one block walks the whole grid-stride loop; a second "block" exercises the nb > 1 indexing
    int64_t nb = grid > 1 ? 2 : 1;
    for (int64_t b = 0; b < nb; b++) {
        BlockCtx c; c.tid = 0; c.nt = 1; c.bid = b; c.nb = nb; c.smem = (char *)sm.data();
        Body::run(c, p);
    }
    if (counter) (*counter)++;
    return NDCONV_OK;
}
static int be_malloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? NDCONV_OK : NDCONV_ERR_INTERNAL; }
static void be_free(void *p) { free(p); }
static int be_h2d(void *d, const void *s, size_t n, stream_t) { memcpy(d, s, n); return NDCONV_OK; }
static int be_d2h(void *d, const void *s, size_t n, stream_t) { memcpy(d, s, n); return NDCONV_OK; }
static int be_sync(stream_t) { return NDCONV_OK; }
static const int kMaxGridMult = 1;
static int be_num_sms(int) { return 4; }
#endif

// ======================================================================================================
// processor
// ======================================================================================================
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return NDCONV_OK;
        be_free(p); p = nullptr; cap = 0;
        size_t want = n + n / 8;
        int st = be_malloc(&p, want);
        if (st) return st;
        cap = want;
        return NDCONV_OK;
    }
    void release() { be_free(p); p = nullptr; cap = 0; }
};

struct KSpecEntry {
    std::vector<unsigned char> key;
    DevBuf buf;
    DevBuf pair;     // paired layout for the sm_100a 2-D fast path
    uint64_t last_use = 0;
};

struct PlanEntry;
struct PlanEntryDeleter { void operator()(PlanEntry *e) const; };

struct ndconv_processor {
    std::vector<std::unique_ptr<PlanEntry, PlanEntryDeleter>> plans;   // per-geometry cache: validated geometry, device border maps / taps, FFT plan
    int64_t plan_hits = 0, plan_misses = 0;
    int device = 0;
    int num_sms = 148;
    stream_t own_stream = nullptr, stream = nullptr;
    int64_t launches = 0;
    DevBuf ws, in_stage, out_stage, meta, kb_stage, kmeta;
    std::map<std::pair<int, int>, void *> tw_c;   // (L, is_double) -> exp(-2 pi i j/L), j < L
    std::map<std::pair<int, int>, void *> tw_r;   // (F, is_double) -> exp(-2 pi i k/F), k <= F/4 + 1
    std::vector<std::unique_ptr<KSpecEntry>> kspecs;
    uint64_t tick = 0;
    Profiler prof;
    LaunchCtx lc() { return LaunchCtx{stream, &launches, &prof}; }
    // host-path slab pipeline (H2D | kernels | D2H overlapped)
    DevBuf pipe_in[2], pipe_out[2], pipe_row;
    DevBuf zero_map;             // one int32 0: identity border map of the dummy leading axes of the rank-3 tile kernel
    // axis-0 split: the tail part runs on a second stream with a workspace of its own, beside the main part
    DevBuf ws_aux;
    stream_t aux_stream = nullptr;
    bool in_split = false, in_tail = false;  // in_tail: the launches being issued belong to the tail part (profiled under tail:* names)
#ifdef NDCONV_CUDA
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
#endif
    stream_t h2d_stream = nullptr, d2h_stream = nullptr;
#ifdef NDCONV_CUDA
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
#endif
    int64_t pipelined_slabs = 0, bounced_slabs = 0;
    // pageable host arrays: pinned bounce buffers (cudaHostAlloc) the slabs are staged through by several host memcpy threads
    void *bounce_in[2] = {nullptr, nullptr}, *bounce_out[2] = {nullptr, nullptr};
    size_t bounce_in_cap = 0, bounce_out_cap = 0;
    bool l2_limit_set = false;
    size_t held() const
    {
        size_t s = ws.cap + in_stage.cap + out_stage.cap + meta.cap + kb_stage.cap + kmeta.cap + pipe_in[0].cap + pipe_in[1].cap + pipe_out[0].cap + pipe_out[1].cap;
        for (auto &k : kspecs) s += k->buf.cap + k->pair.cap;
        return s;
    }
};

static int set_device(const ndconv_processor *p)
{
#ifdef NDCONV_CUDA
    CU_CHECK(cudaSetDevice(p->device));
#else
    (void)p;
#endif
    return NDCONV_OK;
}

template <class R> static int get_tw_c(ndconv_processor *p, int L, const cx<R> **out)
{
    auto key = std::make_pair(L, (int)(sizeof(R) == 8));
    auto it = p->tw_c.find(key);
    if (it == p->tw_c.end()) {
        std::vector<cx<R>> h((size_t)std::max(L, 1));
        for (int j = 0; j < L; j++) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)L;
            h[j].re = (R)cosl(a); h[j].im = (R)sinl(a);
        }
        void *d = nullptr;
        int st = be_malloc(&d, h.size() * sizeof(cx<R>)); if (st) return st;
        st = be_h2d(d, h.data(), h.size() * sizeof(cx<R>), p->stream); if (st) return st;
        st = be_sync(p->stream); if (st) return st;
        it = p->tw_c.emplace(key, d).first;
    }
    *out = (const cx<R> *)it->second;
    return NDCONV_OK;
}
template <class R> static int get_tw_r(ndconv_processor *p, int F, const cx<R> **out)
{
    auto key = std::make_pair(F, (int)(sizeof(R) == 8));
    auto it = p->tw_r.find(key);
    if (it == p->tw_r.end()) {
        int cnt = F / 4 + 2;
        std::vector<cx<R>> h((size_t)cnt);
        for (int k = 0; k < cnt; k++) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)F;
            h[k].re = (R)cosl(a); h[k].im = (R)sinl(a);
        }
        void *d = nullptr;
        int st = be_malloc(&d, h.size() * sizeof(cx<R>)); if (st) return st;
        st = be_h2d(d, h.data(), h.size() * sizeof(cx<R>), p->stream); if (st) return st;
        st = be_sync(p->stream); if (st) return st;
        it = p->tw_r.emplace(key, d).first;
    }
    *out = (const cx<R> *)it->second;
    return NDCONV_OK;
}

// pack a strided host array into standard layout
static void pack_strided(const void *src, int ndim, const int64_t *shape, const int64_t *strides, int es, void *dst)
{
    int64_t total = 1;
    for (int i = 0; i < ndim; i++) total *= shape[i];
    int64_t idx[NDC_MAX_DIM] = {0};
    const unsigned char *s = (const unsigned char *)src;
    unsigned char *d = (unsigned char *)dst;
    for (int64_t e = 0; e < total; e++) {
        int64_t o = 0;
        for (int i = 0; i < ndim; i++) o += idx[i] * strides[i];
        memcpy(d + e * es, s + o * es, es);
        for (int i = ndim - 1; i >= 0; i--) { if (++idx[i] < shape[i]) break; idx[i] = 0; }
    }
}

// device copies of the per-axis border maps (+ optional tap tables) packed into one buffer
struct MetaLayout {
    size_t map_off[NDC_MAX_DIM];
    size_t tap_off_off = 0, tap_lin_off = 0, tap_w_off = 0, total = 0;
};
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int upload_meta(ndconv_processor *p, DevBuf &buf, const Geom &g, const std::vector<int32_t> *maps, const Taps *taps, MetaLayout *ml)
{
    size_t off = 0;
    for (int a = 0; a < g.ndim; a++) { ml->map_off[a] = off; off = align_up(off + maps[a].size() * sizeof(int32_t), 16); }
    if (taps) {
        ml->tap_off_off = off; off = align_up(off + taps->off.size() * sizeof(int32_t), 16);
        ml->tap_lin_off = off; off = align_up(off + taps->lin.size() * sizeof(int64_t), 16);
        ml->tap_w_off = off; off = align_up(off + taps->w.size(), 16);
    }
    ml->total = off;
    std::vector<unsigned char> h(off ? off : 16, 0);
    for (int a = 0; a < g.ndim; a++) memcpy(h.data() + ml->map_off[a], maps[a].data(), maps[a].size() * sizeof(int32_t));
    if (taps && taps->ntap) {
        memcpy(h.data() + ml->tap_off_off, taps->off.data(), taps->off.size() * sizeof(int32_t));
        memcpy(h.data() + ml->tap_lin_off, taps->lin.data(), taps->lin.size() * sizeof(int64_t));
        memcpy(h.data() + ml->tap_w_off, taps->w.data(), taps->w.size());
    }
    int st = buf.reserve(h.size()); if (st) return st;
    return be_h2d(buf.p, h.data(), h.size(), p->stream);
}

static void fill_consts(const ndconv_problem *pr, int ndim, unsigned char cf[][16], unsigned char cb[][16])
{
    for (int a = 0; a < ndim; a++) {
        memset(cf[a], 0, 16); memset(cb[a], 0, 16);
        if (pr->border[a][0].type == NDCONV_BORDER_CONST) memcpy(cf[a], pr->border[a][0].value, 16);
        if (pr->border[a][1].type == NDCONV_BORDER_CONST) memcpy(cb[a], pr->border[a][1].value, 16);
    }
}

// ======================================================================================================
// direct convolution
// ======================================================================================================
template <class T> static int run_direct_t(ndconv_processor *p, const DirectParams &dp, double alg_bytes)
{
    int block = 256;
    int64_t grid = std::min<int64_t>((dp.total + block - 1) / block, (int64_t)p->num_sms * 16 * kMaxGridMult);
    return launch<DirectBody<T>, DirectParams>(p->lc(), "direct_conv", alg_bytes, grid, block, 0, dp);
}

// ======================================================================================================
// FFT convolution
// ======================================================================================================
static const size_t kRowSmemBudget = 96 * 1024;
static const size_t kColSmemBudget = 128 * 1024;
static int cap_last_axis(bool is_cx, bool is_dbl) { return is_cx ? (is_dbl ? 2048 : 4096) : (is_dbl ? 4096 : 8192); }
static int cap_col_axis(bool is_dbl) { return is_dbl ? 512 : 1024; }

struct FftPlan {
    int N = 0;
    bool is_cx = false;
    bool fast = false;            // sm_100a fast path: real f32, rank 2 or 3, power-of-two tiles (kernels_fft_fast.cuh)
    AxisTiling tl[NDC_MAX_DIM];
    FftLen fl[NDC_MAX_DIM];       // complex transform per axis (last axis: F/2 for real input)
    int H = 0, Hp = 0;
    int64_t rows_per_tile = 1, tile_elems = 0, ntiles_total = 1;
};

// sm_100a fast path (kernels_fft_fast.cuh): real f32, rank 2 or 3, power-of-two overlap-save tiles
static bool fast_eligible(const Geom &g)
{
#ifdef NDCONV_CUDA
    static const bool disabled = getenv("NDCONV_DISABLE_OPT") != nullptr;
    const bool cx32 = g.dtype == NDCONV_C32;
    if (disabled || (g.dtype != NDCONV_F32 && !cx32) || g.ndim < 1 || g.ndim > 3) return false;
    const int al = g.ndim - 1;
    if (g.P[al] < (cx32 ? 64 : 128) || g.Kd[al] > (cx32 ? 512 : 1024)) return false;
    int64_t tot = 1;
    for (int a = 0; a < g.ndim; a++) { tot *= g.P[a]; if (a < al && g.Kd[a] > 512) return false; }
    if (tot < (g.ndim == 1 ? 512 : 16384)) return false;
    // the row kernels decode work indices in 32 bits: rows (tile overlap inflates by < 2x per axis) x last-axis tiles must fit
    double work = (double)(g.P[al] / 128 + 2);
    for (int a = 0; a < al; a++) work *= 2.0 * (double)g.P[a] + 16.0;
    return work < 2.0e9;
#else
    (void)g;
    return false;
#endif
}
// tile length of one axis from a power-of-two menu: fewest transformed samples, ties to the longer tile (fewer, larger work
// items: less pitch padding).  `slack` >= 0 (small, latency-bound problems): the SHORTEST tile within (1 + slack) of the fewest
// samples instead -- a launch of less than a wave is as slow as its longest warp, and a shorter tile is a shorter second radix
static void fast_pick_tile(int64_t P, int64_t Kd, const int *menu, int nmenu, AxisTiling *out, double slack = -1.0)
{
    for (int pass = 0; pass < 2 && out->F == 0; pass++) {        // pass 0: at least half of every tile useful; pass 1: anything that works
        double best = 1e300;
        for (int i = 0; i < nmenu; i++) {
            const int F = menu[i];
            const int64_t V = F - Kd + 1;
            if (V < 1 || (pass == 0 && 2 * V < F)) continue;
            const int64_t nt = (P - Kd + 1 + V - 1) / V;
            const double c = (double)nt * F;
            if (c <= best) { best = c; out->F = F; out->V = (int)V; out->ntiles = (int)nt; }
        }
        if (slack < 0 || out->F == 0) continue;
        for (int i = 0; i < nmenu; i++) {                        // menus are ascending: the first tile inside the slack is the shortest
            const int F = menu[i];
            const int64_t V = F - Kd + 1;
            if (V < 1 || (pass == 0 && 2 * V < F)) continue;
            const int64_t nt = (P - Kd + 1 + V - 1) / V;
            if ((double)nt * F <= best * (1.0 + slack)) { out->F = F; out->V = (int)V; out->ntiles = (int)nt; break; }
        }
    }
}

static int make_plan(const Geom &g, FftPlan *pl)
{
    const int N = g.ndim;
    const bool is_cx = dtype_is_complex(g.dtype), is_dbl = (g.dtype == NDCONV_F64 || g.dtype == NDCONV_C64);
    pl->N = N; pl->is_cx = is_cx;
    pl->fast = fast_eligible(g);
    for (int a = 0; a < N; a++) {
        if (pl->fast) {
            static const int menu_last[4] = {256, 512, 1024, 2048}, menu_last_cx[4] = {128, 256, 512, 1024}, menu_col[7] = {16, 32, 64, 128, 256, 512, 1024};
            pl->tl[a].F = 0;
            // problems of fewer than 1.28 M padded samples are well under one wave of row warps: the shortest tile within 10 % of the
            // fewest samples (measured: c2 31.0 -> 27.3 us, c3 36.9 -> 33.3, c3 Complex 32.0 -> 30.7; 25 %: c2 26.6, c3 35.8 / 34.7;
            // forced on larger problems it loses: 64x200x200 64 -> 77 us, 3000^2 96 -> 108, 8192^2 478 -> 576; tools/run_midsize.py).
            // NDCONV_TILE_SLACK=<percent> (negative: off) and NDCONV_TILE_SMALL_K=<thousand samples> override for experiments
            static const double slack_env = getenv("NDCONV_TILE_SLACK") ? atof(getenv("NDCONV_TILE_SLACK")) / 100.0 : 0.10;
            static const double small_k = getenv("NDCONV_TILE_SMALL_K") ? atof(getenv("NDCONV_TILE_SMALL_K")) : 1280.0;
            double tot = 1; for (int b = 0; b < N; b++) tot *= (double)g.P[b];
            const double slack = tot < small_k * 1000.0 ? slack_env : -1.0;
            if (a == N - 1) fast_pick_tile(g.P[a], g.Kd[a], is_cx ? menu_last_cx : menu_last, 4, &pl->tl[a], slack);
            else fast_pick_tile(g.P[a], g.Kd[a], menu_col, 7, &pl->tl[a], slack);
            if (pl->tl[a].F == 0) { pl->fast = false; a = -1; continue; }      // no usable tile: replan everything on the generic path
            if (!factor_radices(a == N - 1 && !is_cx ? pl->tl[a].F / 2 : pl->tl[a].F, &pl->fl[a])) return NDCONV_ERR_INTERNAL;
            continue;
        }
        const bool last = (a == N - 1);
        const bool real_axis = last && !is_cx;
        int cap = last ? cap_last_axis(is_cx, is_dbl) : cap_col_axis(is_dbl);
        if (N == 1) {
            // a 1-D problem is one row: shorter overlap-save tiles = more CTAs in flight and fewer Stockham passes each
            int c1 = 512;
            while (c1 < 8 * g.Kd[a] && c1 < cap) c1 <<= 1;
            cap = std::min(cap, c1);
        }
        int st = plan_axis(g.P[a], g.Kd[a], cap, real_axis, &pl->tl[a]); if (st) return st;
        int L = real_axis ? pl->tl[a].F / 2 : pl->tl[a].F;
        if (!factor_radices(L, &pl->fl[a], is_dbl ? 16 : 32)) { set_error("internal: non-smooth FFT length"); return NDCONV_ERR_INTERNAL; }
    }
    const int Fl = pl->tl[N - 1].F;
    pl->H = is_cx ? Fl : Fl / 2 + 1;
    pl->Hp = (int)align_up((size_t)pl->H, 16);
    pl->rows_per_tile = 1; pl->ntiles_total = 1;
    for (int a = 0; a < N - 1; a++) pl->rows_per_tile *= pl->tl[a].F;
    for (int a = 0; a < N; a++) pl->ntiles_total *= pl->tl[a].ntiles;
    pl->tile_elems = pl->rows_per_tile * pl->Hp;
    return NDCONV_OK;
}

// ======================================================================================================
// plan cache
// ======================================================================================================
struct PlanEntry {
    std::vector<unsigned char> key;
    uint64_t last_use = 0;
    Geom g;                      // xstr already normalised to the device-side layout
    bool host_contiguous = true; // host problems: the caller's array is standard layout (no packing needed)
    DevBuf meta;
    MetaLayout ml;
    int ntap = 0;
    FftPlan pl;
    // axis-0 split (plan_axis0_split): outputs [0, split_out) come from a sub-convolution whose tiles fit exactly, the rest
    // from a second one that the planner gives a shorter tile -- instead of a whole last tile that is mostly padding
    int64_t split_out = 0, split_rows_a = 0;   // outputs / input rows of part A
    bool split_concurrent = false;   // run the tail on the second stream (only when it is a sizeable share of the work)
};
void PlanEntryDeleter::operator()(PlanEntry *e) const { if (e) { e->meta.release(); delete e; } }

// Overlap-save along axis 0 works just as well across two launches as across the tiles of one (the identity the multi-GPU
// slabs use).  When the last axis-0 tile of a fast-path plan is mostly padding -- 32830 output rows =
// 34 x 962 + 122 on c5; 4104 = 4 x 962 + 256 per rank on 8 GPUs -- the problem is cut after the last full tile: part A
// tiles exactly, part B gets whatever shorter tile the planner picks for it (both run on the device copy of the input).  A Circular
// border on axis 0 reads from the far end of the array, which a part does not hold, so it is never split.
static void plan_axis0_split(const ndconv_problem *pr, const Geom &g, const FftPlan &pl, PlanEntry *e)
{
    e->split_out = 0;
#ifdef NDCONV_CUDA
    static const bool disabled = getenv("NDCONV_DISABLE_SPLIT") != nullptr;
    if (disabled || !pl.fast || g.ndim < 2) return;
    const AxisTiling &t = pl.tl[0];
    if (t.ntiles < 2) return;
    if ((g.bf[0] == NDCONV_BORDER_CIRCULAR && g.pf[0] > 0) || (g.bb[0] == NDCONV_BORDER_CIRCULAR && g.pb[0] > 0)) return;
    // part A = the first m tile rows.  Two reasons to cut:
    //  (1) workspace budget: one tile row of workspace per tile row of the problem adds up (c5: 5 GB; a 131072^2 array: 80 GB on
    //      top of 64 GB in and 64 GB out -- more than the 180 GB of a B200), so a plan whose workspace exceeds the budget is
    //      processed in stripes of as many tile rows as fit, one after the other through the same workspace;
    //  (2) the tail: the last tile row is mostly padding, part B gets a shorter tile.
    static const double budget = getenv("NDCONV_WS_BUDGET_MB") ? atof(getenv("NDCONV_WS_BUDGET_MB")) * 1048576.0 : 24.0 * 1073741824.0;
    const int al = g.ndim - 1;
    double ws_tile_row = 8.0 * (pl.is_cx ? (double)pl.tl[al].F : (double)(pl.tl[al].F / 2 + 8)) * (double)pl.tl[al].ntiles;
    for (int a = 0; a < al; a++) ws_tile_row *= (double)pl.tl[a].F * (a > 0 ? (double)pl.tl[a].ntiles : 1.0);
    int64_t m = t.ntiles - 1;
    const bool over_budget = ws_tile_row * (double)t.ntiles > budget;
    if (over_budget) m = std::max<int64_t>(1, std::min<int64_t>(t.ntiles - 1, (int64_t)(budget / ws_tile_row)));
    const int64_t covered = m * t.V;                                        // padded positions whose outputs part A produces
    const int64_t o_split = (covered + g.s[0] - 1) / g.s[0];
    if (o_split <= 0 || o_split >= g.O[0]) return;
    const int64_t pB = o_split * g.s[0];                                    // first padded row part B reads
    const int64_t rowsA = covered + g.Kd[0] - 1 - g.pf[0];                  // input rows of part A (its back pad is 0)
    if (pB < g.pf[0] || rowsA < 1 || rowsA > g.n[0]) return;                // the cut must fall inside the array on both sides
    const int64_t rowsB = g.n[0] - (pB - g.pf[0]);
    if (rowsB < 1 || g.pb[0] >= rowsB || g.pf[0] >= rowsA) return;          // a border may not reach past the part that carries it
    if (!over_budget) {
        static const int menu_col[7] = {16, 32, 64, 128, 256, 512, 1024};
        AxisTiling tb; tb.F = 0;
        fast_pick_tile(g.P[0] - pB, g.Kd[0], menu_col, 7, &tb);
        if (tb.F == 0) return;
        const int64_t full = (int64_t)t.ntiles * t.F, split = (int64_t)(t.ntiles - 1) * t.F + (int64_t)tb.ntiles * tb.F;
        double samples = 1.0;
        for (int a = 0; a < g.ndim; a++) samples *= (double)g.P[a];
        if (samples < 4.0e6 || (full - split) * 64 < full) return;          // three more launches must buy at least 1.5 % of the rows
        // measured: a tail of ~10 % of the rows (one rank of c5 on 8 GPUs) gains 7.5 % from running beside the main part, a tail
        // of ~1 % (c5 on one GPU) loses 2 % (its CTAs delay a few CTAs of the main part's persistent grids)
        e->split_concurrent = (int64_t)tb.ntiles * tb.F * 25 >= split;
    } else e->split_concurrent = false;                                     // stripes share one workspace: one after the other
    e->split_out = o_split;
    e->split_rows_a = rowsA;
#else
    (void)pr; (void)g; (void)pl;
#endif
}

static int get_plan_entry(ndconv_processor *p, const ndconv_problem *pr, int path, PlanEntry **out)
{
    if (!pr) { set_error("null problem"); return NDCONV_ERR_BAD_ARG; }
    const int N = (pr->ndim >= 1 && pr->ndim <= NDC_MAX_DIM) ? pr->ndim : 0;
    std::vector<unsigned char> key;
    key.reserve(512);
    auto push = [&](const void *d, size_t n) { key.insert(key.end(), (const unsigned char *)d, (const unsigned char *)d + n); };
    const size_t es = dtype_size(pr->dtype);
    std::vector<unsigned char> kpacked;
    bool keyable = N > 0 && es > 0 && pr->kernel != nullptr;
    if (keyable) {
        int32_t hdr[5] = {path, pr->dtype, pr->ndim, pr->memory, pr->reverse};
        push(hdr, sizeof(hdr));
        for (int a = 0; a < N; a++) {
            int64_t v[9] = {pr->data_shape[a], pr->data_strides[a], pr->kernel_shape[a], pr->kernel_strides[a], pr->dilation[a],
                            pr->pad[a][0], pr->pad[a][1], pr->stride[a], 0};
            push(v, sizeof(v));
            push(&pr->border[a][0].type, 4); push(pr->border[a][0].value, 16);
            push(&pr->border[a][1].type, 4); push(pr->border[a][1].value, 16);
            if (pr->kernel_shape[a] < 0) keyable = false;
        }
    }
    if (keyable && path == NDCONV_PATH_DIRECT) {
        // the tap table depends on the kernel values
        int64_t kt = 1;
        for (int a = 0; a < N; a++) kt *= pr->kernel_shape[a];
        if (kt > 0 && kt < (1 << 22)) {
            kpacked.resize((size_t)kt * es);
            pack_strided(pr->kernel, N, pr->kernel_shape, pr->kernel_strides, (int)es, kpacked.data());
            push(kpacked.data(), kpacked.size());
        } else keyable = false;
    }
    p->tick++;
    if (keyable) {
        for (auto &e : p->plans) if (e->key == key) { e->last_use = p->tick; p->plan_hits++; *out = e.get(); return NDCONV_OK; }
    }
    p->plan_misses++;
    std::unique_ptr<PlanEntry, PlanEntryDeleter> ent(new PlanEntry());
    ent->key = key;
    ent->last_use = p->tick;
    std::vector<int32_t> maps[NDC_MAX_DIM];
    int st = check_problem(pr, path, &ent->g, maps); if (st) return st;
    Geom &g = ent->g;
    ent->host_contiguous = g.data_contiguous;
    if (pr->memory == NDCONV_MEM_HOST) { int64_t sacc = 1; for (int i = g.ndim - 1; i >= 0; i--) { g.xstr[i] = sacc; sacc *= g.n[i]; } }
    st = set_device(p); if (st) return st;
    if (path == NDCONV_PATH_DIRECT) {
        if (kpacked.empty()) { kpacked.resize((size_t)g.kernel_total * g.es); pack_strided(pr->kernel, g.ndim, g.k, g.kstr, g.es, kpacked.data()); }
        { int64_t sacc = 1; for (int i = g.ndim - 1; i >= 0; i--) { g.kstr[i] = sacc; sacc *= g.k[i]; } }
        Taps taps; build_taps(g, kpacked.data(), taps);
        ent->ntap = taps.ntap;
        st = upload_meta(p, ent->meta, g, maps, &taps, &ent->ml); if (st) return st;
    } else {
        st = make_plan(g, &ent->pl); if (st) return st;
        plan_axis0_split(pr, g, ent->pl, ent.get());
        st = upload_meta(p, ent->meta, g, maps, nullptr, &ent->ml); if (st) return st;
    }
    st = be_sync(p->stream); if (st) return st;      // the staging vector of upload_meta dies here
    if (p->plans.size() >= 32) {
        size_t victim = 0;
        for (size_t i = 1; i < p->plans.size(); i++) if (p->plans[i]->last_use < p->plans[victim]->last_use) victim = i;
        be_sync(p->stream);
        p->plans.erase(p->plans.begin() + victim);
    }
    *out = ent.get();
    p->plans.push_back(std::move(ent));
    return NDCONV_OK;
}

// stage the data array of a host problem on the device (packing strided views), or use a device array in place
static int stage_input2(ndconv_processor *p, const ndconv_problem *pr, const PlanEntry &e, const void **dev_x)
{
    if (pr->memory == NDCONV_MEM_DEVICE) { *dev_x = pr->data; return NDCONV_OK; }
    const Geom &g = e.g;
    size_t bytes = (size_t)g.data_total * g.es;
    int st = p->in_stage.reserve(bytes); if (st) return st;
    if (e.host_contiguous) return (*dev_x = p->in_stage.p, be_h2d(p->in_stage.p, pr->data, bytes, p->stream));
    std::vector<unsigned char> tmp(bytes);
    pack_strided(pr->data, g.ndim, g.n, pr->data_strides, g.es, tmp.data());
    st = be_h2d(p->in_stage.p, tmp.data(), bytes, p->stream); if (st) return st;
    st = be_sync(p->stream); if (st) return st;
    *dev_x = p->in_stage.p;
    return NDCONV_OK;
}

#ifdef NDCONV_CUDA
// ---- sm_100a tile-plus-halo direct convolution (kernels_direct_tile.cuh) -------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_tmap_encoder()
{
    static PFN_tmapEncodeTiled fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
        return (PFN_tmapEncodeTiled)f;
    }();
    return fn;
}

static bool value_is_zero(const ndconv_border &b, int es)
{
    if (b.type == NDCONV_BORDER_ZEROS) return true;
    if (b.type != NDCONV_BORDER_CONST) return false;
    for (int i = 0; i < es; i++) if (b.value[i]) return false;
    return true;
}

template <class T>
static int launch_direct_tile(ndconv_processor *p, const CUtensorMap &tm, const tile::TileParams &tp, int64_t grid, size_t smem, double alg_bytes)
{
    NDC_ONCE_PER_DEVICE(CU_CHECK(cudaFuncSetAttribute(tile::direct_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    const stream_t stm = p->stream;
    return launch_raw(p->lc(), tp.use_tma ? "direct_conv_tile_tma" : "direct_conv_tile", alg_bytes,
                      [&] {
                          static const bool no_pdl = getenv("NDCONV_DISABLE_PDL") != nullptr;
                          cudaLaunchConfig_t cfg = {};
                          cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(tile::kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stm;
                          cudaLaunchAttribute at[1];
                          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
                          cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
                          cudaLaunchKernelEx(&cfg, tile::direct_tile_kernel<T>, tm, tp);
                      });
}

// returns NDCONV_OK with *used = false when the problem is outside the tile kernel's envelope (caller falls back)
static int try_direct_tile(ndconv_processor *p, const ndconv_problem *pr, const PlanEntry &e, const void *dev_x, void *dev_out, bool *used)
{
    *used = false;
    static const bool disabled = getenv("NDCONV_DISABLE_TILE") != nullptr;
    const Geom &g = e.g;
    if (disabled || g.ndim > 3 || e.ntap == 0) return NDCONV_OK;
    const int sh = 3 - g.ndim, es = g.es;
    tile::TileParams tp; memset(&tp, 0, sizeof(tp));
    int st;
    if (!p->zero_map.p) {
        st = p->zero_map.reserve(16); if (st) return st;
        CU_CHECK(cudaMemsetAsync(p->zero_map.p, 0, 16, p->stream));
    }
    const unsigned char *mb = (const unsigned char *)e.meta.p;
    for (int a3 = 0; a3 < 3; a3++) {
        const int a = a3 - sh;
        if (a < 0) {
            tp.n[a3] = 1; tp.xstr[a3] = 0; tp.P[a3] = 1; tp.pf[a3] = 0; tp.Kd[a3] = 1; tp.s[a3] = 1; tp.O[a3] = 1;
            tp.map[a3] = (const int32_t *)p->zero_map.p; tp.front_zero[a3] = tp.back_zero[a3] = 1;
        } else {
            tp.n[a3] = g.n[a]; tp.xstr[a3] = g.xstr[a]; tp.P[a3] = g.P[a]; tp.pf[a3] = g.pf[a]; tp.Kd[a3] = g.Kd[a]; tp.s[a3] = g.s[a]; tp.O[a3] = g.O[a];
            tp.map[a3] = (const int32_t *)(mb + e.ml.map_off[a]);
            memset(tp.cfront[a3], 0, 16); memset(tp.cback[a3], 0, 16);
            if (pr->border[a][0].type == NDCONV_BORDER_CONST) memcpy(tp.cfront[a3], pr->border[a][0].value, 16);
            if (pr->border[a][1].type == NDCONV_BORDER_CONST) memcpy(tp.cback[a3], pr->border[a][1].value, 16);
            tp.front_zero[a3] = value_is_zero(pr->border[a][0], es); tp.back_zero[a3] = value_is_zero(pr->border[a][1], es);
        }
    }
    if (tp.Kd[0] > (1 << 20) || tp.Kd[1] > (1 << 20) || tp.Kd[2] > (1 << 20)) return NDCONV_OK;
    tp.ostr[2] = 1; tp.ostr[1] = tp.O[2]; tp.ostr[0] = tp.O[1] * tp.O[2];
    // tile shape
    auto np2 = [](int64_t v) { int r = 1; while (r < v && r < 256) r <<= 1; return r; };
    int TO2 = np2(tp.O[2]);
    while (TO2 > 32 && (tile::kThreads / TO2) < std::min<int64_t>(tp.O[1], 8)) TO2 >>= 1;
    int TO1 = (int)std::min<int64_t>(tile::kThreads / TO2, np2(tp.O[1]));
    int TO0 = (int)std::min<int64_t>(tile::kMaxTO0, tp.O[0]);
    const int round = 16 / std::min(es, 16);
    auto shape = [&](int T0, int T1, int T2) {
        tp.TO[0] = T0; tp.TO[1] = T1; tp.TO[2] = T2;
        for (int a = 0; a < 3; a++) tp.IT[a] = (int)((tp.TO[a] - 1) * tp.s[a] + tp.Kd[a]);
        tp.IT2p = (tp.IT[2] + (round - 1) + round - 1) / round * round;   // + (round-1): room for the per-tile alignment shift
        return (int64_t)tp.IT[0] * tp.IT[1] * tp.IT2p;
    };
    int64_t elems = shape(TO0, TO1, TO2);
    const int64_t budget = 150 * 1024;
    while (elems * es > budget && TO0 > 1) { TO0--; elems = shape(TO0, TO1, TO2); }
    while (elems * es > budget && TO1 > 1) { TO1 >>= 1; elems = shape(TO0, TO1, TO2); }
    while (elems * es > budget && TO2 > 1) { TO2 >>= 1; elems = shape(TO0, TO1, TO2); }
    if (elems * es > budget) return NDCONV_OK;
    const size_t tap_bytes = align_up((size_t)e.ntap * es, 16) + (size_t)e.ntap * 4 + (size_t)(tp.IT[0] + tp.IT[1] + tp.IT2p) * 4;   // taps + the window's border maps
    const size_t tile_bytes = align_up((size_t)elems * es, 128);
    if (tile_bytes + tap_bytes + 256 > 190 * 1024) return NDCONV_OK;
    tp.tile_elems = (int)elems;
    int64_t grid = 1;
    for (int a = 0; a < 3; a++) { tp.ntile[a] = (int)((tp.O[a] + tp.TO[a] - 1) / tp.TO[a]); grid *= tp.ntile[a]; }
    if (grid > 0x7fffffff) return NDCONV_OK;
    tp.ntap = e.ntap; tp.axis_shift = sh;
    tp.tap_off = (const int32_t *)(mb + e.ml.tap_off_off);
    tp.tap_w = mb + e.ml.tap_w_off;
    tp.x = dev_x; tp.out = dev_out;
    // TMA eligibility: 4/8-byte elements, standard layout, 16-byte aligned rows
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    bool tma = (es == 4 || es == 8) && tp.xstr[2] == 1 && (tp.n[1] == 1 || tp.xstr[1] == tp.n[2]) && (tp.n[0] == 1 || tp.xstr[0] == tp.n[1] * tp.n[2]) &&
               ((tp.n[2] * es) % 16 == 0) && ((uintptr_t)dev_x % 16 == 0) && tp.IT[0] <= 256 && tp.IT[1] <= 256 && tp.IT2p <= 256 && get_tmap_encoder();
    if (tma) {
        cuuint64_t gdim[3] = {(cuuint64_t)tp.n[2], (cuuint64_t)tp.n[1], (cuuint64_t)tp.n[0]};
        cuuint64_t gstr[2] = {(cuuint64_t)tp.n[2] * es, (cuuint64_t)tp.n[1] * tp.n[2] * es};
        cuuint32_t box[3] = {(cuuint32_t)tp.IT2p, (cuuint32_t)tp.IT[1], (cuuint32_t)tp.IT[0]};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = get_tmap_encoder()(&tm, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(dev_x), gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tma = false;
        if (getenv("NDCONV_DEBUG_TMA")) fprintf(stderr, "[ndconv] tmap encode r=%d es=%d gdim=(%llu,%llu,%llu) gstr=(%llu,%llu) box=(%u,%u,%u) x=%p\n", (int)r, es,
            (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2], (unsigned long long)gstr[0], (unsigned long long)gstr[1], box[0], box[1], box[2], dev_x);
    }
    static const bool no_tma = getenv("NDCONV_DISABLE_TMA") != nullptr;
    if (no_tma) tma = false;
    tp.use_tma = tma ? 1 : 0;
    const size_t smem = tile_bytes + tap_bytes + 128;
    const double alg_bytes = (double)es * ((double)g.data_total + (double)g.out_total + (double)e.ntap);
    switch (g.dtype) {
    case NDCONV_I8: case NDCONV_U8: st = launch_direct_tile<uint8_t>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_I16: case NDCONV_U16: st = launch_direct_tile<uint16_t>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_I32: case NDCONV_U32: st = launch_direct_tile<uint32_t>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_I64: case NDCONV_U64: st = launch_direct_tile<uint64_t>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_F32: st = launch_direct_tile<float>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_F64: st = launch_direct_tile<double>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_C32: st = launch_direct_tile<cx<float>>(p, tm, tp, grid, smem, alg_bytes); break;
    case NDCONV_C64: st = launch_direct_tile<cx<double>>(p, tm, tp, grid, smem, alg_bytes); break;
    default: return NDCONV_OK;
    }
    if (st) return st;
    *used = true;
    return NDCONV_OK;
}
#endif

// ======================================================================================================
// direct convolution
// ======================================================================================================
static int conv_direct_impl(ndconv_processor *p, const ndconv_problem *pr, void *out)
{
    PlanEntry *e = nullptr;
    int st = get_plan_entry(p, pr, NDCONV_PATH_DIRECT, &e); if (st) return st;
    if (!out) { set_error("null output pointer"); return NDCONV_ERR_BAD_ARG; }
    st = set_device(p); if (st) return st;
    const Geom &g = e->g;
    const void *dev_x = nullptr;
    st = stage_input2(p, pr, *e, &dev_x); if (st) return st;

    DirectParams dp; memset(&dp, 0, sizeof(dp));
    dp.ndim = g.ndim; dp.ntap = e->ntap; dp.x = dev_x; dp.total = g.out_total;
    const unsigned char *mb = (const unsigned char *)e->meta.p;
    for (int a = 0; a < g.ndim; a++) {
        dp.xstr[a] = g.xstr[a]; dp.n[a] = g.n[a]; dp.P[a] = g.P[a]; dp.pf[a] = g.pf[a]; dp.Kd[a] = g.Kd[a]; dp.s[a] = g.s[a]; dp.O[a] = g.O[a];
        dp.map[a] = (const int32_t *)(mb + e->ml.map_off[a]);
    }
    dp.tap_off = (const int32_t *)(mb + e->ml.tap_off_off);
    dp.tap_lin = (const int64_t *)(mb + e->ml.tap_lin_off);
    dp.tap_w = mb + e->ml.tap_w_off;
    fill_consts(pr, g.ndim, dp.cfront, dp.cback);
    size_t obytes = (size_t)g.out_total * g.es;
    if (pr->memory == NDCONV_MEM_HOST) { st = p->out_stage.reserve(obytes); if (st) return st; dp.out = p->out_stage.p; }
    else dp.out = out;

    const double alg_bytes = (double)g.es * ((double)g.data_total + (double)g.out_total + (double)e->ntap);
    bool tiled = false;
#ifdef NDCONV_CUDA
    st = try_direct_tile(p, pr, *e, dev_x, dp.out, &tiled); if (st) return st;
#endif
    if (tiled) st = NDCONV_OK;
    else switch (g.dtype) {
    case NDCONV_I8: case NDCONV_U8: st = run_direct_t<uint8_t>(p, dp, alg_bytes); break;
    case NDCONV_I16: case NDCONV_U16: st = run_direct_t<uint16_t>(p, dp, alg_bytes); break;
    case NDCONV_I32: case NDCONV_U32: st = run_direct_t<uint32_t>(p, dp, alg_bytes); break;
    case NDCONV_I64: case NDCONV_U64: st = run_direct_t<uint64_t>(p, dp, alg_bytes); break;
    case NDCONV_F32: st = run_direct_t<float>(p, dp, alg_bytes); break;
    case NDCONV_F64: st = run_direct_t<double>(p, dp, alg_bytes); break;
    case NDCONV_C32: st = run_direct_t<cx<float>>(p, dp, alg_bytes); break;
    case NDCONV_C64: st = run_direct_t<cx<double>>(p, dp, alg_bytes); break;
    default: set_error("dtype"); st = NDCONV_ERR_BAD_ARG;
    }
    if (st) return st;
    if (pr->memory == NDCONV_MEM_HOST) {
        st = be_d2h(out, p->out_stage.p, obytes, p->stream); if (st) return st;
        st = be_sync(p->stream); if (st) return st;
    }
    return NDCONV_OK;
}

static FastDiv make_fastdiv(int d)
{
    FastDiv f; f.d = d < 1 ? 1 : d;
    int l = 0;
    while ((1ll << l) < f.d) l++;
    f.shift = 32 + l;
    f.mul = (uint64_t)((((unsigned __int128)1 << f.shift) + (unsigned)f.d - 1) / (unsigned)f.d);     // ceil(2^shift / d): exact for n < 2^32
    return f;
}

template <class R> static int fill_plan_dev(ndconv_processor *p, const FftLen &fl, FftPlanDev<R> *d)
{
    d->L = fl.L; d->npass = fl.npass;
    for (int i = 0; i < NDC_MAX_PASS; i++) d->radix[i] = i < fl.npass ? fl.radix[i] : 1;
    int ns = 1;
    for (int i = 0; i < NDC_MAX_PASS; i++) {
        const int r = d->radix[i];
        d->by_m[i] = make_fastdiv(std::max(1, fl.L / r));
        d->by_ns[i] = make_fastdiv(ns);
        if (i < fl.npass) ns *= r;
    }
    return get_tw_c<R>(p, fl.L, &d->tw);
}

static int pick_rows_per_block(int L, size_t csz)
{
    int B = std::max(1, 1024 / std::max(L, 1));
    B = std::min(B, 16);
    while (B > 1 && 2 * (size_t)B * (L + 1) * csz + (size_t)B * 64 > kRowSmemBudget) B--;
    return B;
}
static int pick_block_threads(int64_t butterflies)
{
    int t = (int)std::min<int64_t>(512, std::max<int64_t>(64, (butterflies + 31) / 32 * 32));
    return t;
}

template <class R>
static int run_row(ndconv_processor *p, int kind /*0 fwd 1 inv 2 1d*/, RowParams<R> &rp, int64_t out_rows, const char *name, double alg_bytes)
{
    const size_t csz = sizeof(cx<R>);
    const int L = rp.plan.L;
    if (kind == 2) rp.B = 1;
    else rp.B = pick_rows_per_block(L, csz);
    size_t smem = align_up(2 * (size_t)rp.B * (L + 1) * csz + (size_t)rp.B * 64 + 64, 16);
    rp.tw_smem_off = 0;
    if (smem + (size_t)L * csz <= 160 * 1024 && L > 1) { rp.tw_smem_off = (int)smem; smem += (size_t)L * csz; }   // twiddle table copy
    if (smem > 227 * 1024) { set_error("internal: row kernel shared memory"); return NDCONV_ERR_INTERNAL; }
    int N = rp.ndim;
    int64_t ntiles_total = 1;
    for (int a = 0; a < N; a++) ntiles_total *= rp.ntiles[a];
    if (kind == 0) rp.nwork = ntiles_total * ((rp.rows_per_tile + rp.B - 1) / rp.B);
    else if (kind == 1) rp.nwork = ((out_rows + rp.B - 1) / rp.B) * rp.ntiles[N - 1];
    else rp.nwork = rp.ntiles[0];
    int block = pick_block_threads((int64_t)rp.B * std::max(L / 4, 1));
    int64_t grid = std::min<int64_t>(rp.nwork, (int64_t)p->num_sms * 4 * kMaxGridMult);
    if (kind == 0) return launch<RowFwdBody<R>, RowParams<R>>(p->lc(), name, alg_bytes, grid, block, smem, rp);
    if (kind == 1) return launch<RowInvBody<R>, RowParams<R>>(p->lc(), name, alg_bytes, grid, block, smem, rp);
    return launch<Row1DBody<R>, RowParams<R>>(p->lc(), name, alg_bytes, grid, block, smem, rp);
}

template <class R>
static int run_col(ndconv_processor *p, const FftPlan &pl, int axis, int mode, cx<R> *ws, const cx<R> *kspec, int64_t ntiles_total,
                   const char *name, double alg_bytes)
{
    ColParams<R> cp; memset(&cp, 0, sizeof(cp));
    cp.ws = ws; cp.kspec = kspec; cp.F = pl.tl[axis].F; cp.mode = mode;
    cp.outer = 1; cp.inner = pl.Hp;
    for (int b = 0; b < axis; b++) cp.outer *= pl.tl[b].F;
    for (int b = axis + 1; b < pl.N - 1; b++) cp.inner *= pl.tl[b].F;
    cp.tile_elems = pl.tile_elems; cp.ntiles_total = ntiles_total;
    int st = fill_plan_dev<R>(p, pl.fl[axis], &cp.plan); if (st) return st;
    int W = 16;
    while (W > 1 && 2 * (size_t)cp.F * W * sizeof(cx<R>) > kColSmemBudget) W >>= 1;
    cp.W = W;
    cp.nwork = ntiles_total * cp.outer * (cp.inner / W);
    size_t smem = align_up(2 * (size_t)cp.F * W * sizeof(cx<R>) + 64, 16);
    cp.tw_smem_off = 0;
    if (cp.F > 1) { cp.tw_smem_off = (int)smem; smem += (size_t)cp.F * sizeof(cx<R>); }
    int block = pick_block_threads((int64_t)cp.F * W / 4);
    int64_t grid = std::min<int64_t>(cp.nwork, (int64_t)p->num_sms * 4 * kMaxGridMult);
    return launch<ColBody<R>, ColParams<R>>(p->lc(), name, alg_bytes, grid, block, smem, cp);
}

// kernel spectrum (cached per processor): conv_fft::padding::kernel (src/conv_fft/padding.rs:78-111) + forward
template <class R>
static int get_kernel_spectrum(ndconv_processor *p, const ndconv_problem *pr, const Geom &g, const FftPlan &pl, const cx<R> **out, KSpecEntry **out_entry = nullptr)
{
    const int N = g.ndim;
    std::vector<unsigned char> kpacked((size_t)g.kernel_total * g.es);
    pack_strided(pr->kernel, N, g.k, g.kstr, g.es, kpacked.data());
    // cache key
    std::vector<unsigned char> key;
    auto push = [&](const void *d, size_t n) { key.insert(key.end(), (const unsigned char *)d, (const unsigned char *)d + n); };
    int hdr[3] = {g.dtype, N, g.reverse ? 1 : 0};
    push(hdr, sizeof(hdr));
    for (int a = 0; a < N; a++) { int64_t v[3] = {g.k[a], g.d[a], (int64_t)pl.tl[a].F}; push(v, sizeof(v)); }
    push(kpacked.data(), kpacked.size());
    p->tick++;
    for (auto &e : p->kspecs) if (e->key == key) { e->last_use = p->tick; *out = (const cx<R> *)e->buf.p; if (out_entry) *out_entry = e.get(); return NDCONV_OK; }

    // dense dilated (and, for no_reverse, flipped) kernel of extent Kd, pre-scaled by 1/prod(F) (the reference divides
    // after the inverse transform: real.rs:278-279, complex.rs:141-142)
    int64_t kdtot = 1, kdstr[NDC_MAX_DIM];
    for (int a = N - 1; a >= 0; a--) { kdstr[a] = kdtot; kdtot *= g.Kd[a]; }
    long double scale = 1.0L;
    for (int a = 0; a < N; a++) scale /= (long double)pl.tl[a].F;
    const int nc = pl.is_cx ? 2 : 1;
    std::vector<R> kb((size_t)kdtot * nc, (R)0);
    {
        int64_t idx[NDC_MAX_DIM] = {0};
        const R *ks = (const R *)kpacked.data();
        for (int64_t e = 0; e < g.kernel_total; e++) {
            int64_t o = 0;
            for (int a = 0; a < N; a++) o += (g.reverse ? idx[a] * g.d[a] : g.Kd[a] - 1 - idx[a] * g.d[a]) * kdstr[a];   // padding.rs:98-108
            for (int c = 0; c < nc; c++) kb[(size_t)o * nc + c] = (R)((long double)ks[(size_t)e * nc + c] * scale);
            for (int a = N - 1; a >= 0; a--) { if (++idx[a] < g.k[a]) break; idx[a] = 0; }
        }
    }
    int st = p->kb_stage.reserve(kb.size() * sizeof(R)); if (st) return st;
    st = be_h2d(p->kb_stage.p, kb.data(), kb.size() * sizeof(R), p->stream); if (st) return st;

    // identity maps of length Kd
    Geom kg = g;
    std::vector<int32_t> kmaps[NDC_MAX_DIM];
    for (int a = 0; a < N; a++) { kmaps[a].resize((size_t)g.Kd[a]); for (int64_t i = 0; i < g.Kd[a]; i++) kmaps[a][(size_t)i] = (int32_t)i; }
    MetaLayout ml;
    st = upload_meta(p, p->kmeta, kg, kmaps, nullptr, &ml); if (st) return st;

    std::unique_ptr<KSpecEntry> ent(new KSpecEntry());
    ent->key = key; ent->last_use = p->tick;
    st = ent->buf.reserve((size_t)pl.tile_elems * sizeof(cx<R>)); if (st) return st;

    RowParams<R> rp; memset(&rp, 0, sizeof(rp));
    rp.ndim = N; rp.is_cx = pl.is_cx ? 1 : 0;
    for (int a = 0; a < N; a++) {
        rp.n[a] = g.Kd[a]; rp.xstr[a] = kdstr[a]; rp.P[a] = g.Kd[a];
        rp.map[a] = (const int32_t *)((const unsigned char *)p->kmeta.p + ml.map_off[a]);
        rp.F[a] = pl.tl[a].F; rp.V[a] = pl.tl[a].F; rp.ntiles[a] = 1; rp.Kd[a] = 1; rp.s[a] = 1; rp.O[a] = 1;
    }
    rp.x = p->kb_stage.p; rp.ws = (cx<R> *)ent->buf.p; rp.H = pl.H; rp.Hp = pl.Hp;
    rp.rows_per_tile = pl.rows_per_tile; rp.tile_elems = pl.tile_elems;
    st = fill_plan_dev<R>(p, pl.fl[N - 1], &rp.plan); if (st) return st;
    if (!pl.is_cx) { st = get_tw_r<R>(p, pl.tl[N - 1].F, &rp.twr); if (st) return st; }
    const double kbytes = (double)pl.tile_elems * sizeof(cx<R>);
    st = run_row<R>(p, 0, rp, 0, "kspec_row_fwd", kbytes); if (st) return st;
    for (int a = N - 2; a >= 0; a--) { st = run_col<R>(p, pl, a, 0, (cx<R> *)ent->buf.p, nullptr, 1, "kspec_col_fwd", 2 * kbytes); if (st) return st; }
    // the staging buffers are reused by the next build: make sure this one has been consumed
    st = be_sync(p->stream); if (st) return st;

    if (p->kspecs.size() >= 8) {   // small LRU
        size_t victim = 0;
        for (size_t i = 1; i < p->kspecs.size(); i++) if (p->kspecs[i]->last_use < p->kspecs[victim]->last_use) victim = i;
        p->kspecs[victim]->buf.release();
        p->kspecs[victim]->pair.release();
        p->kspecs.erase(p->kspecs.begin() + victim);
    }
    *out = (const cx<R> *)ent->buf.p;
    if (out_entry) *out_entry = ent.get();
    p->kspecs.push_back(std::move(ent));
    return NDCONV_OK;
}

#ifdef NDCONV_CUDA
// sm_100a fast path: real f32, rank 2 / 3, power-of-two overlap-save tiles (kernels_fft_fast.cuh)
// every fast-path kernel is launched with programmatic stream serialization (see pdl_wait in kernels_fft_fast.cuh)
template <class P> static void launch_pdl(void (*kernel)(const P), int grid, int block, size_t smem, stream_t stm, const P &prm)
{
    static const bool no_pdl = getenv("NDCONV_DISABLE_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stm;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kernel, prm);
}
template <int T, int N> static void launch_row_n(bool inverse, const fast::RowParams &rp, int grid, stream_t stm)
{
    NDC_ONCE_PER_DEVICE(cudaFuncSetAttribute(fast::row_fwd<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::RowCfg<T>::smem);
                        cudaFuncSetAttribute(fast::row_inv<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::RowCfg<T>::smem));
    launch_pdl<fast::RowParams>(inverse ? fast::row_inv<T, N> : fast::row_fwd<T, N>, grid, 128, fast::RowCfg<T>::smem, stm, rp);
}
template <int T, int N> static void launch_row_cx_n(bool inverse, const fast::RowParams &rp, int grid, stream_t stm)
{
    NDC_ONCE_PER_DEVICE(cudaFuncSetAttribute(fast::row_fwd_c<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::RowCxCfg<T>::smem);
                        cudaFuncSetAttribute(fast::row_inv_c<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::RowCxCfg<T>::smem));
    launch_pdl<fast::RowParams>(inverse ? fast::row_inv_c<T, N> : fast::row_fwd_c<T, N>, grid, 128, fast::RowCxCfg<T>::smem, stm, rp);
}
template <int T> static void launch_row_cx(bool inverse, const fast::RowParams &rp, int grid, stream_t stm)
{
    if (rp.ndim == 2) launch_row_cx_n<T, 2>(inverse, rp, grid, stm);
    else launch_row_cx_n<T, 3>(inverse, rp, grid, stm);
}
template <int T> static void launch_row1d_cx(const fast::RowParams &rp, int grid, stream_t stm)
{
    NDC_ONCE_PER_DEVICE(cudaFuncSetAttribute(fast::row1d_c<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::Row1dCxCfg<T>::smem));
    launch_pdl<fast::RowParams>(fast::row1d_c<T>, grid, 128, fast::Row1dCxCfg<T>::smem, stm, rp);
}
template <int T> static void launch_row1d(const fast::RowParams &rp, int grid, stream_t stm)
{
    NDC_ONCE_PER_DEVICE(cudaFuncSetAttribute(fast::row1d<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast::Row1dCfg<T>::smem));
    launch_pdl<fast::RowParams>(fast::row1d<T>, grid, 128, fast::Row1dCfg<T>::smem, stm, rp);
}
template <int T> static void launch_row(bool inverse, const fast::RowParams &rp, int grid, stream_t stm)
{
    if (rp.ndim == 2) launch_row_n<T, 2>(inverse, rp, grid, stm);
    else launch_row_n<T, 3>(inverse, rp, grid, stm);
}
template <int E, int Tc> static void launch_col_t(const fast::ColParams &cp, int num_sms, stream_t stm)
{
    using C = fast::ColCfg<E, Tc>;
    NDC_ONCE_PER_DEVICE(cudaFuncSetAttribute(fast::col_pass<E, Tc>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem));
    const int per_sm = std::max(1, std::min(16, (int)((200 * 1024) / C::smem)));
    const int grid = (int)std::min<int64_t>(cp.nwork, (int64_t)num_sms * std::max(C::min_blocks, std::min(per_sm, 2048 / C::threads)));
    launch_pdl<fast::ColParams>(fast::col_pass<E, Tc>, grid, C::threads, C::smem, stm, cp);
}
static void launch_col(int F, const fast::ColParams &cp, int num_sms, stream_t stm)
{
    switch (F) {
    case 1024: launch_col_t<32, 32>(cp, num_sms, stm); break;
    case 512: launch_col_t<32, 16>(cp, num_sms, stm); break;
    case 256: launch_col_t<16, 16>(cp, num_sms, stm); break;
    case 128: launch_col_t<16, 8>(cp, num_sms, stm); break;
    case 64: launch_col_t<8, 8>(cp, num_sms, stm); break;
    case 32: launch_col_t<8, 4>(cp, num_sms, stm); break;
    default: launch_col_t<8, 2>(cp, num_sms, stm); break;
    }
}

static int conv_fft_fast(ndconv_processor *p, const ndconv_problem *pr, const Geom &g, const FftPlan &pl, const MetaLayout &ml, const DevBuf &metabuf,
                         const void *dev_x, void *dev_out, KSpecEntry *ent)
{
    using namespace ndc::fast;
    const int N = g.ndim, al = N - 1;
    const bool is_cx = pl.is_cx;                                 // Complex<f32>: C2C rows of L = F columns, no pairing, no pad columns
    const int L = is_cx ? pl.tl[al].F : pl.tl[al].F / 2, pitch = is_cx ? L : L + kPad, T = L / 32;
    int st;
    int64_t rows_per_tile = 1;
    for (int a = 0; a < al; a++) rows_per_tile *= pl.tl[a].F;
    const int64_t tile_elems = rows_per_tile * pitch;
    if (!ent->pair.p) {
        st = ent->pair.reserve((size_t)tile_elems * sizeof(cf)); if (st) return st;
        KfastParams kp; kp.kspec = (const cx<float> *)ent->buf.p; kp.kfast = (cx<float> *)ent->pair.p; kp.rows = rows_per_tile; kp.L = L; kp.Hp = pl.Hp; kp.is_cx = is_cx ? 1 : 0;
        st = launch<KfastBody, KfastParams>(p->lc(), "kspec_fast_layout", (double)tile_elems * 16, p->num_sms * 4, 256, 0, kp); if (st) return st;
    }
    if (N > 1) { st = p->ws.reserve((size_t)pl.ntiles_total * tile_elems * sizeof(cf)); if (st) return st; }
    const cx<float> *tw = nullptr, *twr = nullptr;
    st = get_tw_c<float>(p, L, &tw); if (st) return st;
    if (!is_cx) { st = get_tw_r<float>(p, 2 * L, &twr); if (st) return st; }

    fast::RowParams rp; memset(&rp, 0, sizeof(rp));
    rp.ndim = N;
    for (int a = 0; a < N; a++) {
        rp.n[a] = g.n[a]; rp.xstr[a] = g.xstr[a]; rp.P[a] = g.P[a]; rp.pf[a] = g.pf[a];
        rp.map[a] = (const int32_t *)((const unsigned char *)metabuf.p + ml.map_off[a]);
        rp.F[a] = pl.tl[a].F; rp.V[a] = pl.tl[a].V; rp.ntiles[a] = pl.tl[a].ntiles; rp.Kd[a] = (int)g.Kd[a]; rp.s[a] = g.s[a]; rp.O[a] = g.O[a];
        rp.cfront[a] = pr->border[a][0].type == NDCONV_BORDER_CONST ? *(const float *)pr->border[a][0].value : 0.f;
        rp.cback[a] = pr->border[a][1].type == NDCONV_BORDER_CONST ? *(const float *)pr->border[a][1].value : 0.f;
        if (is_cx) {
            rp.cfront_im[a] = pr->border[a][0].type == NDCONV_BORDER_CONST ? ((const float *)pr->border[a][0].value)[1] : 0.f;
            rp.cback_im[a] = pr->border[a][1].type == NDCONV_BORDER_CONST ? ((const float *)pr->border[a][1].value)[1] : 0.f;
        }
    }
    rp.x = (const float *)dev_x; rp.out = (float *)dev_out; rp.ws = (cf *)p->ws.p; rp.tw = tw; rp.twr = twr;
    rp.rows_per_tile = rows_per_tile; rp.tile_elems = tile_elems;
    if (N == 1) {
        // the whole pipeline in one launch and one pass over memory (fast::row1d): no workspace
        rp.kfast = (const cf *)ent->pair.p; rp.nwork = pl.tl[0].ntiles;
        const int64_t items = (rp.nwork + (32 / T) - 1) / (32 / T);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((items + 3) / 4, (int64_t)p->num_sms * 3 * 8));
        const stream_t stm1 = p->stream;
        return launch_raw(p->lc(), "row1d_fwd_mul_inv", (double)g.es * ((double)g.data_total + (double)g.out_total), [&] {
            if (is_cx) {
                switch (T) {
                case 32: launch_row1d_cx<32>(rp, grid, stm1); break;
                case 16: launch_row1d_cx<16>(rp, grid, stm1); break;
                case 8: launch_row1d_cx<8>(rp, grid, stm1); break;
                default: launch_row1d_cx<4>(rp, grid, stm1); break;
                }
                return;
            }
            switch (T) {
            case 32: launch_row1d<32>(rp, grid, stm1); break;
            case 16: launch_row1d<16>(rp, grid, stm1); break;
            case 8: launch_row1d<8>(rp, grid, stm1); break;
            default: launch_row1d<4>(rp, grid, stm1); break;
            }
        });
    }

    const double csz = 8.0;
    double S = is_cx ? (double)g.P[al] : (double)(g.P[al] / 2 + 1), So = S;           // un-inflated (half) spectrum of the padded array
    for (int a = 0; a < al; a++) { S *= (double)g.P[a]; So *= (double)g.O[a]; }
    const double in_bytes = (double)g.es * (double)g.data_total, out_bytes = (double)g.es * (double)g.out_total;
    const stream_t stm = p->stream;
    auto row_launch = [&](bool inverse) {
        const int64_t items = (rp.nwork + (32 / T) - 1) / (32 / T);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((items + 3) / 4, (int64_t)p->num_sms * 4 * 8));
        if (is_cx) {
            switch (T) {
            case 32: launch_row_cx<32>(inverse, rp, grid, stm); break;
            case 16: launch_row_cx<16>(inverse, rp, grid, stm); break;
            case 8: launch_row_cx<8>(inverse, rp, grid, stm); break;
            default: launch_row_cx<4>(inverse, rp, grid, stm); break;
            }
            return;
        }
        switch (T) {
        case 32: launch_row<32>(inverse, rp, grid, stm); break;
        case 16: launch_row<16>(inverse, rp, grid, stm); break;
        case 8: launch_row<8>(inverse, rp, grid, stm); break;
        default: launch_row<4>(inverse, rp, grid, stm); break;
        }
    };
    auto col_launch = [&](int axis, int mode, const char *name, double bytes) -> int {
        fast::ColParams cp; memset(&cp, 0, sizeof(cp));
        cp.ws = (cf *)p->ws.p; cp.kspec = (const cf *)ent->pair.p; cp.mode = mode;
        cp.outer = 1; cp.inner = pitch;
        for (int b = 0; b < axis; b++) cp.outer *= pl.tl[b].F;
        for (int b = axis + 1; b < al; b++) cp.inner *= pl.tl[b].F;
        cp.tile_elems = tile_elems; cp.ntiles_total = pl.ntiles_total;
        cp.nwork = pl.ntiles_total * cp.outer * (cp.inner / 8);
        const cx<float> *twc = nullptr;
        int s2 = get_tw_c<float>(p, pl.tl[axis].F, &twc); if (s2) return s2;
        cp.tw = twc;
        static const bool no_skip = getenv("NDCONV_COL_NO_SKIP") != nullptr;
        cp.skip = (mode != 0 && !no_skip) ? (int)g.Kd[axis] - 1 : 0;          // the crop discards tile rows [0, Kd - 1): row_inv never reads them
        return launch_raw(p->lc(), name, bytes, [&] { launch_col(pl.tl[axis].F, cp, p->num_sms, stm); });
    };
    // launches of the tail part of an axis-0 split are profiled under their own names: per-launch figures of the main kernels stay
    // comparable, and when the tail runs beside the main part its event-to-event durations include the time it waits for SMs
    const bool tail = p->in_tail;
    rp.nwork = pl.ntiles_total * rows_per_tile;
    st = launch_raw(p->lc(), tail ? "tail:row_fwd_pad_r2c" : "row_fwd_pad_r2c", in_bytes + S * csz, [&] { row_launch(false); }); if (st) return st;
    for (int a = al - 1; a >= 1; a--) { st = col_launch(a, 0, tail ? "tail:col_fwd" : "col_fwd", 2 * S * csz); if (st) return st; }
    st = col_launch(0, 2, tail ? "tail:col_fwd_mul_inv" : "col_fwd_mul_inv", 2 * S * csz + (double)tile_elems * csz); if (st) return st;
    for (int a = 1; a <= al - 1; a++) { st = col_launch(a, 1, tail ? "tail:col_inv" : "col_inv", 2 * S * csz); if (st) return st; }
    rp.nwork = pl.tl[al].ntiles;
    for (int a = 0; a < al; a++) rp.nwork *= g.O[a];
    st = launch_raw(p->lc(), tail ? "tail:row_inv_c2r_crop" : "row_inv_c2r_crop", So * csz + out_bytes, [&] { row_launch(true); }); if (st) return st;
    return NDCONV_OK;
}
#endif

static int conv_fft_impl(ndconv_processor *p, const ndconv_problem *pr, void *out);

template <class R>
static int conv_fft_t(ndconv_processor *p, const ndconv_problem *pr, PlanEntry *pe, void *out)
{
    const Geom &g = pe->g;
    const int N = g.ndim;
    const FftPlan &pl = pe->pl;
    int st;
    const void *dev_x = nullptr;
    st = stage_input2(p, pr, *pe, &dev_x); if (st) return st;
    if (pe->split_out > 0) {
        // two device-resident sub-convolutions along axis 0 (plan_axis0_split); each has its own cached plan
        const Geom gc = pe->g;                      // copies: the nested calls may evict this entry
        const int64_t o_split = pe->split_out, pB = o_split * gc.s[0];
        const int64_t rowsA = pe->split_rows_a;
        const size_t obytes = (size_t)gc.out_total * gc.es;
        void *dev_out = out;
        if (pr->memory == NDCONV_MEM_HOST) { st = p->out_stage.reserve(obytes); if (st) return st; dev_out = p->out_stage.p; }
        int64_t out_row = 1;
        for (int a = 1; a < N; a++) out_row *= gc.O[a];
        ndconv_problem base = *pr;
        base.memory = NDCONV_MEM_DEVICE; base.data = dev_x;
        for (int a = 0; a < N; a++) base.data_strides[a] = gc.xstr[a];
        ndconv_problem subA = base, subB = base;
        subA.data_shape[0] = rowsA; subA.pad[0][1] = 0; subA.border[0][1].type = NDCONV_BORDER_ZEROS;
        subB.data = (const char *)dev_x + (pB - gc.pf[0]) * gc.xstr[0] * (int64_t)gc.es;
        subB.data_shape[0] = gc.n[0] - (pB - gc.pf[0]); subB.pad[0][0] = 0; subB.border[0][0].type = NDCONV_BORDER_ZEROS;
        void *outB = (char *)dev_out + (size_t)o_split * (size_t)out_row * gc.es;
        bool concurrent = false;
#ifdef NDCONV_CUDA
        // The tail's launches are small (ramp and tail dominate them: measured ~65 % of the big launches' rate), so they run on
        // a second stream with their own workspace and fill the idle SMs at the ends of the main part's kernels.  Every cache
        // fill (plans, twiddles, kernel spectra) is synchronous on the host, so the two streams only share read-only state.
        static const bool sequential = getenv("NDCONV_SPLIT_SEQUENTIAL") != nullptr;
        concurrent = !sequential && !p->in_split && pe->split_concurrent;
        if (concurrent) {
            if (!p->aux_stream) {
                CU_CHECK(cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking));
                CU_CHECK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
                CU_CHECK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
            }
            CU_CHECK(cudaEventRecord(p->ev_fork, p->stream));               // the input (and the staging copy above) is ready on the main stream
            CU_CHECK(cudaStreamWaitEvent(p->aux_stream, p->ev_fork, 0));
            p->in_split = true;
            st = conv_fft_impl(p, &subA, dev_out);
            if (!st) {
                std::swap(p->stream, p->aux_stream); std::swap(p->ws, p->ws_aux);
                p->in_tail = true;
                st = conv_fft_impl(p, &subB, outB);
                p->in_tail = false;
                std::swap(p->stream, p->aux_stream); std::swap(p->ws, p->ws_aux);
            }
            p->in_split = false;
            if (st) return st;
            CU_CHECK(cudaEventRecord(p->ev_join, p->aux_stream));
            CU_CHECK(cudaStreamWaitEvent(p->stream, p->ev_join, 0));
        }
#endif
        if (!concurrent) {
            st = conv_fft_impl(p, &subA, dev_out); if (st) return st;
            p->in_tail = true;
            st = conv_fft_impl(p, &subB, outB);
            p->in_tail = false;
            if (st) return st;
        }
        if (pr->memory == NDCONV_MEM_HOST) {
            st = be_d2h(out, dev_out, obytes, p->stream); if (st) return st;
            st = be_sync(p->stream); if (st) return st;
        }
        return NDCONV_OK;
    }
    const MetaLayout &ml = pe->ml;
    DevBuf &metabuf = pe->meta;
    const cx<R> *kspec = nullptr;
    KSpecEntry *kent = nullptr;
    st = get_kernel_spectrum<R>(p, pr, g, pl, &kspec, &kent); if (st) return st;

    size_t obytes = (size_t)g.out_total * g.es;
    void *dev_out = out;
    if (pr->memory == NDCONV_MEM_HOST) { st = p->out_stage.reserve(obytes); if (st) return st; dev_out = p->out_stage.p; }
#ifdef NDCONV_CUDA
    if (pl.fast) {
        if constexpr (sizeof(R) == 4) {
            st = conv_fft_fast(p, pr, g, pl, ml, metabuf, dev_x, dev_out, kent); if (st) return st;
            if (pr->memory == NDCONV_MEM_HOST) {
                st = be_d2h(out, dev_out, obytes, p->stream); if (st) return st;
                st = be_sync(p->stream); if (st) return st;
            }
            return NDCONV_OK;
        }
    }
#endif

    RowParams<R> rp; memset(&rp, 0, sizeof(rp));
    rp.ndim = N; rp.is_cx = pl.is_cx ? 1 : 0;
    for (int a = 0; a < N; a++) {
        rp.n[a] = g.n[a]; rp.xstr[a] = g.xstr[a]; rp.P[a] = g.P[a];
        rp.map[a] = (const int32_t *)((const unsigned char *)metabuf.p + ml.map_off[a]);
        rp.F[a] = pl.tl[a].F; rp.V[a] = pl.tl[a].V; rp.ntiles[a] = pl.tl[a].ntiles; rp.Kd[a] = (int)g.Kd[a];
        rp.s[a] = g.s[a]; rp.O[a] = g.O[a];
    }
    fill_consts(pr, N, rp.cfront, rp.cback);
    rp.x = dev_x; rp.out = dev_out; rp.kspec = kspec; rp.H = pl.H; rp.Hp = pl.Hp;
    rp.rows_per_tile = pl.rows_per_tile; rp.tile_elems = pl.tile_elems;
    st = fill_plan_dev<R>(p, pl.fl[N - 1], &rp.plan); if (st) return st;
    if (!pl.is_cx) { st = get_tw_r<R>(p, pl.tl[N - 1].F, &rp.twr); if (st) return st; }

    // algorithmic bytes per pass: the un-inflated spectrum of the padded array, S complex elements
    const double csz = (double)sizeof(cx<R>);
    double S = pl.is_cx ? (double)g.P[N - 1] : (double)(g.P[N - 1] / 2 + 1), So = S;
    for (int a = 0; a < N - 1; a++) { S *= (double)g.P[a]; So *= (double)g.O[a]; }
    const double in_bytes = (double)g.es * (double)g.data_total, out_bytes = (double)g.es * (double)g.out_total;
    if (N == 1) {
        st = run_row<R>(p, 2, rp, 1, "row1d_fwd_mul_inv", in_bytes + out_bytes); if (st) return st;
    } else {
        st = p->ws.reserve((size_t)pl.ntiles_total * pl.tile_elems * sizeof(cx<R>)); if (st) return st;
        rp.ws = (cx<R> *)p->ws.p;
        st = run_row<R>(p, 0, rp, 0, "row_fwd_pad_r2c", in_bytes + S * csz); if (st) return st;
        for (int a = N - 2; a >= 1; a--) { st = run_col<R>(p, pl, a, 0, rp.ws, nullptr, pl.ntiles_total, "col_fwd", 2 * S * csz); if (st) return st; }
        st = run_col<R>(p, pl, 0, 2, rp.ws, kspec, pl.ntiles_total, "col_fwd_mul_inv", 2 * S * csz + (double)pl.tile_elems * csz); if (st) return st;
        for (int a = 1; a <= N - 2; a++) { st = run_col<R>(p, pl, a, 1, rp.ws, nullptr, pl.ntiles_total, "col_inv", 2 * S * csz); if (st) return st; }
        int64_t out_rows = 1;
        for (int a = 0; a < N - 1; a++) out_rows *= g.O[a];
        st = run_row<R>(p, 1, rp, out_rows, "row_inv_c2r_crop", So * csz + out_bytes); if (st) return st;
    }
    if (pr->memory == NDCONV_MEM_HOST) {
        st = be_d2h(out, dev_out, obytes, p->stream); if (st) return st;
        st = be_sync(p->stream); if (st) return st;
    }
    return NDCONV_OK;
}

static int conv_fft_impl(ndconv_processor *p, const ndconv_problem *pr, void *out);

#ifdef NDCONV_CUDA
// Host-resident problems that are large against PCIe: cut the OUTPUT rows of axis 0 into overlap-save slabs and run
// H2D(slab s+1) | kernels(slab s) | D2H(slab s-1) on three streams.  Axis-0 padding of a slab is materialised while
// staging (each padded row is copied from the source row its border map names; constant / never-written rows are
// filled), which is exactly "pad axis 0 first" of the reference's sequential definition (src/padding/mod.rs:119-153),
// so the slab runs as a plain device problem with no axis-0 border.
static const size_t kPipelineMinBytes = 96u << 20;
static const size_t kPipelineSlabBytes = 64u << 20;      // measured on c5: 160 MB 102.0 ms, 64 MB 100.0 ms, 32 MB 102.4 ms per step (fill + drain vs per-slab overhead)

static bool pipeline_eligible(const ndconv_problem *pr, const Geom &g)
{
    static const bool disabled = getenv("NDCONV_DISABLE_PIPELINE") != nullptr;
    if (disabled || pr->memory != NDCONV_MEM_HOST || !g.data_contiguous) return false;
    const size_t bytes = ((size_t)g.data_total + (size_t)g.out_total) * g.es;
    return bytes >= kPipelineMinBytes && g.O[0] >= 4;
}

// ---- pageable host arrays ---------------------------------------------------------------------------------------------
// cudaMemcpyAsync from / to ordinary pageable memory is staged by the driver on the calling thread: 13-14 GB/s of host<->device
// traffic for one thread, ~20 GB/s for two or more (tools/pageable_probe.py, tools/pageable_sharded_probe.py), against 70-80 GB/s
// from page-locked memory.  A Vec-backed ndarray is pageable, so the pipelined host path stages such arrays itself: several host
// threads memcpy the slab rows between the caller's array and pinned bounce buffers, and the DMA runs from / to those at full rate.
struct CopyTask { char *dst; const char *src; size_t bytes; };
static void parallel_copy(const std::vector<CopyTask> &tasks, int nthreads)
{
    size_t total = 0;
    for (const auto &t : tasks) total += t.bytes;
    if (!total) return;
    nthreads = (int)std::max<size_t>(1, std::min<size_t>((size_t)nthreads, total / (4u << 20) + 1));
    auto work = [&](int w) {
        const size_t lo = total / nthreads * w, hi = (w == nthreads - 1) ? total : total / nthreads * (w + 1);
        size_t pos = 0;
        for (const auto &t : tasks) {
            const size_t a = std::max(lo, pos), b = std::min(hi, pos + t.bytes);
            if (a < b) memcpy(t.dst + (a - pos), t.src + (a - pos), b - a);
            pos += t.bytes;
            if (pos >= hi) break;
        }
    };
    if (nthreads == 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int w = 1; w < nthreads; w++) th.emplace_back(work, w);
    work(0);
    for (auto &x : th) x.join();
}
static bool host_ptr_is_pageable(const void *ptr)
{
    static const bool disabled = getenv("NDCONV_DISABLE_BOUNCE") != nullptr;
    if (disabled) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeUnregistered;
}
static int reserve_bounce(void **buf, size_t *cap, size_t bytes)
{
    if (bytes <= *cap && buf[0] && buf[1]) return NDCONV_OK;
    for (int b = 0; b < 2; b++) { if (buf[b]) cudaFreeHost(buf[b]); buf[b] = nullptr; }
    *cap = 0;
    for (int b = 0; b < 2; b++) {
        if (cudaHostAlloc(&buf[b], bytes, cudaHostAllocDefault) != cudaSuccess) {
            buf[b] = nullptr;
            if (buf[0]) { cudaFreeHost(buf[0]); buf[0] = nullptr; }
            set_error("cudaHostAlloc failed for a bounce buffer");
            return NDCONV_ERR_CUDA;
        }
    }
    *cap = bytes;
    return NDCONV_OK;
}

// [o_lo, o_hi): the output rows of axis 0 this call produces (the whole axis for a single-GPU call, one slab of it per GPU in
// ndconv_conv_fft_sharded); `out` is always the base of the full output array
static int conv_fft_host_pipelined(ndconv_processor *p, const ndconv_problem *pr, const Geom &g, const std::vector<int32_t> &map0, void *out,
                                   int64_t o_lo = 0, int64_t o_hi = -1)
{
    if (o_hi < 0) o_hi = g.O[0];
    if (o_hi <= o_lo) return NDCONV_OK;
    const int64_t O0 = o_hi - o_lo;
    const int N = g.ndim;
    int st;
    if (!p->h2d_stream) {
        CU_CHECK(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
        CU_CHECK(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            CU_CHECK(cudaEventCreateWithFlags(&p->ev_h2d[b], cudaEventDisableTiming));
            CU_CHECK(cudaEventCreateWithFlags(&p->ev_comp[b], cudaEventDisableTiming));
            CU_CHECK(cudaEventCreateWithFlags(&p->ev_d2h[b], cudaEventDisableTiming));
        }
    }
    int64_t in_row_elems = 1, out_row_elems = 1;
    for (int a = 1; a < N; a++) { in_row_elems *= g.n[a]; out_row_elems *= g.O[a]; }
    const size_t in_row_bytes = (size_t)in_row_elems * g.es, out_row_bytes = (size_t)out_row_elems * g.es;
    // slab height in output rows: a multiple of the axis-0 tile payload so no tile is cut
    FftPlan fullpl; st = make_plan(g, &fullpl); if (st) return st;
    int64_t V0 = std::max<int64_t>(1, fullpl.tl[0].V / g.s[0]);
    static const size_t slab_bytes = getenv("NDCONV_PIPE_SLAB_MB") ? (size_t)atoll(getenv("NDCONV_PIPE_SLAB_MB")) << 20 : kPipelineSlabBytes;   // experiments
    int64_t rows = (int64_t)(slab_bytes / std::max<size_t>(1, std::max(in_row_bytes * g.s[0], out_row_bytes)));
    if (fullpl.fast && rows < V0) {
        // a slab shorter than one tile row of the full plan: take the payload of the largest power-of-two tile that fits the budget
        // (the slab's own plan picks that tile); shorter slabs shorten the fill and drain of the three-stage pipeline
        for (int F = fullpl.tl[0].F / 2; F >= 64 && rows < V0; F /= 2) { const int64_t v = (F - g.Kd[0] + 1) / g.s[0]; if (v >= 1 && 2 * (F - g.Kd[0] + 1) >= F) V0 = v; else break; }
    }
    rows = std::max<int64_t>(V0, rows / V0 * V0);
    if (rows >= O0) rows = std::max<int64_t>(1, (O0 + 1) / 2);
    const int64_t nslab = (O0 + rows - 1) / rows;
    const int64_t max_in_rows = (rows - 1) * g.s[0] + g.Kd[0];
    for (int b = 0; b < 2; b++) {
        st = p->pipe_in[b].reserve((size_t)max_in_rows * in_row_bytes); if (st) return st;
        st = p->pipe_out[b].reserve((size_t)rows * out_row_bytes); if (st) return st;
    }
    // pageable caller arrays are staged through pinned bounce buffers by host memcpy threads (see parallel_copy)
    bool bounce_x = host_ptr_is_pageable(pr->data), bounce_y = host_ptr_is_pageable(out);
    // measured (16384^2, k = 63^2, 24-core host): driver-staged 164 ms, 8 threads 74.8 ms, 16 threads 56.8 ms, pinned arrays 26.9 ms
    static const int copy_threads = getenv("NDCONV_HOST_COPY_THREADS") ? std::max(1, atoi(getenv("NDCONV_HOST_COPY_THREADS")))
                                                                        : (int)std::min(16u, std::max(4u, std::thread::hardware_concurrency() / 2));
    // no pinned memory to be had (locked-memory limit): fall back to the driver's own staging of pageable copies
    if (bounce_x && reserve_bounce(p->bounce_in, &p->bounce_in_cap, (size_t)max_in_rows * in_row_bytes)) { cudaGetLastError(); bounce_x = false; }
    if (bounce_y && reserve_bounce(p->bounce_out, &p->bounce_out_cap, (size_t)rows * out_row_bytes)) { cudaGetLastError(); bounce_y = false; }
    struct OutPending { bool live = false; int64_t ob = 0, oe = 0; } pend[2];       // D2H into bounce_out[b] issued, copy-out to the caller's array still due
    auto drain_tasks = [&](int b, std::vector<CopyTask> &tasks) -> int {              // slab in bounce_out[b] -> caller's rows (after its D2H has finished)
        if (!pend[b].live) return NDCONV_OK;
        CU_CHECK(cudaEventSynchronize(p->ev_d2h[b]));
        tasks.push_back(CopyTask{(char *)out + (size_t)pend[b].ob * out_row_bytes, (const char *)p->bounce_out[b], (size_t)(pend[b].oe - pend[b].ob) * out_row_bytes});
        pend[b].live = false;
        return NDCONV_OK;
    };
    // one host row holding the front / back constants of axis 0 (for constant border rows)
    const bool has_const = g.bf[0] == NDCONV_BORDER_CONST || g.bb[0] == NDCONV_BORDER_CONST;
    std::vector<unsigned char> crow_f, crow_b;
    if (has_const) {
        crow_f.resize(in_row_bytes); crow_b.resize(in_row_bytes);
        for (int64_t e = 0; e < in_row_elems; e++) { memcpy(&crow_f[e * g.es], pr->border[0][0].value, g.es); memcpy(&crow_b[e * g.es], pr->border[0][1].value, g.es); }
        st = p->pipe_row.reserve(2 * in_row_bytes); if (st) return st;
        st = be_h2d(p->pipe_row.p, crow_f.data(), in_row_bytes, p->h2d_stream); if (st) return st;
        st = be_h2d((char *)p->pipe_row.p + in_row_bytes, crow_b.data(), in_row_bytes, p->h2d_stream); if (st) return st;
    }
    const stream_t comp = p->stream;
    const char *hx = (const char *)pr->data;
    char *hout = (char *)out;
    int64_t prev_pb = 0, prev_pe = 0;
    for (int64_t sidx = 0; sidx < nslab; sidx++) {
        const int b = (int)(sidx & 1);
        const int64_t ob = o_lo + sidx * rows, oe = std::min<int64_t>(o_hi, ob + rows);
        const int64_t pb = ob * g.s[0], pe = (oe - 1) * g.s[0] + g.Kd[0];       // padded rows read by this slab
        // ---- H2D: materialise padded rows [pb, pe) of axis 0 ----
        if (sidx >= 2) CU_CHECK(cudaStreamWaitEvent(p->h2d_stream, p->ev_comp[b], 0));   // kernels of slab s-2 have consumed pipe_in[b]
        char *din = (char *)p->pipe_in[b].p;
        // the first Kd0 - 1 (halo) rows of this slab were uploaded with the previous one: copy them on the device instead of
        // sending them over PCIe again (c5: 62 of 1024 rows per slab, 6 % of the H2D bytes)
        int64_t i_start = pb;
        if (sidx >= 1 && prev_pe > pb) {
            const int64_t nrow = std::min(prev_pe, pe) - pb;
            CU_CHECK(cudaMemcpyAsync(din, (const char *)p->pipe_in[b ^ 1].p + (size_t)(pb - prev_pb) * in_row_bytes, (size_t)nrow * in_row_bytes,
                                     cudaMemcpyDeviceToDevice, p->h2d_stream));
            i_start = pb + nrow;
        }
        prev_pb = pb; prev_pe = pe;
        if (bounce_x || bounce_y) {
            // host side of this slab: rows of x into bounce_in[b] (free once the H2D of slab s-2 has read it), together with the
            // copy-out of slab s-2 from bounce_out[b] (its D2H was issued a whole iteration ago), on `copy_threads` threads
            std::vector<CopyTask> tasks;
            if (bounce_x) {
                if (sidx >= 2) CU_CHECK(cudaEventSynchronize(p->ev_h2d[b]));
                for (int64_t i = i_start; i < pe;) {
                    const int32_t m = map0[(size_t)i];
                    if (m < 0) { i++; continue; }
                    int64_t run = 1;
                    while (i + run < pe && map0[(size_t)(i + run)] == m + (int32_t)run) run++;
                    tasks.push_back(CopyTask{(char *)p->bounce_in[b] + (size_t)(i - pb) * in_row_bytes, hx + (size_t)m * in_row_bytes, (size_t)run * in_row_bytes});
                    i += run;
                }
            }
            if (bounce_y) { st = drain_tasks(b, tasks); if (st) return st; }
            parallel_copy(tasks, copy_threads);
            p->bounced_slabs++;
        }
        const char *hsrc = bounce_x ? (const char *)p->bounce_in[b] : nullptr;       // bounce layout = slab layout: row i at (i - pb)
        for (int64_t i = i_start; i < pe;) {
            const int32_t m = map0[(size_t)i];
            char *drow = din + (size_t)(i - pb) * in_row_bytes;
            if (m >= 0) {
                int64_t run = 1;
                while (i + run < pe && map0[(size_t)(i + run)] == m + (int32_t)run) run++;
                st = be_h2d(drow, hsrc ? hsrc + (size_t)(i - pb) * in_row_bytes : hx + (size_t)m * in_row_bytes, (size_t)run * in_row_bytes, p->h2d_stream); if (st) return st;
                i += run;
            } else if (m == NDC_MAP_INIT || (m == NDC_MAP_CONST_FRONT && g.bf[0] != NDCONV_BORDER_CONST) || (m == NDC_MAP_CONST_BACK && g.bb[0] != NDCONV_BORDER_CONST)) {
                CU_CHECK(cudaMemsetAsync(drow, 0, in_row_bytes, p->h2d_stream));
                i++;
            } else {
                const char *srow = (const char *)p->pipe_row.p + (m == NDC_MAP_CONST_BACK ? in_row_bytes : 0);
                CU_CHECK(cudaMemcpyAsync(drow, srow, in_row_bytes, cudaMemcpyDeviceToDevice, p->h2d_stream));
                i++;
            }
        }
        CU_CHECK(cudaEventRecord(p->ev_h2d[b], p->h2d_stream));
        // ---- kernels ----
        CU_CHECK(cudaStreamWaitEvent(comp, p->ev_h2d[b], 0));
        if (sidx >= 2) CU_CHECK(cudaStreamWaitEvent(comp, p->ev_d2h[b], 0));             // pipe_out[b] has been drained
        ndconv_problem sub = *pr;
        sub.memory = NDCONV_MEM_DEVICE;
        sub.data = din;
        sub.data_shape[0] = pe - pb;
        { int64_t stn = 1; for (int a = N - 1; a >= 0; a--) { sub.data_strides[a] = stn; stn *= (a == 0 ? pe - pb : g.n[a]); } }
        sub.pad[0][0] = sub.pad[0][1] = 0;
        sub.border[0][0].type = sub.border[0][1].type = NDCONV_BORDER_ZEROS;
        st = conv_fft_impl(p, &sub, p->pipe_out[b].p); if (st) return st;
        CU_CHECK(cudaEventRecord(p->ev_comp[b], comp));
        // ---- D2H ----
        CU_CHECK(cudaStreamWaitEvent(p->d2h_stream, p->ev_comp[b], 0));
        st = be_d2h(bounce_y ? (char *)p->bounce_out[b] : hout + (size_t)ob * out_row_bytes, p->pipe_out[b].p, (size_t)(oe - ob) * out_row_bytes, p->d2h_stream); if (st) return st;
        CU_CHECK(cudaEventRecord(p->ev_d2h[b], p->d2h_stream));
        if (bounce_y) { pend[b].live = true; pend[b].ob = ob; pend[b].oe = oe; }
        p->pipelined_slabs++;
    }
    st = be_sync(p->h2d_stream); if (st) return st;
    st = be_sync(comp); if (st) return st;
    st = be_sync(p->d2h_stream); if (st) return st;
    if (bounce_y) {
        std::vector<CopyTask> tasks;
        for (int b = 0; b < 2; b++) { st = drain_tasks(b, tasks); if (st) return st; }
        parallel_copy(tasks, copy_threads);
    }
    return NDCONV_OK;
}
#endif

// A dilated kernel longer than the largest shared-memory FFT tile of its axis has no overlap-save tiling here (a two-level
// transform would be needed).  The reference handles such kernels, so instead of refusing them the float / Complex problem is
// evaluated by the direct kernel: exact tap-by-tap summation, O(outputs x taps) -- slow but inside the conv_fft tolerance.
static bool kernel_exceeds_fft_tiles(const ndconv_problem *pr)
{
    if (!pr || pr->ndim < 1 || pr->ndim > NDC_MAX_DIM || !dtype_is_float(pr->dtype)) return false;
    const bool is_cx = dtype_is_complex(pr->dtype), is_dbl = (pr->dtype == NDCONV_F64 || pr->dtype == NDCONV_C64);
    for (int a = 0; a < pr->ndim; a++) {
        if (pr->kernel_shape[a] < 1 || pr->dilation[a] < 1) return false;       // malformed: let the FFT path report it
        const int64_t Kd = (pr->kernel_shape[a] - 1) * pr->dilation[a] + 1;
        const int cap = a == pr->ndim - 1 ? cap_last_axis(is_cx, is_dbl) : cap_col_axis(is_dbl);
        if (Kd > cap) return true;
    }
    return false;
}

static int conv_direct_impl(ndconv_processor *p, const ndconv_problem *pr, void *out);

static int conv_fft_impl(ndconv_processor *p, const ndconv_problem *pr, void *out)
{
    if (kernel_exceeds_fft_tiles(pr)) return conv_direct_impl(p, pr, out);
#ifdef NDCONV_CUDA
    if (pr && pr->memory == NDCONV_MEM_HOST) {
        Geom g; std::vector<int32_t> maps[NDC_MAX_DIM];
        int st0 = check_problem(pr, NDCONV_PATH_FFT, &g, maps); if (st0) return st0;
        if (!out) { set_error("null output pointer"); return NDCONV_ERR_BAD_ARG; }
        if (pipeline_eligible(pr, g)) { st0 = set_device(p); if (st0) return st0; return conv_fft_host_pipelined(p, pr, g, maps[0], out); }
    }
#endif
    PlanEntry *pe = nullptr;
    int st = get_plan_entry(p, pr, NDCONV_PATH_FFT, &pe); if (st) return st;
    if (!out) { set_error("null output pointer"); return NDCONV_ERR_BAD_ARG; }
    st = set_device(p); if (st) return st;
    if (pe->g.dtype == NDCONV_F32 || pe->g.dtype == NDCONV_C32) return conv_fft_t<float>(p, pr, pe, out);
    return conv_fft_t<double>(p, pr, pe, out);
}

// ======================================================================================================
// public N-d FFT of the processors: Processor::{forward, backward} (src/conv_fft/processor/mod.rs:91-118), with the
// reference's rotated spectrum layout (real.rs:126-154, complex.rs:48-53): axis 0 ends up last.
// Any length (rustfft / realfft take any, real.rs:40,62, complex.rs:56): an axis that is {2,3,5,7}-smooth and fits one
// shared-memory transform (even, when it is the real axis) runs in the row / column kernels of the convolution pipeline;
// every other axis -- longer, odd real, prime factors above 7 -- runs as global-memory Stockham passes (GPassBody).
// ======================================================================================================
// radices of the global passes: register butterflies first, every remaining prime factor as a generic pass
static std::vector<int> global_radices(int64_t n, bool is_dbl)
{
    std::vector<int> r;
    if (!is_dbl) while (n % 16 == 0 && n != 32) { r.push_back(16); n /= 16; }
    while (n % 8 == 0) { r.push_back(8); n /= 8; }
    while (n % 4 == 0) { r.push_back(4); n /= 4; }
    while (n % 2 == 0) { r.push_back(2); n /= 2; }
    for (int64_t f = 3; f * f <= n; f += 2) while (n % f == 0) { r.push_back((int)f); n /= f; }
    if (n > 1) r.push_back((int)n);
    return r;
}

// transform axis (outer, n, inner) of `*cur` into `*alt` pass by pass; on return *cur holds the result
template <class R>
static int run_global_axis(ndconv_processor *p, int64_t n, int64_t inner, int64_t outer, bool inverse, cx<R> **cur, cx<R> **alt, const char *name)
{
    if (n <= 1) return NDCONV_OK;
    const cx<R> *tw = nullptr;
    int st = get_tw_c<R>(p, (int)n, &tw); if (st) return st;
    int64_t Ns = 1;
    for (int r : global_radices(n, sizeof(R) == 8)) {
        GPassParams<R> gp; memset(&gp, 0, sizeof(gp));
        gp.in = *cur; gp.out = *alt; gp.n = n; gp.m = n / r; gp.Ns = Ns; gp.radix = r; gp.inner = inner; gp.outer = outer; gp.tw = tw; gp.inv = inverse ? 1 : 0;
        const bool reg = r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16;
        gp.nwork = outer * inner * (reg ? gp.m : n);
        const int64_t grid = std::min<int64_t>((gp.nwork + 255) / 256, (int64_t)p->num_sms * 16 * kMaxGridMult);
        st = launch<GPassBody<R>, GPassParams<R>>(p->lc(), name, 2.0 * (double)(outer * n * inner) * sizeof(cx<R>), grid, 256, 0, gp); if (st) return st;
        std::swap(*cur, *alt);
        Ns *= r;
    }
    return NDCONV_OK;
}

template <class R>
static int run_gmove(ndconv_processor *p, int mode, const void *src, void *dst, int64_t rows, int64_t n, int64_t spitch, int64_t dpitch, int64_t H, R scale, const char *name)
{
    GMoveParams<R> mp; memset(&mp, 0, sizeof(mp));
    mp.src = src; mp.dst = dst; mp.rows = rows; mp.n = n; mp.spitch = spitch; mp.dpitch = dpitch; mp.H = H; mp.mode = mode; mp.scale = scale;
    const int64_t total = rows * dpitch;
    const int64_t grid = std::min<int64_t>((total + 255) / 256, (int64_t)p->num_sms * 16 * kMaxGridMult);
    return launch<GMoveBody<R>, GMoveParams<R>>(p->lc(), name, (double)(rows * (spitch + dpitch)) * sizeof(cx<R>), grid, 256, 0, mp);
}

template <class R>
static int fft_nd_t(ndconv_processor *p, bool is_cx, int N, const int64_t *shape, const void *in, void *out, int memory, bool inverse)
{
    const bool is_dbl = sizeof(R) == 8;
    int64_t total = 1;
    for (int a = 0; a < N; a++) {
        if (shape[a] < 1) { set_error("fft: empty axis"); return NDCONV_ERR_DATA_SHAPE; }
        if (shape[a] > std::numeric_limits<int32_t>::max()) { set_error("fft: axis longer than 2^31-1"); return NDCONV_ERR_UNSUPPORTED; }
        total *= shape[a];
    }
    FftPlan pl; pl.N = N; pl.is_cx = is_cx;
    bool in_smem[NDC_MAX_DIM], any_col_global = false;
    for (int a = 0; a < N; a++) {
        const bool last = a == N - 1, real_axis = last && !is_cx;
        const int cap = last ? cap_last_axis(is_cx, is_dbl) : cap_col_axis(is_dbl);
        pl.tl[a].F = (int)shape[a]; pl.tl[a].V = (int)shape[a]; pl.tl[a].ntiles = 1;
        in_smem[a] = shape[a] <= cap && !(real_axis && (shape[a] & 1)) &&
                     factor_radices(real_axis ? (int)shape[a] / 2 : (int)shape[a], &pl.fl[a], is_dbl ? 16 : 32);
        if (!last && !in_smem[a]) any_col_global = true;
    }
    const int64_t Fl = shape[N - 1];
    const bool last_global = !in_smem[N - 1];
    pl.H = (int)(is_cx ? Fl : Fl / 2 + 1);
    pl.Hp = (int)align_up((size_t)pl.H, 16);
    pl.rows_per_tile = 1;
    for (int a = 0; a < N - 1; a++) pl.rows_per_tile *= pl.tl[a].F;
    pl.tile_elems = pl.rows_per_tile * pl.Hp; pl.ntiles_total = 1;
    const int64_t rows = pl.rows_per_tile;
    const int es = (int)sizeof(R) * (is_cx ? 2 : 1);
    const size_t real_bytes = (size_t)total * es, spec_elems = (size_t)rows * pl.H, spec_bytes = spec_elems * sizeof(cx<R>);
    int st = set_device(p); if (st) return st;
    // one reservation: spectra workspace [rows][Hp], its ping-pong twin when a strided axis takes the global passes, and two
    // dense [rows][Fl] complex buffers when the last axis does
    const size_t ws_elems = (size_t)pl.tile_elems, row_elems = last_global ? (size_t)rows * (size_t)Fl : 0;
    st = p->ws.reserve((ws_elems * (any_col_global ? 2 : 1) + 2 * row_elems) * sizeof(cx<R>)); if (st) return st;
    cx<R> *ws_a = (cx<R> *)p->ws.p, *ws_b = any_col_global ? ws_a + ws_elems : nullptr;
    cx<R> *row_a = ws_a + ws_elems * (any_col_global ? 2 : 1), *row_b = row_a + row_elems;
    // identity border maps (no padding)
    Geom g; g.ndim = N; g.es = es;
    std::vector<int32_t> maps[NDC_MAX_DIM];
    for (int a = 0; a < N; a++) {
        maps[a].resize(last_global ? 1 : (size_t)shape[a]);       // the row kernels are the only readers of the maps
        for (size_t i = 0; i < maps[a].size(); i++) maps[a][i] = (int32_t)i;
    }
    MetaLayout ml;
    st = upload_meta(p, p->kmeta, g, maps, nullptr, &ml); if (st) return st;
    st = be_sync(p->stream); if (st) return st;
    // staging for host callers
    const void *dev_in = in; void *dev_out = out;
    const size_t in_bytes = inverse ? spec_bytes : real_bytes, out_bytes = inverse ? real_bytes : spec_bytes;
    if (memory == NDCONV_MEM_HOST) {
        st = p->in_stage.reserve(in_bytes); if (st) return st;
        st = p->out_stage.reserve(out_bytes); if (st) return st;
        st = be_h2d(p->in_stage.p, in, in_bytes, p->stream); if (st) return st;
        dev_in = p->in_stage.p; dev_out = p->out_stage.p;
    }
    RowParams<R> rp; memset(&rp, 0, sizeof(rp));
    rp.ndim = N; rp.is_cx = is_cx ? 1 : 0;
    int64_t xs = 1;
    for (int a = N - 1; a >= 0; a--) {
        rp.n[a] = shape[a]; rp.xstr[a] = xs; xs *= shape[a]; rp.P[a] = shape[a];
        rp.map[a] = (const int32_t *)((const unsigned char *)p->kmeta.p + ml.map_off[a]);
        rp.F[a] = pl.tl[a].F; rp.V[a] = pl.tl[a].F; rp.ntiles[a] = 1; rp.Kd[a] = 1; rp.s[a] = 1; rp.O[a] = shape[a];
    }
    rp.ws = ws_a; rp.H = pl.H; rp.Hp = pl.Hp; rp.rows_per_tile = pl.rows_per_tile; rp.tile_elems = pl.tile_elems;
    if (!last_global) {
        st = fill_plan_dev<R>(p, pl.fl[N - 1], &rp.plan); if (st) return st;
        if (!is_cx) { st = get_tw_r<R>(p, (int)Fl, &rp.twr); if (st) return st; }
    }
    PermuteParams<R> pp; pp.n0 = N > 1 ? shape[0] : 1; pp.rest_rows = N > 1 ? pl.rows_per_tile / shape[0] : 1; pp.H = pl.H; pp.Hp = pl.Hp;
    const int64_t pgrid = std::min<int64_t>((int64_t)(spec_elems + 255) / 256, (int64_t)p->num_sms * 16 * kMaxGridMult);
    const double sb = (double)spec_bytes;
    // strided axis a of the workspace: inner = Hp * prod F[b > a], outer = prod F[b < a]
    auto col_axis = [&](int a, bool inv, cx<R> **cur, cx<R> **alt) -> int {
        if (in_smem[a]) return run_col<R>(p, pl, a, inv ? 1 : 0, *cur, nullptr, 1, inv ? "fft_col_inv" : "fft_col_fwd", 2 * sb);
        int64_t inner = pl.Hp, outer = 1;
        for (int b = a + 1; b < N - 1; b++) inner *= shape[b];
        for (int b = 0; b < a; b++) outer *= shape[b];
        return run_global_axis<R>(p, shape[a], inner, outer, inv, cur, alt, inv ? "fft_gpass_inv" : "fft_gpass_fwd");
    };
    cx<R> *cur = ws_a, *alt = ws_b;
    if (!inverse) {
        if (!last_global) {
            rp.x = dev_in;
            st = run_row<R>(p, 0, rp, 0, "fft_row_fwd", (double)real_bytes + sb); if (st) return st;
        } else {
            cx<R> *rc = row_a, *ra = row_b;
            st = run_gmove<R>(p, is_cx ? 1 : 0, dev_in, rc, rows, Fl, Fl, Fl, Fl, (R)0, "fft_row_pack"); if (st) return st;
            st = run_global_axis<R>(p, Fl, 1, rows, false, &rc, &ra, "fft_gpass_fwd"); if (st) return st;
            st = run_gmove<R>(p, 1, rc, ws_a, rows, Fl, Fl, pl.Hp, pl.H, (R)0, "fft_row_pack"); if (st) return st;
        }
        for (int a = N - 2; a >= 0; a--) { st = col_axis(a, false, &cur, &alt); if (st) return st; }
        pp.src = cur; pp.dst = (cx<R> *)dev_out; pp.to_rotated = 1;
        st = launch<PermuteBody<R>, PermuteParams<R>>(p->lc(), "fft_layout_rotate", 2 * sb, pgrid, 256, 0, pp); if (st) return st;
    } else {
        pp.src = (const cx<R> *)dev_in; pp.dst = ws_a; pp.to_rotated = 0;
        st = launch<PermuteBody<R>, PermuteParams<R>>(p->lc(), "fft_layout_rotate", 2 * sb, pgrid, 256, 0, pp); if (st) return st;
        for (int a = 0; a <= N - 2; a++) { st = col_axis(a, true, &cur, &alt); if (st) return st; }
        const R scale = (R)(1.0L / (long double)total);                          // real.rs:278-279, complex.rs:141-142
        if (!last_global) {
            rp.ws = cur; rp.out = dev_out; rp.scale = scale;
            st = run_row<R>(p, 1, rp, pl.rows_per_tile, "fft_row_inv", (double)real_bytes + sb); if (st) return st;
        } else {
            cx<R> *rc = row_a, *ra = row_b;
            st = run_gmove<R>(p, is_cx ? 1 : 2, cur, rc, rows, Fl, pl.Hp, Fl, pl.H, (R)0, "fft_row_pack"); if (st) return st;
            st = run_global_axis<R>(p, Fl, 1, rows, true, &rc, &ra, "fft_gpass_inv"); if (st) return st;
            st = run_gmove<R>(p, is_cx ? 4 : 3, rc, dev_out, rows, Fl, Fl, Fl, Fl, scale, "fft_row_pack"); if (st) return st;
        }
    }
    if (memory == NDCONV_MEM_HOST) {
        st = be_d2h(out, dev_out, out_bytes, p->stream); if (st) return st;
        st = be_sync(p->stream); if (st) return st;
    }
    return NDCONV_OK;
}

static int fft_nd(ndconv_processor *p, int dtype, int ndim, const int64_t *shape, const void *in, void *out, int memory, bool inverse)
{
    if (!p || !shape || !in || !out || ndim < 1 || ndim > NDC_MAX_DIM) { set_error("fft: bad arguments"); return NDCONV_ERR_BAD_ARG; }
    switch (dtype) {
    case NDCONV_F32: return fft_nd_t<float>(p, false, ndim, shape, in, out, memory, inverse);
    case NDCONV_F64: return fft_nd_t<double>(p, false, ndim, shape, in, out, memory, inverse);
    case NDCONV_C32: return fft_nd_t<float>(p, true, ndim, shape, in, out, memory, inverse);
    case NDCONV_C64: return fft_nd_t<double>(p, true, ndim, shape, in, out, memory, inverse);
    }
    set_error("fft: dtype must be f32/f64/Complex (integer FFT is documented as broken in the reference)");
    return NDCONV_ERR_UNSUPPORTED;
}

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

const char *ndconv_version(void) { return NDCONV_VERSION_STRING; }
int ndconv_is_emulation(void)
{
#ifdef NDCONV_CUDA
    return 0;
#else
    return 1;
#endif
}
const char *ndconv_last_error_string(void) { return get_error(); }
const char *ndconv_status_string(int s)
{
    switch (s) {
    case NDCONV_OK: return "ok";
    case NDCONV_ERR_DATA_SHAPE: return "DataShape";
    case NDCONV_ERR_KERNEL_SHAPE: return "KernelShape";
    case NDCONV_ERR_MISMATCH_SHAPE: return "MismatchShape";
    case NDCONV_ERR_PANIC: return "ReferencePanic";
    case NDCONV_ERR_BAD_ARG: return "BadArgument";
    case NDCONV_ERR_UNSUPPORTED: return "Unsupported";
    case NDCONV_ERR_CUDA: return "CudaError";
    case NDCONV_ERR_INTERNAL: return "InternalError";
    }
    return "unknown";
}
size_t ndconv_dtype_size(int dtype) { return dtype_size(dtype); }

int ndconv_device_count(void)
{
#ifdef NDCONV_CUDA
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_error(std::string("CUDA error: ") + cudaGetErrorString(e)); cudaGetLastError(); return -NDCONV_ERR_CUDA; }
    return n;
#else
    return 1;
#endif
}

int ndconv_unfold_conv_mode(int mode, int ndim, const int64_t *kernel_shape, const int64_t *dilation, const int64_t *padding,
                            const int64_t *strides, int64_t out_pad[][2], int64_t *out_stride)
{
    return unfold_mode(mode, ndim, kernel_shape, dilation, padding, strides, out_pad, out_stride);
}
int64_t ndconv_good_fft_size(int64_t n) { return good_size_cc(n); }
int64_t ndconv_plan_fft_size(int64_t n, int real_axis) { return smooth_ge(n, real_axis != 0); }

int ndconv_out_shape(const ndconv_problem *problem, int path, int64_t *out_shape)
{
    Geom g; std::vector<int32_t> maps[NDC_MAX_DIM];
    int st = check_problem(problem, path, &g, maps); if (st) return st;
    if (out_shape) for (int i = 0; i < g.ndim; i++) out_shape[i] = g.O[i];
    return NDCONV_OK;
}

int ndconv_border_index_map(int64_t n, int64_t pad_front, int64_t pad_back, int border_front, int border_back, int32_t *out_map)
{
    std::vector<int32_t> m;
    int st = build_border_map(n, pad_front, pad_back, border_front, border_back, m); if (st) return st;
    if (out_map) memcpy(out_map, m.data(), m.size() * sizeof(int32_t));
    return NDCONV_OK;
}

int ndconv_processor_create(int device, ndconv_processor **out)
{
    if (!out) { set_error("null out"); return NDCONV_ERR_BAD_ARG; }
    *out = nullptr;
    std::unique_ptr<ndconv_processor> p(new ndconv_processor());
    p->device = device;
#ifdef NDCONV_CUDA
    int n = 0;
    CU_CHECK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) { set_error("no such CUDA device"); return NDCONV_ERR_CUDA; }
    CU_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error(std::string("device is not sm_100 (Blackwell B200): ") + prop.name + "; this library has no other code path"); return NDCONV_ERR_CUDA; }
    CU_CHECK(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
#endif
    p->stream = p->own_stream;
    p->num_sms = be_num_sms(device);
    *out = p.release();
    return NDCONV_OK;
}

int ndconv_processor_destroy(ndconv_processor *p)
{
    if (!p) return NDCONV_OK;
    set_device(p);
    be_sync(p->stream);
    p->ws.release(); p->in_stage.release(); p->out_stage.release(); p->meta.release(); p->kb_stage.release(); p->kmeta.release();
    for (int b = 0; b < 2; b++) { p->pipe_in[b].release(); p->pipe_out[b].release(); }
    p->pipe_row.release(); p->zero_map.release(); p->ws_aux.release();
#ifdef NDCONV_CUDA
    if (p->aux_stream) { cudaStreamSynchronize(p->aux_stream); cudaStreamDestroy(p->aux_stream); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (int b = 0; b < 2; b++) { if (p->bounce_in[b]) cudaFreeHost(p->bounce_in[b]); if (p->bounce_out[b]) cudaFreeHost(p->bounce_out[b]); }
    if (p->h2d_stream) cudaStreamDestroy(p->h2d_stream);
    if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
    for (int b = 0; b < 2; b++) { if (p->ev_h2d[b]) cudaEventDestroy(p->ev_h2d[b]); if (p->ev_comp[b]) cudaEventDestroy(p->ev_comp[b]); if (p->ev_d2h[b]) cudaEventDestroy(p->ev_d2h[b]); }
#endif
    for (auto &kv : p->tw_c) be_free(kv.second);
    for (auto &kv : p->tw_r) be_free(kv.second);
    for (auto &k : p->kspecs) { k->buf.release(); k->pair.release(); }
#ifdef NDCONV_CUDA
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
#endif
    delete p;
    return NDCONV_OK;
}

int ndconv_processor_set_stream(ndconv_processor *p, void *cuda_stream)
{
    if (!p) { set_error("null processor"); return NDCONV_ERR_BAD_ARG; }
    p->stream = cuda_stream ? (stream_t)cuda_stream : p->own_stream;
    return NDCONV_OK;
}
int ndconv_processor_synchronize(ndconv_processor *p)
{
    if (!p) { set_error("null processor"); return NDCONV_ERR_BAD_ARG; }
    int st = set_device(p); if (st) return st;
    return be_sync(p->stream);
}
int64_t ndconv_processor_launch_count(const ndconv_processor *p) { return p ? p->launches : 0; }
int64_t ndconv_processor_workspace_bytes(const ndconv_processor *p) { return p ? (int64_t)p->held() : 0; }

int ndconv_processor_set_profiling(ndconv_processor *p, int enable)
{
    if (!p) { set_error("null processor"); return NDCONV_ERR_BAD_ARG; }
    p->prof.on = enable != 0;
    return NDCONV_OK;
}

int ndconv_processor_get_profile(ndconv_processor *p, int max_entries, char (*names)[64], double *total_ms, int64_t *launches, double *alg_bytes)
{
    if (!p) return 0;
    int n = 0;
#ifdef NDCONV_CUDA
    set_device(p);
    cudaStreamSynchronize(p->stream);
    for (auto &r : p->prof.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        int i = 0;
        for (; i < n; i++) if (!strncmp(names[i], r.name, 63)) break;
        if (i == n) {
            if (n >= max_entries) continue;
            strncpy(names[n], r.name, 63); names[n][63] = 0; total_ms[n] = 0; launches[n] = 0; alg_bytes[n] = 0; n++;
        }
        total_ms[i] += ms; launches[i] += 1; alg_bytes[i] += r.bytes;
        p->prof.pool.push_back(r.a); p->prof.pool.push_back(r.b);
    }
    p->prof.recs.clear();
#else
    (void)max_entries; (void)names; (void)total_ms; (void)launches; (void)alg_bytes;
#endif
    return n;
}

static int with_processor(ndconv_processor *p, const ndconv_problem *pr, void *out, int (*fn)(ndconv_processor *, const ndconv_problem *, void *))
{
    if (p) return fn(p, pr, out);
    if (pr && pr->memory == NDCONV_MEM_DEVICE) { set_error("device-resident problems need a processor"); return NDCONV_ERR_BAD_ARG; }
    // validate before touching the device so shape errors surface exactly as in the reference
    if (pr) { Geom g; std::vector<int32_t> maps[NDC_MAX_DIM]; int st = check_problem(pr, fn == conv_direct_impl ? NDCONV_PATH_DIRECT : NDCONV_PATH_FFT, &g, maps); if (st) return st; }
    ndconv_processor *tmp = nullptr;
    int st = ndconv_processor_create(0, &tmp); if (st) return st;
    st = fn(tmp, pr, out);
    std::string keep = get_error();
    ndconv_processor_destroy(tmp);
    set_error(keep);
    return st;
}

int ndconv_conv_direct(ndconv_processor *p, const ndconv_problem *problem, void *out) { return with_processor(p, problem, out, conv_direct_impl); }
int ndconv_conv_fft(ndconv_processor *p, const ndconv_problem *problem, void *out) { return with_processor(p, problem, out, conv_fft_impl); }
int ndconv_conv_fft_par(ndconv_processor *p, const ndconv_problem *problem, void *out) { return with_processor(p, problem, out, conv_fft_impl); }

int ndconv_fft_forward(ndconv_processor *p, int dtype, int ndim, const int64_t *shape, const void *in, void *out, int memory)
{
    return fft_nd(p, dtype, ndim, shape, in, out, memory, false);
}
int ndconv_fft_backward(ndconv_processor *p, int dtype, int ndim, const int64_t *shape, const void *spectrum, void *out, int memory)
{
    return fft_nd(p, dtype, ndim, shape, spectrum, out, memory, true);
}

int ndconv_slab_plan(const ndconv_problem *problem, int path, int n_slabs, int slab, ndconv_slab *out)
{
    Geom g; std::vector<int32_t> maps[NDC_MAX_DIM];
    int st = check_problem(problem, path, &g, maps); if (st) return st;
    return slab_plan(g, n_slabs, slab, out);
}

// One host-resident convolution over several GPUs of this process (conv_fft_par with more than one device
// configured): the output rows of axis 0 are split into one contiguous slab per handle (slab_plan), and every handle runs the
// pipelined host path on its own rows from its own host thread -- H2D | kernels | D2H per GPU, each over its own PCIe link, no
// data-path collective.  Handles may live on the same device (the rows are then interleaved on that device's streams).
int ndconv_conv_fft_sharded(ndconv_processor *const *handles, int n_handles, const ndconv_problem *problem, void *out)
{
    if (!handles || n_handles < 1) { set_error("conv_fft_sharded: no processors"); return NDCONV_ERR_BAD_ARG; }
    for (int i = 0; i < n_handles; i++) if (!handles[i]) { set_error("conv_fft_sharded: null processor"); return NDCONV_ERR_BAD_ARG; }
    if (problem && problem->memory != NDCONV_MEM_HOST) { set_error("conv_fft_sharded: the problem must be host-resident (device-resident slabs: one call per rank + halo exchange)"); return NDCONV_ERR_BAD_ARG; }
#ifdef NDCONV_CUDA
    if (n_handles > 1 && problem && !kernel_exceeds_fft_tiles(problem)) {
        Geom g; std::vector<int32_t> maps[NDC_MAX_DIM];
        int st = check_problem(problem, NDCONV_PATH_FFT, &g, maps); if (st) return st;
        if (!out) { set_error("null output pointer"); return NDCONV_ERR_BAD_ARG; }
        if (pipeline_eligible(problem, g) && g.O[0] >= 4 * (int64_t)n_handles) {
            std::vector<int> status((size_t)n_handles, NDCONV_OK);
            std::vector<std::string> message((size_t)n_handles);
            std::vector<std::thread> workers;
            for (int i = 0; i < n_handles; i++) {
                workers.emplace_back([&, i]() {
                    ndconv_slab sl;
                    int s2 = slab_plan(g, n_handles, i, &sl);
                    if (!s2) s2 = set_device(handles[i]);
                    if (!s2) s2 = conv_fft_host_pipelined(handles[i], problem, g, maps[0], out, sl.out_begin, sl.out_end);
                    status[(size_t)i] = s2;
                    if (s2) message[(size_t)i] = get_error();       // the error string is thread-local
                });
            }
            for (auto &w : workers) w.join();
            for (int i = 0; i < n_handles; i++) if (status[(size_t)i]) { set_error(message[(size_t)i]); return status[(size_t)i]; }
            return NDCONV_OK;
        }
    }
#endif
    return conv_fft_impl(handles[0], problem, out);
}

// Independent convolutions distributed whole (north star: "batched independent convolutions are distributed whole"):
// problem i runs on handle i % n_processors, every handle driven by its own host thread on its own stream and workspace.  Handles on
// one device overlap their (latency-bound, less-than-a-wave) launches; handles on several devices spread the batch over the GPUs.
// The first failing status is returned (problems are independent: the others still run).
int ndconv_conv_fft_batch(ndconv_processor *const *handles, int n_handles, const ndconv_problem *problems, void *const *outs, int n_problems)
{
    if (!handles || n_handles < 1 || n_problems < 0 || (n_problems > 0 && (!problems || !outs))) { set_error("conv_fft_batch: bad arguments"); return NDCONV_ERR_BAD_ARG; }
    for (int i = 0; i < n_handles; i++) if (!handles[i]) { set_error("conv_fft_batch: null processor"); return NDCONV_ERR_BAD_ARG; }
    const int nt = std::min(n_handles, n_problems);
    std::vector<int> status((size_t)std::max(nt, 1), NDCONV_OK);
    std::vector<std::string> message((size_t)std::max(nt, 1));
    auto work = [&](int h) {
        for (int i = h; i < n_problems; i += n_handles) {
            const int st = conv_fft_impl(handles[h], &problems[i], outs[i]);
            if (st && !status[(size_t)h]) { status[(size_t)h] = st; message[(size_t)h] = get_error(); }      // the error string is thread-local
        }
        // device-resident problems are only enqueued: the caller synchronises the processors (ndconv_processor_synchronize)
    };
    if (nt <= 1) { if (nt == 1) work(0); }
    else {
        std::vector<std::thread> workers;
        for (int h = 0; h < nt; h++) workers.emplace_back(work, h);
        for (auto &w : workers) w.join();
    }
    for (int h = 0; h < nt; h++) if (status[(size_t)h]) { set_error(message[(size_t)h]); return status[(size_t)h]; }
    return NDCONV_OK;
}

void *ndconv_host_alloc(size_t bytes)
{
#ifdef NDCONV_CUDA
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); set_error("cudaHostAlloc failed"); return nullptr; }
    return p;
#else
    return malloc(bytes ? bytes : 1);
#endif
}
void ndconv_host_free(void *ptr)
{
#ifdef NDCONV_CUDA
    if (ptr) cudaFreeHost(ptr);
#else
    free(ptr);
#endif
}

// page-lock an allocation the caller already owns (an ndarray's Vec): host calls on it then run at the pinned rate.  Registering
// costs ~0.1 ms per MB (measured, tools/pageable_probe.py) -- worth it for a buffer used by more than ~2 calls, not per call.
int ndconv_host_register(void *ptr, size_t bytes)
{
    if (!ptr || !bytes) { set_error("host_register: bad arguments"); return NDCONV_ERR_BAD_ARG; }
#ifdef NDCONV_CUDA
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); set_error(std::string("cudaHostRegister failed: ") + cudaGetErrorString(e)); return NDCONV_ERR_CUDA; }
#endif
    return NDCONV_OK;
}
int ndconv_host_unregister(void *ptr)
{
    if (!ptr) { set_error("host_unregister: bad arguments"); return NDCONV_ERR_BAD_ARG; }
#ifdef NDCONV_CUDA
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); set_error(std::string("cudaHostUnregister failed: ") + cudaGetErrorString(e)); return NDCONV_ERR_CUDA; }
#endif
    return NDCONV_OK;
}

}  // extern "C"
