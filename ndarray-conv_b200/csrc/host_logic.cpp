// host_logic.cpp -- see host_logic.h.  Citations are relative to the reference root.
#include "host_logic.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace ndc {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *get_error() { return g_err.c_str(); }

size_t dtype_size(int dtype)
{
    switch (dtype) {
    case NDCONV_I32: case NDCONV_F32: case NDCONV_U32: return 4;
    case NDCONV_I64: case NDCONV_F64: case NDCONV_C32: case NDCONV_U64: return 8;
    case NDCONV_C64: case NDCONV_I128: case NDCONV_U128: return 16;
    case NDCONV_I8: case NDCONV_U8: return 1;
    case NDCONV_I16: case NDCONV_U16: return 2;
    }
    return 0;
}
bool dtype_is_float(int dtype) { return dtype == NDCONV_F32 || dtype == NDCONV_F64 || dtype == NDCONV_C32 || dtype == NDCONV_C64; }
bool dtype_is_complex(int dtype) { return dtype == NDCONV_C32 || dtype == NDCONV_C64; }

// ConvMode::unfold, src/conv/mod.rs:28-66
int unfold_mode(int mode, int ndim, const int64_t *kshape, const int64_t *dil, const int64_t *padding,
                const int64_t *strides, int64_t out_pad[][2], int64_t *out_stride)
{
    if (ndim < 1 || ndim > NDC_MAX_DIM || !kshape || !dil || !out_pad || !out_stride) { set_error("unfold: bad arguments"); return NDCONV_ERR_BAD_ARG; }
    if ((mode == NDCONV_MODE_CUSTOM || mode == NDCONV_MODE_EXPLICIT) && (!padding || !strides)) { set_error("unfold: Custom/Explicit need padding and strides"); return NDCONV_ERR_BAD_ARG; }
    for (int i = 0; i < ndim; i++) {
        int64_t kd = kshape[i] * dil[i] - dil[i] + 1;   // :35-37
        switch (mode) {
        case NDCONV_MODE_FULL: out_pad[i][0] = out_pad[i][1] = kd - 1; out_stride[i] = 1; break;   // :40-43
        case NDCONV_MODE_SAME:                                                                     // :44-55
            if (kd % 2 == 0) { out_pad[i][0] = (kd - 1) / 2 + 1; out_pad[i][1] = (kd - 1) / 2; }
            else out_pad[i][0] = out_pad[i][1] = (kd - 1) / 2;
            out_stride[i] = 1; break;
        case NDCONV_MODE_VALID: out_pad[i][0] = out_pad[i][1] = 0; out_stride[i] = 1; break;         // :56-59
        case NDCONV_MODE_CUSTOM: out_pad[i][0] = out_pad[i][1] = padding[i]; out_stride[i] = strides[i]; break;   // :60-63
        case NDCONV_MODE_EXPLICIT: out_pad[i][0] = padding[2 * i]; out_pad[i][1] = padding[2 * i + 1]; out_stride[i] = strides[i]; break;  // :64
        default: set_error("unfold: unknown ConvMode"); return NDCONV_ERR_BAD_ARG;
        }
    }
    return NDCONV_OK;
}

// src/conv_fft/good_size.rs:6-31
int64_t good_size_cc(int64_t n)
{
    int64_t best = 1;
    while (best < n) best <<= 1;
    for (;;) { int64_t f = best / 4 * 3; if (f < n) break; if (f == n) return n; best = f; }
    for (;;) { int64_t f = best / 6 * 5; if (f < n) break; if (f == n) return n; best = f; }
    return best;
}

// Symbolic replay of src/padding/half_dim.rs on one axis.  sym[i] = data index the plane i ends up
// holding, or a code: CONST_FRONT / CONST_BACK (a filled plane), INIT (never written).
int build_border_map(int64_t n, int64_t pf, int64_t pb, int bf, int bb, std::vector<int32_t> &map)
{
    if (n < 1 || pf < 0 || pb < 0) { set_error("border map: bad extents"); return NDCONV_ERR_BAD_ARG; }
    int64_t P = n + pf + pb;
    if (P > std::numeric_limits<int32_t>::max()) { set_error("axis longer than 2^31-1"); return NDCONV_ERR_UNSUPPORTED; }
    map.assign((size_t)P, NDC_MAP_INIT);
    for (int64_t i = 0; i < n; i++) map[(size_t)(pf + i)] = (int32_t)i;   // padding_const, src/padding/mod.rs:175-200
    for (int64_t j = 0; j < pf; j++) {                                      // *_front loops run j ascending
        switch (bf) {
        case NDCONV_BORDER_ZEROS: case NDCONV_BORDER_CONST: map[j] = NDC_MAP_CONST_FRONT; break;   // half_dim.rs:30-49
        case NDCONV_BORDER_REPLICATE: map[j] = map[pf]; break;                                     // :113-129
        case NDCONV_BORDER_REFLECT: {                                                              // :192-211
            int64_t r = (pf - j) + pf;
            if (r >= P) { set_error("Reflect front padding reaches past the padded axis: the reference panics (half_dim.rs:200-208)"); return NDCONV_ERR_PANIC; }
            map[j] = map[r]; break; }
        case NDCONV_BORDER_CIRCULAR: { int64_t r = P - pb - (pf - j); map[j] = map[r]; break; }    // :277-296
        default: set_error("unknown BorderType"); return NDCONV_ERR_BAD_ARG;
        }
    }
    int64_t bi = P - pb - 1;
    for (int64_t j = n + pf; j < P; j++) {
        switch (bb) {
        case NDCONV_BORDER_ZEROS: case NDCONV_BORDER_CONST: map[j] = NDC_MAP_CONST_BACK; break;    // :72-94
        case NDCONV_BORDER_REPLICATE: map[j] = map[bi]; break;                                     // :151-173
        case NDCONV_BORDER_REFLECT:                                                                // :233-258
            if (j - bi > bi) { set_error("Reflect back padding underflows: the reference panics (half_dim.rs:246)"); return NDCONV_ERR_PANIC; }
            map[j] = map[bi - (j - bi)]; break;
        case NDCONV_BORDER_CIRCULAR: map[j] = map[pf + (j - bi - 1)]; break;                       // :318-343
        default: set_error("unknown BorderType"); return NDCONV_ERR_BAD_ARG;
        }
    }
    return NDCONV_OK;
}

int check_problem(const ndconv_problem *pr, int path, Geom *g, std::vector<int32_t> maps[NDC_MAX_DIM])
{
    if (!pr || !g) { set_error("null problem"); return NDCONV_ERR_BAD_ARG; }
    if (pr->ndim < 1 || pr->ndim > NDC_MAX_DIM) { set_error("ndim must be 1..6"); return NDCONV_ERR_BAD_ARG; }
    size_t es = dtype_size(pr->dtype);
    if (!es) { set_error("unknown dtype"); return NDCONV_ERR_BAD_ARG; }
    if (path == NDCONV_PATH_FFT && !dtype_is_float(pr->dtype)) {
        set_error("conv_fft takes f32/f64/Complex (integer FFT is documented as broken in the reference, processor/mod.rs:44-52)");
        return NDCONV_ERR_UNSUPPORTED;
    }
    g->ndim = pr->ndim; g->dtype = pr->dtype; g->es = (int)es; g->reverse = pr->reverse != 0;
    int N = pr->ndim;
    g->data_total = 1; g->kernel_total = 1;
    for (int i = 0; i < N; i++) {
        if (pr->data_shape[i] < 0 || pr->kernel_shape[i] < 0 || pr->pad[i][0] < 0 || pr->pad[i][1] < 0) { set_error("negative extent"); return NDCONV_ERR_BAD_ARG; }
        g->n[i] = pr->data_shape[i]; g->xstr[i] = pr->data_strides[i];
        g->k[i] = pr->kernel_shape[i]; g->kstr[i] = pr->kernel_strides[i];
        g->d[i] = pr->dilation[i];
        g->pf[i] = pr->pad[i][0]; g->pb[i] = pr->pad[i][1];
        g->s[i] = pr->stride[i];
        g->bf[i] = pr->border[i][0].type; g->bb[i] = pr->border[i][1].type;
        g->data_total *= g->n[i]; g->kernel_total *= g->k[i];
    }
    if (g->data_total == 0) { set_error("Data shape shouldn't have ZERO."); return NDCONV_ERR_DATA_SHAPE; }          // conv/mod.rs:136-139
    if (g->kernel_total == 0) {                                                                                      // conv/mod.rs:141-144
        set_error("Kernel shape shouldn't have ZERO.");
        return path == NDCONV_PATH_FFT ? NDCONV_ERR_DATA_SHAPE : NDCONV_ERR_KERNEL_SHAPE;                            // conv_fft/mod.rs:210-213 (sic)
    }
    if (!pr->data || !pr->kernel) { set_error("null data / kernel pointer"); return NDCONV_ERR_BAD_ARG; }
    for (int i = 0; i < N; i++) {
        if (g->d[i] < 1) { set_error("dilation must be >= 1"); return NDCONV_ERR_PANIC; }
        g->Kd[i] = g->k[i] * g->d[i] - g->d[i] + 1;                                                                  // conv/mod.rs:146-147
        g->P[i] = g->n[i] + g->pf[i] + g->pb[i];
    }
    bool mismatch = false;
    for (int i = 0; i < N; i++) if (g->Kd[i] > g->P[i]) mismatch = true;
    // conv pads first (conv/mod.rs:150) and checks afterwards (:152-158); conv_fft checks first (conv_fft/mod.rs:222-227)
    if (path == NDCONV_PATH_FFT && mismatch) { set_error("ConvMode does not match KernelWithDilation size"); return NDCONV_ERR_MISMATCH_SHAPE; }
    for (int i = 0; i < N; i++) {
        int st = build_border_map(g->n[i], g->pf[i], g->pb[i], g->bf[i], g->bb[i], maps[i]);
        if (st) return st;
    }
    if (mismatch) { set_error("ConvMode does not match KernelWithDilation size"); return NDCONV_ERR_MISMATCH_SHAPE; }
    g->out_total = 1;
    for (int i = 0; i < N; i++) {
        if (g->s[i] < 1) { set_error("stride 0: the reference divides by zero (conv/mod.rs:165)"); return NDCONV_ERR_PANIC; }
        g->O[i] = (g->P[i] - g->Kd[i]) / g->s[i] + 1;                                                                // conv/mod.rs:162-167
        g->out_total *= g->O[i];
    }
    // standard layout?
    int64_t s = 1; bool contig = true;
    for (int i = N - 1; i >= 0; i--) { if (g->n[i] != 1 && g->xstr[i] != s) contig = false; s *= g->n[i]; }
    g->data_contiguous = contig;
    return NDCONV_OK;
}

template <class T> static bool is_zero_t(const void *p) { return *(const T *)p == (T)0; }
static bool weight_is_zero(int dtype, const void *p)
{
    switch (dtype) {
    case NDCONV_F32: return is_zero_t<float>(p);
    case NDCONV_F64: return is_zero_t<double>(p);
    case NDCONV_C32: return ((const float *)p)[0] == 0.0f && ((const float *)p)[1] == 0.0f;
    case NDCONV_C64: return ((const double *)p)[0] == 0.0 && ((const double *)p)[1] == 0.0;
    default: { size_t es = dtype_size(dtype); const unsigned char *c = (const unsigned char *)p; for (size_t i = 0; i < es; i++) if (c[i]) return false; return true; }
    }
}

// src/dilation/mod.rs:34-60
void build_taps(const Geom &g, const void *kernel, Taps &t)
{
    int N = g.ndim;
    t.ntap = 0; t.off.clear(); t.lin.clear(); t.w.clear();
    int64_t idx[NDC_MAX_DIM] = {0};
    const unsigned char *kb = (const unsigned char *)kernel;
    for (int64_t e = 0; e < g.kernel_total; e++) {
        int64_t ko = 0, lin = 0;
        for (int i = 0; i < N; i++) {
            int64_t ki = g.reverse ? g.k[i] - 1 - idx[i] : idx[i];   // slice step -1 when reversed, :35-42
            ko += ki * g.kstr[i];
            lin += idx[i] * g.d[i] * g.xstr[i];                      // :44-45, :52-57 (on the un-padded data strides)
        }
        const unsigned char *w = kb + ko * g.es;
        if (!weight_is_zero(g.dtype, w)) {                           // :49
            for (int i = 0; i < NDC_MAX_DIM; i++) t.off.push_back(i < N ? (int32_t)(idx[i] * g.d[i]) : 0);
            t.lin.push_back(lin);
            t.w.insert(t.w.end(), w, w + g.es);
            t.ntap++;
        }
        for (int i = N - 1; i >= 0; i--) { if (++idx[i] < g.k[i]) break; idx[i] = 0; }
    }
}

// ---- FFT planning --------------------------------------------------------------------------------
bool factor_radices(int L, FftLen *out, int max_radix)
{
    if (L < 1) return false;
    FftLen f; f.L = L; f.npass = 0;
    int r = L;
    // large power-of-two radices first, then odd primes
    // greedy: as few passes as possible (each pass is one shared-memory round trip).  max_radix = 16 for f64: a radix-32
    // butterfly of doubles needs 128 registers for its data alone and spills
    while (max_radix >= 32 && r % 32 == 0 && r != 64 && r != 128 && r != 256 && f.npass < NDC_MAX_PASS) { f.radix[f.npass++] = 32; r /= 32; }
    while (r % 16 == 0 && r != 32 && r != 64 && f.npass < NDC_MAX_PASS) { f.radix[f.npass++] = 16; r /= 16; }
    while (r % 8 == 0 && f.npass < NDC_MAX_PASS) { f.radix[f.npass++] = 8; r /= 8; }
    while (r % 4 == 0 && f.npass < NDC_MAX_PASS) { f.radix[f.npass++] = 4; r /= 4; }
    while (r % 2 == 0 && f.npass < NDC_MAX_PASS) { f.radix[f.npass++] = 2; r /= 2; }
    const int odd[3] = {7, 5, 3};
    for (int p : odd) while (r % p == 0 && f.npass < NDC_MAX_PASS) { f.radix[f.npass++] = p; r /= p; }
    if (r != 1) return false;
    if (out) *out = f;
    return true;
}

static bool is_smooth(int64_t n)
{
    if (n < 1) return false;
    for (int p : {2, 3, 5, 7}) while (n % p == 0) n /= p;
    return n == 1;
}

int64_t smooth_ge(int64_t n, bool even)
{
    if (n < 1) n = 1;
    if (even && n < 2) n = 2;
    for (int64_t f = n;; f++) {
        if (even && (f & 1)) continue;
        if (is_smooth(f)) return f;
    }
}

double fft_len_cost(int F, bool real_axis)
{
    int L = real_axis ? F / 2 : F;
    FftLen fl;
    if (!factor_radices(L, &fl)) return 1e300;
    // one global read + one global write, plus one shared-memory round trip per Stockham pass
    return (double)F * (2.0 + 0.6 * fl.npass);
}

int plan_axis(int64_t P, int64_t Kd, int cap, bool real_axis, AxisTiling *out)
{
    if (Kd > cap) {
        set_error("dilated kernel extent exceeds the largest shared-memory FFT tile on an axis (see DESIGN.md: envelope)");
        return NDCONV_ERR_UNSUPPORTED;
    }
    int64_t M = P - Kd + 1;              // alias-free positions needed: m in [Kd-1, P)
    double best = 1e300; AxisTiling bt;
    int64_t lo = std::max<int64_t>(Kd, real_axis ? 2 : 1);
    for (int64_t F = lo; F <= cap; F++) {
        if (real_axis && (F & 1)) continue;
        if (!is_smooth(F)) continue;
        int64_t V = F - Kd + 1;
        if (V < 1) continue;
        int64_t nt = (M + V - 1) / V;
        if (nt == 1 && F < P) continue;   // a single tile must hold the whole padded axis
        double c = (double)nt * fft_len_cost((int)F, real_axis);
        if (c < best) { best = c; bt.F = (int)F; bt.V = (int)V; bt.ntiles = (int)nt; }
        if (nt == 1 && F >= P && c > best * 1.5) break;
    }
    if (best >= 1e300) { set_error("no FFT tiling found for axis"); return NDCONV_ERR_UNSUPPORTED; }
    *out = bt;
    return NDCONV_OK;
}

// Balanced split of axis-0 output rows; pad_begin/pad_end = rows of the padded axis the slab reads.
int slab_plan(const Geom &g, int n_slabs, int slab, ndconv_slab *out)
{
    if (n_slabs < 1 || slab < 0 || slab >= n_slabs || !out) { set_error("slab_plan: bad arguments"); return NDCONV_ERR_BAD_ARG; }
    int64_t O = g.O[0];
    int64_t base = O / n_slabs, rem = O % n_slabs;
    int64_t b = slab * base + std::min<int64_t>(slab, rem);
    int64_t e = b + base + (slab < rem ? 1 : 0);
    out->out_begin = b; out->out_end = e;
    if (e > b) { out->pad_begin = b * g.s[0]; out->pad_end = (e - 1) * g.s[0] + g.Kd[0]; }   // out[o] reads padded rows o*s .. o*s+Kd-1 (conv/mod.rs:188-196)
    else { out->pad_begin = out->pad_end = 0; }
    return NDCONV_OK;
}

}  // namespace ndc
