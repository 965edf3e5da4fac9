/*
 * ndconv_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the algorithm of TYPEmber/ndarray-conv v0.6.1 for the
 * `conv` / `conv_fft` hot path.  It exists so that the CUDA product can be checked
 * against the reference's semantics in an image that has no Rust toolchain.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may
 * load it.  The product (ndarray-conv_b200/) never links, imports or calls it.
 *
 * Pinning: the reference cannot be compiled here (no rustc/cargo), so this oracle is
 * pinned against the literal known-answer vectors of the reference's own unit tests
 * (tests/golden/reference_kats.json, transcribed from src/padding/mod.rs:468-680,
 * src/conv_fft/padding.rs:124-181, src/dilation/mod.rs:278-376, src/conv/tests.rs:555-583,
 * src/conv_fft/tests.rs:209-237) and against torch-CPU generated vectors for the
 * libtorch-derived tests (tests/golden/make_torch_golden.py).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  Compile with -ffp-contract=off: the reference's Rust never fuses
 * a*b+c, and bit-exact float parity of the direct path depends on that.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_MAX_DIM 6

/* status codes (mirror src/lib.rs:148-159 Error variants; 4 = a Rust panic) */
enum { ORC_OK = 0, ORC_DATA_SHAPE = 1, ORC_KERNEL_SHAPE = 2, ORC_MISMATCH_SHAPE = 3, ORC_PANIC = 4, ORC_BAD_ARG = 5 };
/* ConvMode codes, src/lib.rs:80-105 */
enum { ORC_MODE_FULL = 0, ORC_MODE_SAME = 1, ORC_MODE_VALID = 2, ORC_MODE_CUSTOM = 3, ORC_MODE_EXPLICIT = 4 };
/* BorderType codes, src/lib.rs:131-143 */
enum { ORC_B_ZEROS = 0, ORC_B_CONST = 1, ORC_B_REFLECT = 2, ORC_B_REPLICATE = 3, ORC_B_CIRCULAR = 4 };
/* element types */
enum { ORC_I32 = 0, ORC_I64 = 1, ORC_F32 = 2, ORC_F64 = 3, ORC_C32 = 4, ORC_C64 = 5,
       ORC_I8 = 6, ORC_I16 = 7, ORC_U8 = 8, ORC_U16 = 9, ORC_U32 = 10, ORC_U64 = 11, ORC_I128 = 12, ORC_U128 = 13 };

int orc_elem_size(int dtype)
{
    switch (dtype) {
    case ORC_I32: case ORC_F32: case ORC_U32: return 4;
    case ORC_I64: case ORC_F64: case ORC_C32: case ORC_U64: return 8;
    case ORC_C64: case ORC_I128: case ORC_U128: return 16;
    case ORC_I8: case ORC_U8: return 1;
    case ORC_I16: case ORC_U16: return 2;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * ConvMode::unfold, src/conv/mod.rs:28-66.  kd[i] = k*d - d + 1 (:35-37).
 * `custom` holds [N] pads for Custom, [N][2] for Explicit.  out_pad is [N][2].
 * ---------------------------------------------------------------------------------- */
int orc_unfold(int mode, int ndim, const int64_t *kshape, const int64_t *dil,
               const int64_t *custom, const int64_t *custom_strides,
               int64_t *out_pad, int64_t *out_strides)
{
    for (int i = 0; i < ndim; i++) {
        int64_t kd = kshape[i] * dil[i] - dil[i] + 1;
        switch (mode) {
        case ORC_MODE_FULL:                       /* :40-43 */
            out_pad[2 * i] = out_pad[2 * i + 1] = kd - 1; out_strides[i] = 1; break;
        case ORC_MODE_SAME:                       /* :44-55 */
            if (kd % 2 == 0) { out_pad[2 * i] = (kd - 1) / 2 + 1; out_pad[2 * i + 1] = (kd - 1) / 2; }
            else { out_pad[2 * i] = out_pad[2 * i + 1] = (kd - 1) / 2; }
            out_strides[i] = 1; break;
        case ORC_MODE_VALID:                      /* :56-59 */
            out_pad[2 * i] = out_pad[2 * i + 1] = 0; out_strides[i] = 1; break;
        case ORC_MODE_CUSTOM:                     /* :60-63 */
            out_pad[2 * i] = out_pad[2 * i + 1] = custom[i]; out_strides[i] = custom_strides[i]; break;
        case ORC_MODE_EXPLICIT:                   /* :64 */
            out_pad[2 * i] = custom[2 * i]; out_pad[2 * i + 1] = custom[2 * i + 1];
            out_strides[i] = custom_strides[i]; break;
        default: return ORC_BAD_ARG;
        }
    }
    return ORC_OK;
}

/* good_size_cc, src/conv_fft/good_size.rs:6-31 (integer division kept as is). */
int64_t orc_good_size_cc(int64_t n)
{
    int64_t best = 1;
    while (best < n) best <<= 1;                 /* next_power_of_two */
    for (;;) {
        int64_t f = best / 4 * 3;
        if (f < n) break;
        if (f == n) return n;
        best = f;
    }
    for (;;) {
        int64_t f = best / 6 * 5;
        if (f < n) break;
        if (f == n) return n;
        best = f;
    }
    return best;
}

/* ------------------------------------------------------------------------------------
 * Padding, sequential restatement.  src/padding/mod.rs:84-153 (drivers :175-452),
 * src/padding/dim.rs:34-157, src/padding/half_dim.rs:30-343.
 * The buffer has shape bshape (>= P on every axis: the FFT path hands in an fft_size
 * buffer and pads inside its [0,P) corner, src/conv_fft/padding.rs:47-59); planes span
 * the [0,P) corner on every other axis exactly as ndarray's index_axis on the slice does.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int ndim; int es;
    int64_t P[ORC_MAX_DIM];        /* padded extent (the slice the reference pads in) */
    int64_t bstr[ORC_MAX_DIM];     /* buffer strides in elements */
    char *buf;
} orc_view;

static void plane_iter_init(const orc_view *v, int axis, int64_t *idx) { (void)v; (void)axis; for (int i = 0; i < ORC_MAX_DIM; i++) idx[i] = 0; }

/* advance idx over all axes except `axis`, row-major; returns 0 when done */
static int plane_iter_next(const orc_view *v, int axis, int64_t *idx)
{
    for (int i = v->ndim - 1; i >= 0; i--) {
        if (i == axis) continue;
        if (++idx[i] < v->P[i]) return 1;
        idx[i] = 0;
    }
    return 0;
}

static int64_t plane_off(const orc_view *v, int axis, const int64_t *idx, int64_t j)
{
    int64_t o = 0;
    for (int i = 0; i < v->ndim; i++) o += (i == axis ? j : idx[i]) * v->bstr[i];
    return o;
}

/* index_axis_mut(dim, j).fill(c): half_dim.rs:40-48, :85-93 */
static int plane_fill(orc_view *v, int axis, int64_t j, const void *c)
{
    if (j < 0 || j >= v->P[axis]) return ORC_PANIC;
    int64_t idx[ORC_MAX_DIM]; plane_iter_init(v, axis, idx);
    do { memcpy(v->buf + plane_off(v, axis, idx, j) * v->es, c, v->es); } while (plane_iter_next(v, axis, idx));
    return ORC_OK;
}

/* index_axis_mut(dim, dst).assign(&index_axis(dim, src)): half_dim.rs:126, :206-208 ... */
static int plane_copy(orc_view *v, int axis, int64_t dst, int64_t src)
{
    if (dst < 0 || dst >= v->P[axis] || src < 0 || src >= v->P[axis]) return ORC_PANIC;  /* ndarray index panic */
    int64_t idx[ORC_MAX_DIM]; plane_iter_init(v, axis, idx);
    do {
        memmove(v->buf + plane_off(v, axis, idx, dst) * v->es, v->buf + plane_off(v, axis, idx, src) * v->es, v->es);
    } while (plane_iter_next(v, axis, idx));
    return ORC_OK;
}

static int half_front(orc_view *v, int axis, int64_t n, int64_t pf, int64_t pb, int border, const void *c)
{
    (void)n;
    int64_t P = v->P[axis];
    int st = ORC_OK;
    for (int64_t j = 0; j < pf && st == ORC_OK; j++) {
        switch (border) {
        case ORC_B_ZEROS: case ORC_B_CONST: st = plane_fill(v, axis, j, c); break;       /* half_dim.rs:30-49 */
        case ORC_B_REPLICATE: st = plane_copy(v, axis, j, pf); break;                    /* :113-129 */
        case ORC_B_REFLECT: st = plane_copy(v, axis, j, (pf - j) + pf); break;           /* :192-211 */
        case ORC_B_CIRCULAR: st = plane_copy(v, axis, j, P - pb - (pf - j)); break;      /* :277-296 */
        default: st = ORC_BAD_ARG;
        }
    }
    return st;
}

static int half_back(orc_view *v, int axis, int64_t n, int64_t pf, int64_t pb, int border, const void *c)
{
    int64_t P = v->P[axis];
    int st = ORC_OK;
    int64_t bi = P - pb - 1;                                                              /* border_index */
    for (int64_t j = n + pf; j < P && st == ORC_OK; j++) {
        switch (border) {
        case ORC_B_ZEROS: case ORC_B_CONST: st = plane_fill(v, axis, j, c); break;       /* :72-94 */
        case ORC_B_REPLICATE: st = plane_copy(v, axis, j, bi); break;                    /* :151-173 */
        case ORC_B_REFLECT:                                                               /* :233-258 */
            if (j - bi > bi) st = ORC_PANIC;      /* usize underflow -> panic */
            else st = plane_copy(v, axis, j, bi - (j - bi));
            break;
        case ORC_B_CIRCULAR: st = plane_copy(v, axis, j, pf + (j - bi - 1)); break;      /* :318-343 */
        default: st = ORC_BAD_ARG;
        }
    }
    return st;
}

/*
 * padding_in (src/padding/mod.rs:119-153) into buf[bshape] whose [0,P) corner receives
 * the padded data.  buf must be pre-initialised by the caller (zeros; or the Const value
 * for the allocating `padding` with PaddingMode::Const, :89-99).
 * borders[ndim][2] / cvals[ndim][2][es]: per-side border (PaddingMode lowered the way
 * the Custom/Explicit drivers :346-452 do; the five plain modes = same border on all sides).
 */
int orc_padding_in(int ndim, int es, const void *data, const int64_t *nshape,
                   const int64_t *pads /*[ndim][2]*/, const int32_t *borders /*[ndim][2]*/,
                   const void *cvals /*[ndim][2] elements*/,
                   void *buf, const int64_t *bshape)
{
    if (ndim < 1 || ndim > ORC_MAX_DIM) return ORC_BAD_ARG;
    orc_view v; v.ndim = ndim; v.es = es; v.buf = (char *)buf;
    int64_t s = 1;
    for (int i = ndim - 1; i >= 0; i--) { v.bstr[i] = s; s *= bshape[i]; }
    for (int i = 0; i < ndim; i++) {
        v.P[i] = nshape[i] + pads[2 * i] + pads[2 * i + 1];
        if (v.P[i] > bshape[i]) return ORC_BAD_ARG;
    }
    /* padding_const: centre slice .assign(input), mod.rs:175-200 */
    {
        int64_t idx[ORC_MAX_DIM] = {0};
        int64_t total = 1; for (int i = 0; i < ndim; i++) total *= nshape[i];
        const char *src = (const char *)data;
        for (int64_t e = 0; e < total; e++) {
            int64_t o = 0;
            for (int i = 0; i < ndim; i++) o += (idx[i] + pads[2 * i]) * v.bstr[i];
            memcpy(v.buf + o * es, src + e * es, es);
            for (int i = ndim - 1; i >= 0; i--) { if (++idx[i] < nshape[i]) break; idx[i] = 0; }
        }
    }
    /* per-dim drivers in order 0..N-1: front then back (dim.rs:34-157; mod.rs:416-451) */
    for (int i = 0; i < ndim; i++) {
        const char *cv = (const char *)cvals + (size_t)(2 * i) * es;
        int st = half_front(&v, i, nshape[i], pads[2 * i], pads[2 * i + 1], borders[2 * i], cv);
        if (st) return st;
        st = half_back(&v, i, nshape[i], pads[2 * i], pads[2 * i + 1], borders[2 * i + 1], cv + es);
        if (st) return st;
    }
    return ORC_OK;
}

/*
 * Closed-form padding on the well-defined domain: an independent
 * second statement used to cross-check the sequential one.  Returns ORC_BAD_ARG outside
 * the domain (Reflect pf,pb <= n-1; Circular pf <= n).
 */
int orc_padding_closed_form(int ndim, int es, const void *data, const int64_t *nshape,
                            const int64_t *pads, const int32_t *borders, const void *cvals, void *out)
{
    int64_t P[ORC_MAX_DIM], nstr[ORC_MAX_DIM];
    int64_t tot = 1, s = 1;
    for (int i = ndim - 1; i >= 0; i--) { nstr[i] = s; s *= nshape[i]; }
    for (int i = 0; i < ndim; i++) {
        P[i] = nshape[i] + pads[2 * i] + pads[2 * i + 1]; tot *= P[i];
        if (borders[2 * i] == ORC_B_REFLECT && pads[2 * i] > nshape[i] - 1) return ORC_BAD_ARG;
        if (borders[2 * i + 1] == ORC_B_REFLECT && pads[2 * i + 1] > nshape[i] - 1) return ORC_BAD_ARG;
        if (borders[2 * i] == ORC_B_CIRCULAR && pads[2 * i] > nshape[i]) return ORC_BAD_ARG;
    }
    int64_t idx[ORC_MAX_DIM] = {0};
    for (int64_t e = 0; e < tot; e++) {
        const char *cst = NULL; int64_t so = 0;
        for (int i = 0; i < ndim; i++) {
            int64_t n = nshape[i], pf = pads[2 * i], c = idx[i], src;
            if (c >= pf && c < pf + n) src = c - pf;
            else {
                int side = c < pf ? 0 : 1;
                int b = borders[2 * i + side];
                int64_t t = c - pf - n;
                if (b == ORC_B_ZEROS || b == ORC_B_CONST) { cst = (const char *)cvals + (size_t)(2 * i + side) * es; src = 0; }
                else if (b == ORC_B_REPLICATE) src = side ? n - 1 : 0;
                else if (b == ORC_B_REFLECT) src = side ? n - 2 - t : pf - c;
                else src = side ? t % n : n + c - pf;
            }
            so += src * nstr[i];
        }
        /* highest-numbered axis with a constant hit wins (later axes overwrite earlier ones) */
        memcpy((char *)out + e * es, cst ? cst : (const char *)data + so * es, es);
        for (int i = ndim - 1; i >= 0; i--) { if (++idx[i] < P[i]) break; idx[i] = 0; }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * gen_offset_list, src/dilation/mod.rs:34-60: row-major over the (optionally reversed)
 * kernel, zero weights dropped, offset = sum idx*dilation*pds_stride.
 * Returns the number of taps; offsets/widx (index into the kernel array) are filled.
 * ---------------------------------------------------------------------------------- */
static int is_zero_elem(int dtype, const void *p)
{
    switch (dtype) {
    case ORC_F32: return *(const float *)p == 0.0f;          /* -0.0 == 0.0 too, as in Rust */
    case ORC_F64: return *(const double *)p == 0.0;
    case ORC_C32: return ((const float *)p)[0] == 0.0f && ((const float *)p)[1] == 0.0f;
    case ORC_C64: return ((const double *)p)[0] == 0.0 && ((const double *)p)[1] == 0.0;
    default: { int es = orc_elem_size(dtype); const char *c = (const char *)p; for (int i = 0; i < es; i++) if (c[i]) return 0; return 1; }
    }
}

int64_t orc_gen_offset_list(int dtype, int ndim, const void *kernel, const int64_t *kshape,
                            const int64_t *dil, int reverse, const int64_t *pds_strides,
                            int64_t *offsets, int64_t *widx)
{
    int es = orc_elem_size(dtype);
    int64_t tot = 1, kstr[ORC_MAX_DIM], s = 1;
    for (int i = ndim - 1; i >= 0; i--) { kstr[i] = s; s *= kshape[i]; }
    for (int i = 0; i < ndim; i++) tot *= kshape[i];
    int64_t idx[ORC_MAX_DIM] = {0}, cnt = 0;
    for (int64_t e = 0; e < tot; e++) {
        int64_t ko = 0, off = 0;
        for (int i = 0; i < ndim; i++) {
            int64_t ki = reverse ? kshape[i] - 1 - idx[i] : idx[i];     /* slice step -1, :35-42 */
            ko += ki * kstr[i];
            off += idx[i] * dil[i] * pds_strides[i];                    /* :44-45, :52-57 */
        }
        if (!is_zero_elem(dtype, (const char *)kernel + ko * es)) {      /* :49 */
            offsets[cnt] = off; widx[cnt] = ko; cnt++;
        }
        for (int i = ndim - 1; i >= 0; i--) { if (++idx[i] < kshape[i]) break; idx[i] = 0; }
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------
 * ConvExt::conv, src/conv/mod.rs:128-200.  Inputs are standard-layout arrays.
 * pads/strides already unfolded (orc_unfold).  init_const: PaddingMode::Const value
 * for the allocating `padding` (:89-99) or NULL -> zero.
 * ---------------------------------------------------------------------------------- */
#define MAC_LOOP(T, ZERO, MACEXPR)                                                       \
    {                                                                                    \
        const T *pd = (const T *)padded; const T *kw = (const T *)kernel; T *po = (T *)out; \
        int64_t oidx[ORC_MAX_DIM] = {0};                                                 \
        for (int64_t e = 0; e < ototal; e++) {                                           \
            int64_t base = 0;                                                            \
            for (int i = 0; i < ndim; i++) base += oidx[i] * strides[i] * pstr[i];       \
            T acc = ZERO;                                                                \
            for (int64_t t = 0; t < ntap; t++) { const T a = pd[base + offs[t]]; const T w = kw[widx[t]]; MACEXPR; } \
            po[e] = acc;                                                                 \
            for (int i = ndim - 1; i >= 0; i--) { if (++oidx[i] < oshape[i]) break; oidx[i] = 0; } \
        }                                                                                \
    }

typedef struct { float re, im; } cf32;
typedef struct { double re, im; } cf64;

int orc_conv_direct(int dtype, int ndim, const void *data, const int64_t *nshape,
                    const void *kernel, const int64_t *kshape, const int64_t *dil, int reverse,
                    const int64_t *pads, const int64_t *strides, const int32_t *borders,
                    const void *cvals, void *out, int64_t *out_shape, int conv_fft_error_quirk)
{
    int es = orc_elem_size(dtype);
    if (!es || ndim < 1 || ndim > ORC_MAX_DIM) return ORC_BAD_ARG;
    int64_t dtot = 1, ktot = 1;
    for (int i = 0; i < ndim; i++) { dtot *= nshape[i]; ktot *= kshape[i]; }
    if (dtot == 0) return ORC_DATA_SHAPE;                                   /* :136-139 */
    if (ktot == 0) return conv_fft_error_quirk ? ORC_DATA_SHAPE : ORC_KERNEL_SHAPE;  /* :141-144; conv_fft/mod.rs:211-213 */
    int64_t kd[ORC_MAX_DIM], P[ORC_MAX_DIM], pstr[ORC_MAX_DIM], oshape[ORC_MAX_DIM];
    int64_t ptot = 1;
    for (int i = 0; i < ndim; i++) {
        kd[i] = kshape[i] * dil[i] - dil[i] + 1;                            /* :146-147 */
        P[i] = nshape[i] + pads[2 * i] + pads[2 * i + 1]; ptot *= P[i];
    }
    /* self.padding(...): :150 (happens BEFORE the mismatch check) */
    char *padded = (char *)calloc((size_t)ptot, es);
    if (!padded) return ORC_BAD_ARG;
    int st = orc_padding_in(ndim, es, data, nshape, pads, borders, cvals, padded, P);
    if (st) { free(padded); return st; }
    for (int i = 0; i < ndim; i++) if (kd[i] > P[i]) { free(padded); return ORC_MISMATCH_SHAPE; }  /* :152-158 */
    { int64_t s = 1; for (int i = ndim - 1; i >= 0; i--) { pstr[i] = s; s *= P[i]; } }
    int64_t *offs = (int64_t *)malloc(sizeof(int64_t) * (size_t)ktot);
    int64_t *widx = (int64_t *)malloc(sizeof(int64_t) * (size_t)ktot);
    int64_t ntap = orc_gen_offset_list(dtype, ndim, kernel, kshape, dil, reverse, pstr, offs, widx);  /* :160 */
    int64_t ototal = 1;
    for (int i = 0; i < ndim; i++) {
        if (strides[i] <= 0) { free(padded); free(offs); free(widx); return ORC_PANIC; }   /* div by zero :165 */
        oshape[i] = (P[i] - kd[i]) / strides[i] + 1; ototal *= oshape[i];   /* :162-167 */
        if (out_shape) out_shape[i] = oshape[i];
    }
    if (out) {
        /* hot loop :188-196; ints wrap (release build) -> unsigned arithmetic */
        switch (dtype) {
        case ORC_I32: case ORC_U32: MAC_LOOP(uint32_t, 0u, acc += a * w) break;
        case ORC_I64: case ORC_U64: MAC_LOOP(uint64_t, 0ull, acc += a * w) break;
        case ORC_I128: case ORC_U128: MAC_LOOP(unsigned __int128, 0, acc += a * w) break;   /* i128 / u128: wrapping, as every integer type */
        case ORC_I8: case ORC_U8: MAC_LOOP(uint8_t, 0, acc = (uint8_t)(acc + (uint8_t)(a * w))) break;
        case ORC_I16: case ORC_U16: MAC_LOOP(uint16_t, 0, acc = (uint16_t)(acc + (uint16_t)(a * w))) break;
        case ORC_F32: MAC_LOOP(float, 0.0f, acc += a * w) break;
        case ORC_F64: MAC_LOOP(double, 0.0, acc += a * w) break;
        case ORC_C32: { const cf32 Z = {0.0f, 0.0f};
            /* num::Complex Mul: (re*re - im*im, re*im + im*re) */
            MAC_LOOP(cf32, Z, { float pr = a.re * w.re - a.im * w.im; float pi = a.re * w.im + a.im * w.re; acc.re += pr; acc.im += pi; }) } break;
        case ORC_C64: { const cf64 Z = {0.0, 0.0};
            MAC_LOOP(cf64, Z, { double pr = a.re * w.re - a.im * w.im; double pi = a.re * w.im + a.im * w.re; acc.re += pr; acc.im += pi; }) } break;
        default: st = ORC_BAD_ARG;
        }
    }
    free(padded); free(offs); free(widx);
    return st;
}

/* ------------------------------------------------------------------------------------
 * FFT pipeline restatement, src/conv_fft/mod.rs:185-292.
 * The butterflies of the reference live in rustfft "6.4" / realfft "3.5" (Cargo.toml:17-18,
 * not under /root/reference, Cargo.lock git-ignored => patch version unpinned).  Their
 * published contract is the unnormalised DFT X[k] = sum x[j] e^{-2 pi i jk/n} (forward),
 * e^{+...} (inverse), realfft's R2C returning bins 0..n/2.  We restate that contract with
 * a plain recursive mixed-radix DFT (any n; O(n p) for a prime factor p), computed in the
 * precision selected by `REAL`.
 * ---------------------------------------------------------------------------------- */
#define REAL double
#define SUF(x) x##_f64
#include "fft_pipeline.inc"
#undef REAL
#undef SUF
#define REAL float
#define SUF(x) x##_f32
#include "fft_pipeline.inc"
#undef REAL
#undef SUF

/*
 * orc_conv_fft: dtype in {F32,F64,C32,C64}.  precision: 0 = native (f32 data -> f32 FFT,
 * the faithful restatement), 1 = force f64 arithmetic (high-precision check of the
 * placement/crop logic; inputs and output stay in `dtype`).
 * use_ref_size: 1 -> fft_size = good_size_cc(max(P,Kd)) as the reference (:229-231).
 */
int orc_conv_fft(int dtype, int ndim, const void *data, const int64_t *nshape,
                 const void *kernel, const int64_t *kshape, const int64_t *dil, int reverse,
                 const int64_t *pads, const int64_t *strides, const int32_t *borders,
                 const void *cvals, void *out, int64_t *out_shape, int precision)
{
    int is_cx = (dtype == ORC_C32 || dtype == ORC_C64);
    int is_dbl = (dtype == ORC_F64 || dtype == ORC_C64);
    if (!(dtype == ORC_F32 || dtype == ORC_F64 || is_cx)) return ORC_BAD_ARG;
    if (is_dbl || precision == 1)
        return orc_conv_fft_impl_f64(is_cx, is_dbl, ndim, data, nshape, kernel, kshape, dil, reverse, pads, strides, borders, cvals, out, out_shape);
    return orc_conv_fft_impl_f32(is_cx, is_dbl, ndim, data, nshape, kernel, kshape, dil, reverse, pads, strides, borders, cvals, out, out_shape);
}
