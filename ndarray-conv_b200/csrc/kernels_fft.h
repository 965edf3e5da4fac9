// kernels_fft.h -- FFT convolution kernels (conv_fft_proc_impl, src/conv_fft/mod.rs:185-292).
//
// Pipeline per overlap-save tile:
//   RowFwd   : border-mapped gather (padded buffer never materialised; replaces conv_fft::padding::data,
//              src/conv_fft/padding.rs:30-62 + padding_in, src/padding/mod.rs:119-153) + R2C/C2C of the last
//              axis in shared memory (replaces real.rs:116-124 / complex.rs:56-82) -> spectra workspace
//   Col      : strided-axis C2C in place, W adjacent columns per block staged through shared memory (no
//              transpose; replaces the permuted_axes + transpose-copy + C2C of real.rs:129-152).  On axis 0 the
//              forward pass, the multiply by the cached kernel spectrum (mod.rs:268) and the inverse pass
//              (real.rs:244-256) are one kernel (mode FWD_MUL_INV).
//   RowInv   : C2R/C2C inverse of the last axis + crop [Kd-1, P) + stride decimation fused into the store
//              (real.rs:263-280 + mod.rs:282-289).  The 1/len scale is folded into the kernel spectrum.
//   Row1D    : N == 1: all of the above in one launch.
//
// Generic bodies: any {2,3,5,7}-smooth length, runtime radix list, ping-pong Stockham in shared memory.
#pragma once
#include "common.h"

namespace ndc {

// exact unsigned division by a run-time constant (n < 2^31): q = (n * mul) >> shift, filled on the host (api.cu)
struct FastDiv {
    uint64_t mul;
    int shift;
    int d;
};
HD int fdiv(int n, const FastDiv &f) { return (int)(((uint64_t)(uint32_t)n * f.mul) >> f.shift); }

template <class R> struct FftPlanDev {
    int L;                       // complex transform length
    int npass;
    int radix[NDC_MAX_PASS];
    FastDiv by_m[NDC_MAX_PASS];  // butterflies per transform of the pass: L / radix
    FastDiv by_ns[NDC_MAX_PASS]; // product of the earlier radices
    const cx<R> *tw;             // tw[j] = exp(-2 pi i j / L), j in [0, L)
};

// ---- small DFTs in registers -------------------------------------------------------------------
template <class R> HD cx<R> mul_neg_i(cx<R> a, bool inverse) { return inverse ? cx<R>{-a.im, a.re} : cx<R>{a.im, -a.re}; }  // a * (-i) forward, a * (+i) inverse

template <class R> HD void dft2(cx<R> *v)
{
    cx<R> a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
}
template <class R> HD void dft4(cx<R> *v, bool inv)
{
    cx<R> a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), a2 = cadd(v[1], v[3]), a3 = mul_neg_i(csub(v[1], v[3]), inv);
    v[0] = cadd(a0, a2); v[2] = csub(a0, a2); v[1] = cadd(a1, a3); v[3] = csub(a1, a3);
}
template <class R> HD void dft8(cx<R> *v, bool inv)
{
    cx<R> e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    dft4(e, inv); dft4(o, inv);
    const R h = (R)0.70710678118654752440;
    // w8^q, q = 1..3 (forward: e^{-2 pi i q/8}; inverse: conjugate)
    cx<R> w1 = cx<R>{h, inv ? h : -h}, w3 = cx<R>{-h, inv ? h : -h};
    o[1] = cmul(o[1], w1); o[2] = mul_neg_i(o[2], inv); o[3] = cmul(o[3], w3);
    for (int q = 0; q < 4; q++) { v[q] = cadd(e[q], o[q]); v[q + 4] = csub(e[q], o[q]); }
}
// cos / sin (2 pi m / 32): the in-register twiddles of the composite radix-16 / radix-32 butterflies
HD constexpr double w32_cos(int m)
{
    constexpr double t[32] = {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476, 0.5555702330196023, 0.38268343236508984, 0.19509032201612833, 6.123233995736766e-17, -0.1950903220161282, -0.3826834323650897, -0.555570233019602, -0.7071067811865475, -0.8314696123025453, -0.9238795325112867, -0.9807852804032304, -1.0, -0.9807852804032304, -0.9238795325112868, -0.8314696123025455, -0.7071067811865477, -0.5555702330196022, -0.38268343236509034, -0.19509032201612866, -1.8369701987210297e-16, 0.1950903220161283, 0.38268343236509, 0.5555702330196018, 0.7071067811865474, 0.8314696123025452, 0.9238795325112865, 0.9807852804032303};
    return t[m & 31];
}
HD constexpr double w32_sin(int m)
{
    constexpr double t[32] = {0.0, 0.19509032201612825, 0.3826834323650898, 0.5555702330196022, 0.7071067811865475, 0.8314696123025452, 0.9238795325112867, 0.9807852804032304, 1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025455, 0.7071067811865476, 0.5555702330196022, 0.3826834323650899, 0.1950903220161286, 1.2246467991473532e-16, -0.19509032201612836, -0.38268343236508967, -0.555570233019602, -0.7071067811865475, -0.8314696123025452, -0.9238795325112865, -0.9807852804032303, -1.0, -0.9807852804032304, -0.9238795325112866, -0.8314696123025455, -0.7071067811865477, -0.5555702330196022, -0.3826834323650904, -0.19509032201612872};
    return t[m & 31];
}
// Cooley-Tukey N = N1*N2 in registers: n = N2*n1 + n2, k = k1 + N1*k2.  Natural order in, natural order out.
template <class R, int N1, int N2, class DFT1, class DFT2> HD void dft_ct(cx<R> *v, bool inv, DFT1 d1, DFT2 d2)
{
    constexpr int N = N1 * N2;
    cx<R> a[N2][N1];
#pragma unroll
    for (int n2 = 0; n2 < N2; n2++) {
#pragma unroll
        for (int n1 = 0; n1 < N1; n1++) a[n2][n1] = v[N2 * n1 + n2];
        d1(a[n2], inv);                                   // a[n2][k1]
    }
#pragma unroll
    for (int k1 = 0; k1 < N1; k1++) {
        cx<R> b[N2];
#pragma unroll
        for (int n2 = 0; n2 < N2; n2++) {
            const int m = (n2 * k1 * (32 / N)) & 31;      // W_N^{n2 k1} = W_32^{n2 k1 32/N}
            if (m == 0) b[n2] = a[n2][k1];
            else {
                const cx<R> w = cx<R>{(R)w32_cos(m), (R)(inv ? w32_sin(m) : -w32_sin(m))};
                b[n2] = cmul(a[n2][k1], w);
            }
        }
        d2(b, inv);                                       // b[k2]
#pragma unroll
        for (int k2 = 0; k2 < N2; k2++) v[k1 + N1 * k2] = b[k2];
    }
}
template <class R> struct Dft4Fn { HD void operator()(cx<R> *v, bool inv) const { dft4(v, inv); } };
template <class R> struct Dft8Fn { HD void operator()(cx<R> *v, bool inv) const { dft8(v, inv); } };
template <class R> HD void dft16(cx<R> *v, bool inv) { dft_ct<R, 4, 4>(v, inv, Dft4Fn<R>(), Dft4Fn<R>()); }
template <class R> HD void dft32(cx<R> *v, bool inv) { dft_ct<R, 4, 8>(v, inv, Dft4Fn<R>(), Dft8Fn<R>()); }

// odd prime radices: O(R^2) with exact constants
template <class R, int RDX> HD void dft_odd(cx<R> *v, bool inv)
{
    // cos/sin(2 pi j / RDX), j = 0..RDX-1
    R cs[RDX], sn[RDX];
    if (RDX == 3) {
        const double c[3] = {1.0, -0.5, -0.5}, s[3] = {0.0, 0.86602540378443864676, -0.86602540378443864676};
        for (int j = 0; j < RDX; j++) { cs[j] = (R)c[j % 3]; sn[j] = (R)s[j % 3]; }
    } else if (RDX == 5) {
        const double c[5] = {1.0, 0.30901699437494742410, -0.80901699437494742410, -0.80901699437494742410, 0.30901699437494742410};
        const double s[5] = {0.0, 0.95105651629515357212, 0.58778525229247312917, -0.58778525229247312917, -0.95105651629515357212};
        for (int j = 0; j < RDX; j++) { cs[j] = (R)c[j % 5]; sn[j] = (R)s[j % 5]; }
    } else {
        const double c[7] = {1.0, 0.62348980185873353053, -0.22252093395631440429, -0.90096886790241912624,
                             -0.90096886790241912624, -0.22252093395631440429, 0.62348980185873353053};
        const double s[7] = {0.0, 0.78183148246802980871, 0.97492791218182360702, 0.43388373911755812048,
                             -0.43388373911755812048, -0.97492791218182360702, -0.78183148246802980871};
        for (int j = 0; j < RDX; j++) { cs[j] = (R)c[j % 7]; sn[j] = (R)s[j % 7]; }
    }
    cx<R> o[RDX];
    for (int q = 0; q < RDX; q++) {
        cx<R> acc = v[0];
        for (int t = 1; t < RDX; t++) {
            int j = (t * q) % RDX;
            cx<R> w = cx<R>{cs[j], inv ? sn[j] : -sn[j]};   // forward e^{-i theta}
            acc = cadd(acc, cmul(v[t], w));
        }
        o[q] = acc;
    }
    for (int q = 0; q < RDX; q++) v[q] = o[q];
}
template <class R, int RDX> HD void dft(cx<R> *v, bool inv)
{
    if constexpr (RDX == 2) dft2(v);
    else if constexpr (RDX == 4) dft4(v, inv);
    else if constexpr (RDX == 8) dft8(v, inv);
    else if constexpr (RDX == 16) dft16(v, inv);
    else if constexpr (RDX == 32) dft32(v, inv);
    else dft_odd<R, RDX>(v, inv);
}

// ---- one Stockham pass over `nbatch` transforms in shared memory ----------------------------------
// element (b, j) lives at b*bstride + j*estride.  batch_fast: consecutive threads take consecutive b.
template <class R, int RDX>
HD void stockham_pass_r(const BlockCtx &c, const cx<R> *in, cx<R> *out, int L, int Ns, const cx<R> *tw, bool inv,
                        int nbatch, int estride, int bstride, bool batch_fast, const FastDiv &by_m, const FastDiv &by_ns)
{
    const int m = L / RDX;
    const int total = nbatch * m;
    const int step = L / (Ns * RDX);
#if defined(__CUDA_ARCH__)
    const int bshift = 31 - __clz(nbatch > 0 ? nbatch : 1);                           // batch_fast callers pass a power of two
#else
    const int bshift = 31 - __builtin_clz((unsigned)(nbatch > 0 ? nbatch : 1));
#endif
    for (int idx = c.tid; idx < total; idx += c.nt) {
        int b, j;
        if (batch_fast) { b = idx & (nbatch - 1); j = idx >> bshift; } else { b = fdiv(idx, by_m); j = idx - b * m; }
        const int jq = fdiv(j, by_ns);
        const int k = j - jq * Ns;
        cx<R> v[RDX];
        const cx<R> *src = in + (int64_t)b * bstride;
        for (int t = 0; t < RDX; t++) v[t] = src[(int64_t)(j + t * m) * estride];
        if (Ns > 1) {
            for (int t = 1; t < RDX; t++) {
                cx<R> w = tw[t * k * step];
                v[t] = inv ? cmulc(v[t], w) : cmul(v[t], w);
            }
        }
        dft<R, RDX>(v, inv);
        cx<R> *dst = out + (int64_t)b * bstride;
        const int j0 = jq * Ns * RDX + k;
        for (int t = 0; t < RDX; t++) dst[(int64_t)(j0 + t * Ns) * estride] = v[t];
    }
}

// Full transform: data starts in `a`, ping-pongs with `b`; returns the buffer holding the result.
template <class R>
HD cx<R> *fft_smem(const BlockCtx &c, cx<R> *a, cx<R> *b, const FftPlanDev<R> &pl, bool inv,
                   int nbatch, int estride, int bstride, bool batch_fast, const cx<R> *tw_override = nullptr)
{
    const cx<R> *tw = tw_override ? tw_override : pl.tw;       // shared-memory copy of the twiddle table when the kernel made one
    int Ns = 1;
    cx<R> *in = a, *out = b;
    for (int p = 0; p < pl.npass; p++) {
        const int r = pl.radix[p];
        switch (r) {
        case 2: stockham_pass_r<R, 2>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        case 3: stockham_pass_r<R, 3>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        case 4: stockham_pass_r<R, 4>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        case 5: stockham_pass_r<R, 5>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        case 7: stockham_pass_r<R, 7>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        case 16: stockham_pass_r<R, 16>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        case 32: if constexpr (sizeof(R) == 4) stockham_pass_r<R, 32>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;   // f64 plans stop at radix 16 (factor_radices)
        default: stockham_pass_r<R, 8>(c, in, out, pl.L, Ns, tw, inv, nbatch, estride, bstride, batch_fast, pl.by_m[p], pl.by_ns[p]); break;
        }
        Ns *= r;
        c.sync();
        cx<R> *t = in; in = out; out = t;
    }
    return in;
}

// ---- parameters ------------------------------------------------------------------------------------
template <class R> struct RowParams {
    int ndim, is_cx;
    int64_t n[NDC_MAX_DIM], xstr[NDC_MAX_DIM], P[NDC_MAX_DIM];
    const int32_t *map[NDC_MAX_DIM];
    unsigned char cfront[NDC_MAX_DIM][16], cback[NDC_MAX_DIM][16];   // R (real input) or cx<R>
    int F[NDC_MAX_DIM], V[NDC_MAX_DIM], ntiles[NDC_MAX_DIM], Kd[NDC_MAX_DIM];
    int64_t s[NDC_MAX_DIM], O[NDC_MAX_DIM];
    const void *x;
    void *out;
    cx<R> *ws;                  // spectra workspace [tile][r0]..[r_{N-2}][Hp]
    const cx<R> *kspec;         // kernel spectrum  [r0]..[r_{N-2}][Hp] (Row1D only)
    int H, Hp;                  // valid bins per row / row pitch (complex elements)
    int B;                      // rows per block iteration
    FftPlanDev<R> plan;         // last-axis complex transform (L = F/2 for real input, F for complex)
    const cx<R> *twr;           // twr[k] = exp(-2 pi i k / F), k in [0, L/2]   (real input only)
    int64_t rows_per_tile;      // prod_{a<N-1} F[a]
    int64_t tile_elems;         // rows_per_tile * Hp
    int64_t nwork;
    R scale;                    // applied by the final store; 0 = none (conv_fft folds 1/len into the kernel spectrum)
    int tw_smem_off;            // > 0: byte offset of a shared-memory copy of the twiddle table (loaded once per CTA)
};

template <class R> struct ColParams {
    cx<R> *ws;
    const cx<R> *kspec;
    int F, W, mode;             // mode: 0 forward, 1 inverse, 2 forward * kspec -> inverse
    int64_t inner, outer, tile_elems, ntiles_total;
    FftPlanDev<R> plan;
    int64_t nwork;              // ntiles_total * outer * (inner / W)
    int tw_smem_off;            // > 0: byte offset of a shared-memory copy of the twiddle table
};

// copy the twiddle table into shared memory once per CTA (saves one dependent global load per butterfly input)
template <class R> HD const cx<R> *stage_twiddles(const BlockCtx &c, const FftPlanDev<R> &pl, int off)
{
    if (off <= 0) return nullptr;
    cx<R> *s = (cx<R> *)(c.smem + off);
    for (int i = c.tid; i < pl.L; i += c.nt) s[i] = pl.tw[i];
    c.sync();
    return s;
}

// ---- row helpers -----------------------------------------------------------------------------------
// Resolution of the axes below the last one for one row of one tile (see padded_at in kernels_direct.h).
template <class R> struct RowSrc {
    bool zero;          // beyond the padded extent, or never-written cell: the whole row reads 0 unless the last axis is constant
    bool beyond;        // beyond the padded extent on some lower axis: zero even where the last axis is constant
    bool has_const;     // a lower axis is in a constant border
    const unsigned char *cval;
    int64_t base;       // element offset of the row in x
};

template <class R> HD RowSrc<R> resolve_row(const RowParams<R> &p, const int64_t *coord /* padded coords of axes < N-1 */)
{
    RowSrc<R> r; r.zero = false; r.beyond = false; r.has_const = false; r.cval = nullptr; r.base = 0;
    for (int a = p.ndim - 2; a >= 0; a--) {
        if (coord[a] >= p.P[a]) { r.beyond = true; r.zero = true; return r; }
    }
    for (int a = p.ndim - 2; a >= 0; a--) {
        int32_t m = p.map[a][coord[a]];
        if (m >= 0) r.base += (int64_t)m * p.xstr[a];
        else if (m == NDC_MAP_INIT) r.zero = true;
        else { r.has_const = true; r.cval = (m == NDC_MAP_CONST_FRONT) ? p.cfront[a] : p.cback[a]; return r; }
    }
    return r;
}

// value of the padded, zero-extended signal at last-axis padded coordinate cl for a resolved row
template <class R, class E> HD E gather_elem(const RowParams<R> &p, const RowSrc<R> &rs, int64_t cl, E zero)
{
    const int a = p.ndim - 1;
    if (rs.beyond || cl >= p.P[a]) return zero;
    int32_t m = p.map[a][cl];
    if (m == NDC_MAP_CONST_FRONT) return *(const E *)p.cfront[a];
    if (m == NDC_MAP_CONST_BACK) return *(const E *)p.cback[a];
    if (rs.has_const) return *(const E *)rs.cval;
    if (m == NDC_MAP_INIT || rs.zero) return zero;
    return ((const E *)p.x)[rs.base + (int64_t)m * p.xstr[a]];
}

// R2C post-processing of a half-length complex transform: z[0..L) -> X[0..L]; dst may be global or shared.
template <class R> HD void r2c_post(const BlockCtx &c, const cx<R> *z, cx<R> *dst, int L, const cx<R> *twr, int nbatch, int zstride, int64_t dstride)
{
    const int half = L / 2 + 1;
    for (int idx = c.tid; idx < nbatch * half; idx += c.nt) {
        const int k = idx % half, b = idx / half;
        const cx<R> *zz = z + (int64_t)b * zstride;
        cx<R> *d = dst + (int64_t)b * dstride;
        const int k2 = L - k;
        cx<R> zk = zz[k], zk2 = (k == 0) ? zz[0] : zz[k2];
        cx<R> e = cx<R>{(R)0.5 * (zk.re + zk2.re), (R)0.5 * (zk.im - zk2.im)};        // (zk + conj(zk2)) / 2
        cx<R> o = cx<R>{(R)0.5 * (zk.im + zk2.im), (R)-0.5 * (zk.re - zk2.re)};       // (zk - conj(zk2)) / (2i)
        cx<R> wo = cmul(twr[k], o);
        d[k] = cadd(e, wo);
        if (k2 != k) d[k2] = cconj(csub(e, wo));
    }
}

// C2R pre-processing: X[0..L] -> z[0..L) whose inverse half-length transform interleaves the real output (x F).
template <class R> HD void c2r_pre(const BlockCtx &c, const cx<R> *X, cx<R> *z, int L, const cx<R> *twr, int nbatch, int xstride, int zstride)
{
    const int half = L / 2 + 1;
    for (int idx = c.tid; idx < nbatch * half; idx += c.nt) {
        const int k = idx % half, b = idx / half;
        const cx<R> *xx = X + (int64_t)b * xstride;
        cx<R> *zz = z + (int64_t)b * zstride;
        const int k2 = L - k;
        cx<R> xk = xx[k], xk2 = xx[k2];
        if (k == 0) { xk.im = 0; xk2.im = 0; }                 // realfft ignores the imaginary part of DC / Nyquist
        cx<R> e = cx<R>{xk.re + xk2.re, xk.im - xk2.im};       // xk + conj(xk2)
        cx<R> dd = cx<R>{xk.re - xk2.re, xk.im + xk2.im};      // xk - conj(xk2)
        cx<R> o = cmulc(dd, twr[k]);                           // conj(w^k) * dd
        zz[k] = cx<R>{e.re - o.im, e.im + o.re};               // e + i o
        if (k != 0 && k2 != k) zz[k2] = cx<R>{e.re + o.im, -e.im + o.re};   // conj(e) + i conj(o)
    }
}

// decode helpers
HD void decode_rowmajor(int64_t v, const int *dims, int nd, int64_t *out)
{
    for (int a = nd - 1; a >= 0; a--) { out[a] = v % dims[a]; v /= dims[a]; }
}

// ---- RowFwd ------------------------------------------------------------------------------------------
// work item = B consecutive rows of one tile.  Shared memory: 2 buffers of B * (L + 1) complex.
template <class R> struct RowFwdBody {
    static HD void run(const BlockCtx &c, const RowParams<R> &p)
    {
        const int N = p.ndim, L = p.plan.L, B = p.B, zs = L + 1;
        const int F = p.F[N - 1];
        cx<R> *bufA = (cx<R> *)c.smem, *bufB = bufA + (size_t)B * zs;
        RowSrc<R> *rsrc = (RowSrc<R> *)(bufB + (size_t)B * zs);
        const cx<R> *stw = stage_twiddles(c, p.plan, p.tw_smem_off);
        const int64_t groups_per_tile = (p.rows_per_tile + B - 1) / B;
        for (int64_t w = c.bid; w < p.nwork; w += c.nb) {
            const int64_t tile = w / groups_per_tile, row0 = (w % groups_per_tile) * B;
            const int nrows = (int)((p.rows_per_tile - row0) < B ? (p.rows_per_tile - row0) : B);
            int64_t t[NDC_MAX_DIM];
            decode_rowmajor(tile, p.ntiles, N, t);
            const int64_t cl0 = t[N - 1] * p.V[N - 1];
            // resolve the lower axes once per row
            for (int b = c.tid; b < nrows; b += c.nt) {
                int64_t r[NDC_MAX_DIM], coord[NDC_MAX_DIM];
                decode_rowmajor(row0 + b, p.F, N - 1, r);
                for (int a = 0; a < N - 1; a++) coord[a] = t[a] * p.V[a] + r[a];
                rsrc[b] = resolve_row(p, coord);
            }
            c.sync();
            // gather
            const int per_row = F;   // scalars (real) or complex elements per row
            for (int idx = c.tid; idx < nrows * per_row; idx += c.nt) {
                const int i = idx % per_row, b = idx / per_row;
                const RowSrc<R> rs = rsrc[b];
                if (p.is_cx) bufA[(size_t)b * zs + i] = gather_elem<R, cx<R>>(p, rs, cl0 + i, cx<R>{(R)0, (R)0});
                else ((R *)(bufA + (size_t)b * zs))[i] = gather_elem<R, R>(p, rs, cl0 + i, (R)0);
            }
            c.sync();
            cx<R> *res = fft_smem(c, bufA, bufB, p.plan, false, nrows, 1, zs, false, stw);
            cx<R> *dst = p.ws + tile * p.tile_elems + row0 * p.Hp;
            if (p.is_cx) {
                for (int idx = c.tid; idx < nrows * p.Hp; idx += c.nt) {
                    const int k = idx % p.Hp, b = idx / p.Hp;
                    dst[(int64_t)b * p.Hp + k] = k < L ? res[(size_t)b * zs + k] : cx<R>{(R)0, (R)0};
                }
            } else {
                r2c_post(c, res, dst, L, p.twr, nrows, zs, p.Hp);
                const int padc = p.Hp - p.H;
                for (int idx = c.tid; idx < nrows * padc; idx += c.nt) dst[(int64_t)(idx / padc) * p.Hp + p.H + idx % padc] = cx<R>{(R)0, (R)0};
            }
            c.sync();
        }
    }
};

// crop [Kd-1, F) of the tile, global position m = tile*V + i, keep m < P and (m-Kd+1) % s == 0
template <class R> HD void crop_store_row(const BlockCtx &c, const RowParams<R> &p, const cx<R> *res, int zs, int nrows,
                                          const int64_t *orow_base /* [B] output element offset of each row, or -1 */, int64_t tl)
{
    const int N = p.ndim, a = N - 1, F = p.F[a], Kd = p.Kd[a];
    const int64_t s = p.s[a];
    const int nvalid = F - Kd + 1;
    for (int idx = c.tid; idx < nrows * nvalid; idx += c.nt) {
        const int i = Kd - 1 + idx % nvalid, b = idx / nvalid;
        if (orow_base[b] < 0) continue;
        const int64_t m = tl * p.V[a] + i;
        if (m >= p.P[a]) continue;
        const int64_t q = m - (Kd - 1);
        if (q % s) continue;
        const int64_t o = q / s;
        if (o >= p.O[a]) continue;
        const R sc = p.scale != (R)0 ? p.scale : (R)1;
        if (p.is_cx) { const cx<R> v = res[(size_t)b * zs + i]; ((cx<R> *)p.out)[orow_base[b] + o] = cx<R>{v.re * sc, v.im * sc}; }
        else ((R *)p.out)[orow_base[b] + o] = ((const R *)(res + (size_t)b * zs))[i] * sc;
    }
}

// ---- RowInv ------------------------------------------------------------------------------------------
// work item = (B consecutive output rows, last-axis tile).  Output rows are enumerated so that no block is
// launched for rows the crop / stride would discard.
template <class R> struct RowInvBody {
    static HD void run(const BlockCtx &c, const RowParams<R> &p)
    {
        const int N = p.ndim, L = p.plan.L, B = p.B, zs = L + 1;
        cx<R> *bufA = (cx<R> *)c.smem, *bufB = bufA + (size_t)B * zs;
        int64_t *orow_base = (int64_t *)(bufB + (size_t)B * zs);
        const cx<R> *stw = stage_twiddles(c, p.plan, p.tw_smem_off);
        int64_t out_rows = 1;
        for (int a = 0; a < N - 1; a++) out_rows *= p.O[a];
        const int64_t groups = (out_rows + B - 1) / B;
        const int ntl = p.ntiles[N - 1];
        for (int64_t w = c.bid; w < p.nwork; w += c.nb) {
            const int64_t tl = w % ntl, row0 = (w / ntl) * B;
            const int nrows = (int)((out_rows - row0) < B ? (out_rows - row0) : B);
            (void)groups;
            // load H bins per row (and record where each row goes)
            for (int idx = c.tid; idx < nrows * p.H; idx += c.nt) {
                const int k = idx % p.H, b = idx / p.H;
                int64_t o[NDC_MAX_DIM];
                int odims[NDC_MAX_DIM];
                for (int a = 0; a < N - 1; a++) odims[a] = (int)p.O[a];
                decode_rowmajor(row0 + b, odims, N - 1, o);
                int64_t tile = 0, row = 0, obase = 0;
                for (int a = 0; a < N - 1; a++) {
                    const int64_t q = o[a] * p.s[a];            // m - (Kd-1)
                    const int64_t ta = q / p.V[a];
                    const int64_t ra = q - ta * p.V[a] + p.Kd[a] - 1;
                    tile = tile * p.ntiles[a] + ta;
                    row = row * p.F[a] + ra;
                    obase = obase * p.O[a] + o[a];
                }
                tile = tile * ntl + tl;
                bufA[(size_t)b * zs + k] = p.ws[tile * p.tile_elems + row * p.Hp + k];
                if (k == 0) orow_base[b] = obase * p.O[N - 1];
            }
            c.sync();
            cx<R> *res;
            if (p.is_cx) res = fft_smem(c, bufA, bufB, p.plan, true, nrows, 1, zs, false, stw);
            else {
                c2r_pre(c, bufA, bufB, L, p.twr, nrows, zs, zs);
                c.sync();
                res = fft_smem(c, bufB, bufA, p.plan, true, nrows, 1, zs, false, stw);
            }
            crop_store_row(c, p, res, zs, nrows, orow_base, tl);
            c.sync();
        }
    }
};

// ---- Row1D (N == 1): gather -> FFT -> x kernel spectrum -> inverse FFT -> crop, one launch --------------
template <class R> struct Row1DBody {
    static HD void run(const BlockCtx &c, const RowParams<R> &p)
    {
        const int L = p.plan.L, zs = L + 1, F = p.F[0];
        cx<R> *bufA = (cx<R> *)c.smem, *bufB = bufA + zs;
        int64_t *orow_base = (int64_t *)(bufB + zs);
        const cx<R> *stw = stage_twiddles(c, p.plan, p.tw_smem_off);
        for (int64_t w = c.bid; w < p.nwork; w += c.nb) {
            const int64_t cl0 = w * p.V[0];
            RowSrc<R> rs; rs.zero = false; rs.beyond = false; rs.has_const = false; rs.cval = nullptr; rs.base = 0;
            for (int i = c.tid; i < F; i += c.nt) {
                if (p.is_cx) bufA[i] = gather_elem<R, cx<R>>(p, rs, cl0 + i, cx<R>{(R)0, (R)0});
                else ((R *)bufA)[i] = gather_elem<R, R>(p, rs, cl0 + i, (R)0);
            }
            if (c.tid == 0) orow_base[0] = 0;
            c.sync();
            cx<R> *res = fft_smem(c, bufA, bufB, p.plan, false, 1, 1, zs, false, stw);
            cx<R> *oth = (res == bufA) ? bufB : bufA;
            if (p.is_cx) {
                for (int k = c.tid; k < L; k += c.nt) res[k] = cmul(res[k], p.kspec[k]);
                c.sync();
                res = fft_smem(c, res, oth, p.plan, true, 1, 1, zs, false, stw);
            } else {
                r2c_post(c, res, oth, L, p.twr, 1, zs, zs);
                c.sync();
                for (int k = c.tid; k <= L; k += c.nt) oth[k] = cmul(oth[k], p.kspec[k]);
                c.sync();
                c2r_pre(c, oth, res, L, p.twr, 1, zs, zs);
                c.sync();
                res = fft_smem(c, res, oth, p.plan, true, 1, 1, zs, false, stw);
            }
            crop_store_row(c, p, res, zs, 1, orow_base, w);
            c.sync();
        }
    }
};

// ---- Col ---------------------------------------------------------------------------------------------
// work item = (tile, outer index, block of W adjacent inner columns).  Shared memory: 2 x F x W complex.
template <class R> struct ColBody {
    static HD void run(const BlockCtx &c, const ColParams<R> &p)
    {
        const int F = p.F, W = p.W;
        cx<R> *bufA = (cx<R> *)c.smem, *bufB = bufA + (size_t)F * W;
        const cx<R> *stw = stage_twiddles(c, p.plan, p.tw_smem_off);
        const int64_t iblocks = p.inner / W;
        for (int64_t w = c.bid; w < p.nwork; w += c.nb) {
            const int64_t ib = w % iblocks, o = (w / iblocks) % p.outer, tile = w / (iblocks * p.outer);
            const int64_t rel = o * F * p.inner + ib * W;          // offset inside a tile (same for kspec)
            cx<R> *g = p.ws + tile * p.tile_elems + rel;
            for (int idx = c.tid; idx < F * W; idx += c.nt) bufA[idx] = g[(int64_t)(idx / W) * p.inner + idx % W];
            c.sync();
            cx<R> *res = fft_smem(c, bufA, bufB, p.plan, p.mode == 1, W, W, 1, true, stw);
            if (p.mode == 2) {
                const cx<R> *ks = p.kspec + rel;
                for (int idx = c.tid; idx < F * W; idx += c.nt) res[idx] = cmul(res[idx], ks[(int64_t)(idx / W) * p.inner + idx % W]);
                c.sync();
                res = fft_smem(c, res, res == bufA ? bufB : bufA, p.plan, true, W, W, 1, true, stw);
            }
            for (int idx = c.tid; idx < F * W; idx += c.nt) g[(int64_t)(idx / W) * p.inner + idx % W] = res[idx];
            c.sync();
        }
    }
};

// ---- global-memory Stockham pass (public processors only: axes outside the shared-memory envelope) ---------------
// rustfft takes any length (real.rs:40,62; complex.rs:56); the shared-memory bodies above take {2,3,5,7}-smooth lengths
// of at most one tile.  Everything else -- longer axes, odd real axes, prime factors above 7 -- runs as out-of-place
// autosort Stockham passes through HBM, one launch per radix: element (o, i, c) of the transformed axis lives at
// (o * n + i) * inner + c.  Radices 2..16 are register butterflies (one thread per butterfly); any other prime factor r
// is evaluated as an r-term sum per output (one thread per output, O(n * r) for the pass) with the twiddle and the
// radix-r root taken from the one table exp(-2 pi i j / n).
template <class R> struct GPassParams {
    const cx<R> *in;
    cx<R> *out;
    int64_t n, m, Ns;           // axis length, n / radix, product of the earlier radices
    int radix;
    int64_t inner, outer;
    const cx<R> *tw;            // tw[j] = exp(-2 pi i j / n), j in [0, n)
    int inv;
    int64_t nwork;              // register radix: outer * m * inner butterflies; generic: outer * n * inner outputs
};

template <class R, int RDX> HD void gpass_butterfly(const GPassParams<R> &p, int64_t o, int64_t j, int64_t cc)
{
    const int64_t k = j % p.Ns, jq = j / p.Ns, step = p.n / (p.Ns * RDX);
    const cx<R> *src = p.in + (o * p.n) * p.inner + cc;
    cx<R> v[RDX];
    for (int t = 0; t < RDX; t++) v[t] = src[(j + t * p.m) * p.inner];
    if (p.Ns > 1) {
        for (int t = 1; t < RDX; t++) {
            const cx<R> w = p.tw[t * k * step];
            v[t] = p.inv ? cmulc(v[t], w) : cmul(v[t], w);
        }
    }
    dft<R, RDX>(v, p.inv != 0);
    cx<R> *dst = p.out + (o * p.n) * p.inner + cc;
    const int64_t j0 = jq * p.Ns * RDX + k;
    for (int t = 0; t < RDX; t++) dst[(j0 + t * p.Ns) * p.inner] = v[t];
}

template <class R> struct GPassBody {
    static HD void run(const BlockCtx &c, const GPassParams<R> &p)
    {
        const int r = p.radix;
        const bool reg = r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16;
        for (int64_t e = c.bid * c.nt + c.tid; e < p.nwork; e += c.nb * c.nt) {
            const int64_t cc = e % p.inner, rest = e / p.inner;
            if (reg) {
                const int64_t j = rest % p.m, o = rest / p.m;
                switch (r) {
                case 2: gpass_butterfly<R, 2>(p, o, j, cc); break;
                case 3: gpass_butterfly<R, 3>(p, o, j, cc); break;
                case 4: gpass_butterfly<R, 4>(p, o, j, cc); break;
                case 5: gpass_butterfly<R, 5>(p, o, j, cc); break;
                case 7: gpass_butterfly<R, 7>(p, o, j, cc); break;
                case 8: gpass_butterfly<R, 8>(p, o, j, cc); break;
                default: gpass_butterfly<R, 16>(p, o, j, cc); break;
                }
            } else {
                // output (j, q): X = sum_t in[j + t m] * W_n^{t (k step + m q)}; the exponent advances by d modulo n
                const int64_t i = rest % p.n, o = rest / p.n;
                const int64_t q = i / p.m, j = i % p.m;
                const int64_t k = j % p.Ns, jq = j / p.Ns, step = p.n / (p.Ns * r);
                const int64_t d = (k * step + p.m * q) % p.n;
                const cx<R> *src = p.in + (o * p.n + j) * p.inner + cc;
                double are = 0.0, aim = 0.0;                   // r can be a long prime: sum in double for either element type
                int64_t ex = 0;
                for (int t = 0; t < r; t++) {
                    const cx<R> x = src[(int64_t)t * p.m * p.inner];
                    const cx<R> w = p.tw[ex];
                    const double wr = (double)w.re, wi = p.inv ? -(double)w.im : (double)w.im;
                    are += (double)x.re * wr - (double)x.im * wi;
                    aim += (double)x.re * wi + (double)x.im * wr;
                    ex += d; if (ex >= p.n) ex -= p.n;
                }
                p.out[(o * p.n + jq * p.Ns * r + k + q * p.Ns) * p.inner + cc] = cx<R>{(R)are, (R)aim};
            }
        }
    }
};

// last-axis re-packing around the global passes: rows of `spitch` source elements -> rows of `dpitch` destination elements
template <class R> struct GMoveParams {
    const void *src;
    void *dst;
    int64_t rows, n, spitch, dpitch, H;
    int mode;   // 0 real -> complex; 1 complex copy of H bins, zero up to dpitch; 2 Hermitian half (H = n/2+1 bins) -> n bins;
                // 3 real part * scale -> real; 4 complex * scale -> complex
    R scale;
};
template <class R> struct GMoveBody {
    static HD void run(const BlockCtx &c, const GMoveParams<R> &p)
    {
        const int64_t total = p.rows * p.dpitch;
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int64_t k = e % p.dpitch, row = e / p.dpitch;
            const cx<R> zero = cx<R>{(R)0, (R)0};
            if (p.mode == 0) ((cx<R> *)p.dst)[e] = cx<R>{((const R *)p.src)[row * p.spitch + k], (R)0};
            else if (p.mode == 1) ((cx<R> *)p.dst)[e] = k < p.H ? ((const cx<R> *)p.src)[row * p.spitch + k] : zero;
            else if (p.mode == 2) {
                const cx<R> *s = (const cx<R> *)p.src + row * p.spitch;
                ((cx<R> *)p.dst)[e] = k < p.H ? s[k] : cconj(s[p.n - k]);
            } else if (p.mode == 3) ((R *)p.dst)[e] = ((const cx<R> *)p.src)[row * p.spitch + k].re * p.scale;
            else { const cx<R> v = ((const cx<R> *)p.src)[row * p.spitch + k]; ((cx<R> *)p.dst)[e] = cx<R>{v.re * p.scale, v.im * p.scale}; }
        }
    }
};

// ---- spectrum layout of the public processors -------------------------------------------------------
// natural [n0][rest (pitch-padded last axis)]  <->  rotated [rest (dense)][n0]: "axis 0 moves to the end".
template <class R> struct PermuteParams {
    const cx<R> *src;
    cx<R> *dst;
    int64_t n0, rest_rows, H, Hp;     // rest = rest_rows x H valid bins, rows pitched by Hp on the natural side
    int to_rotated;                   // 1: natural -> rotated, 0: rotated -> natural
};
template <class R> struct PermuteBody {
    static HD void run(const BlockCtx &c, const PermuteParams<R> &p)
    {
        const int64_t total = p.n0 * p.rest_rows * p.H;
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int64_t i0 = e % p.n0, rest = e / p.n0;          // rotated side is dense: (rest, i0), i0 fastest
            const int64_t k = rest % p.H, rr = rest / p.H;
            const int64_t nat = (i0 * p.rest_rows + rr) * p.Hp + k;
            if (p.to_rotated) p.dst[e] = p.src[nat];
            else p.dst[nat] = p.src[e];
        }
    }
};


// ---- Bluestein (chirp-z) for axes with a large prime factor: a length-n DFT as a length-M circular convolution, M >= 2n - 1 a power of two
//   X[k] = w[k] * sum_j (x[j] w[j]) conj(w[k - j]),   w[j] = exp(-+ i pi j^2 / n)
// mode 0: dst[o][j][i] = src[o][j][i] * w[j] for j < n, 0 for n <= j < M      ([outer][n][inner] -> [outer][M][inner])
// mode 1: dst[o][m][i] *= bspec[m]                                           (spectrum of the chirp kernel, 1 / M folded in)
// mode 2: dst[o][k][i] = src[o][k][i] * w[k] for k < n                        ([outer][M][inner] -> [outer][n][inner])
template <class R> struct ChirpParams {
    const cx<R> *src;
    cx<R> *dst;
    const cx<R> *w;        // chirp, n entries (already conjugated for the inverse direction)
    const cx<R> *bspec;    // M entries
    int64_t n, M, inner, outer;
    int mode;
};
template <class R> struct ChirpBody {
    static HD void run(const BlockCtx &c, const ChirpParams<R> &p)
    {
        const int64_t len = p.mode == 2 ? p.n : p.M, total = p.outer * len * p.inner;
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int64_t i = e % p.inner, j = (e / p.inner) % len, o = e / (p.inner * len);
            if (p.mode == 0) p.dst[e] = j < p.n ? cmul(p.src[(o * p.n + j) * p.inner + i], p.w[j]) : cx<R>{(R)0, (R)0};
            else if (p.mode == 1) p.dst[e] = cmul(p.dst[e], p.bspec[j]);
            else p.dst[e] = cmul(p.src[(o * p.M + j) * p.inner + i], p.w[j]);
        }
    }
};

// out[e] += add[e] over `n` real scalars (float or double): accumulation of the partial convolutions of a split kernel
template <class R> struct AccumParams { R *out; const R *add; int64_t n; };
template <class R> struct AccumBody {
    static HD void run(const BlockCtx &c, const AccumParams<R> &p)
    {
        for (int64_t e = c.bid * c.nt + c.tid; e < p.n; e += c.nb * c.nt) p.out[e] = p.out[e] + p.add[e];
    }
};
}  // namespace ndc
