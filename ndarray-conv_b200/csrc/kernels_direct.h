// kernels_direct.h -- direct N-d convolution (ConvExt::conv, src/conv/mod.rs:128-200).
//
// out[o] = sum_t B[o*s + off_t] * w_t over the compacted tap list (zero weights dropped, row-major,
// kernel flipped unless no_reverse -- gen_offset_list, src/dilation/mod.rs:34-60), accumulated in
// the element type in tap order with un-fused multiply/add, so results are bit-identical to the
// reference for integers (wrapping) AND floats.  B is the padded array of src/padding/mod.rs:84-117;
// it is never materialised: each coordinate goes through the per-axis border index map.
#pragma once
#include "common.h"

namespace ndc {

// ---- element arithmetic, un-fused -------------------------------------------------------------
template <class T> struct Elem;

#define NDC_INT_ELEM(T)                                                            \
    template <> struct Elem<T> {                                                   \
        static HD T zero() { return (T)0; }                                        \
        static HD T mac(T acc, T a, T w) { return (T)(acc + (T)(a * w)); }         \
    };
NDC_INT_ELEM(uint8_t)
NDC_INT_ELEM(uint16_t)
NDC_INT_ELEM(uint32_t)
NDC_INT_ELEM(uint64_t)
#undef NDC_INT_ELEM
// i128 / u128 (T: NumAssign + Copy admits them, src/conv/mod.rs:118-121): two 64-bit halves, wrapping multiply-add
typedef unsigned __int128 u128_t;
template <> struct Elem<u128_t> {
    static HD u128_t zero() { return (u128_t)0; }
    static HD u128_t mac(u128_t acc, u128_t a, u128_t w) { return acc + a * w; }
};

HD float f_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;   // host emulation is compiled with -ffp-contract=off
#endif
}
HD float f_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
HD double f_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
HD double f_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

template <> struct Elem<float> {
    static HD float zero() { return 0.0f; }
    static HD float mac(float acc, float a, float w) { return f_add(acc, f_mul(a, w)); }
};
template <> struct Elem<double> {
    static HD double zero() { return 0.0; }
    static HD double mac(double acc, double a, double w) { return f_add(acc, f_mul(a, w)); }
};
template <class R> struct Elem<cx<R>> {
    static HD cx<R> zero() { return cx<R>{(R)0, (R)0}; }
    // num::Complex Mul: (re*re - im*im, re*im + im*re), then +=
    static HD cx<R> mac(cx<R> acc, cx<R> a, cx<R> w)
    {
        R pr = f_add(f_mul(a.re, w.re), -f_mul(a.im, w.im));
        R pi = f_add(f_mul(a.re, w.im), f_mul(a.im, w.re));
        return cx<R>{f_add(acc.re, pr), f_add(acc.im, pi)};
    }
};

struct DirectParams {
    int ndim, ntap;
    const void *x;
    void *out;
    int64_t xstr[NDC_MAX_DIM], n[NDC_MAX_DIM], P[NDC_MAX_DIM], pf[NDC_MAX_DIM], Kd[NDC_MAX_DIM], s[NDC_MAX_DIM], O[NDC_MAX_DIM];
    const int32_t *map[NDC_MAX_DIM];   // border index maps, length P[a]
    const int32_t *tap_off;            // [ntap][NDC_MAX_DIM]
    const int64_t *tap_lin;            // [ntap]
    const void *tap_w;                 // [ntap] T
    unsigned char cfront[NDC_MAX_DIM][16], cback[NDC_MAX_DIM][16];
    int64_t total;
};

// value of the padded array at padded coordinate c[] (highest-numbered constant axis wins;
// never-written cells read the zero-initialised buffer): SURVEY A.3 / DESIGN.md "border maps".
template <class T> HD T padded_at(const DirectParams &p, const int64_t *c)
{
    const T *x = (const T *)p.x;
    int64_t src = 0;
    bool init = false;
    for (int a = p.ndim - 1; a >= 0; a--) {
        int32_t m = p.map[a][c[a]];
        if (m >= 0) src += (int64_t)m * p.xstr[a];
        else if (m == NDC_MAP_INIT) init = true;
        else return *(const T *)(m == NDC_MAP_CONST_FRONT ? p.cfront[a] : p.cback[a]);
    }
    return init ? Elem<T>::zero() : x[src];
}

template <class T> struct DirectBody {
    static HD void run(const BlockCtx &c, const DirectParams &p)
    {
        const T *x = (const T *)p.x;
        const T *w = (const T *)p.tap_w;
        T *out = (T *)p.out;
        for (int64_t e = c.bid * c.nt + c.tid; e < p.total; e += c.nb * c.nt) {
            int64_t o[NDC_MAX_DIM], b[NDC_MAX_DIM];
            int64_t r = e;
            bool interior = true;
            int64_t base = 0;
            for (int a = p.ndim - 1; a >= 0; a--) {
                o[a] = r % p.O[a]; r /= p.O[a];
                b[a] = o[a] * p.s[a];                       // first padded coordinate of the window
                interior = interior && (b[a] >= p.pf[a]) && (b[a] + p.Kd[a] <= p.pf[a] + p.n[a]);
                base += (b[a] - p.pf[a]) * p.xstr[a];
            }
            T acc = Elem<T>::zero();
            if (interior) {
                for (int t = 0; t < p.ntap; t++) acc = Elem<T>::mac(acc, x[base + p.tap_lin[t]], w[t]);
            } else {
                for (int t = 0; t < p.ntap; t++) {
                    int64_t cc[NDC_MAX_DIM];
                    for (int a = 0; a < p.ndim; a++) cc[a] = b[a] + p.tap_off[t * NDC_MAX_DIM + a];
                    acc = Elem<T>::mac(acc, padded_at<T>(p, cc), w[t]);
                }
            }
            out[e] = acc;
        }
    }
};

}  // namespace ndc
