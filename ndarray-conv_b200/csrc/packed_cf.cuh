// packed_cf.cuh -- complex f32 arithmetic on sm_100a's packed FP32 pipe (device only).
//
// Blackwell executes add/mul/fma on a PAIR of f32 held in an aligned 64-bit register pair with one instruction
// (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  The SASS operand selectors (.HI_LO / .LO_HI swap,
// per-half negation, scalar broadcast) make a complex value (re = low half, im = high half) a natural operand:
//   complex add / sub          1 FADD2
//   a +/- (-i) b               1 FADD2 (swap + sign pattern folded into the operand by ptxas)
//   complex multiply           FMUL2 + FFMA2
// so a radix-32 butterfly costs about half the issue slots of the scalar formulation in kernels_fft.h.  The layout in memory
// is the same float2 {re, im}, so values are loaded and stored as 64-bit words without packing instructions.
#pragma once
#include <cstdint>

#ifdef NDCONV_CUDA
namespace ndc {
namespace pk {

struct pcf { uint64_t v; };    // (re, im) packed: re in bits 0..31

__device__ __forceinline__ pcf mk(float re, float im) { pcf r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(re), "f"(im)); return r; }
__device__ __forceinline__ float re(pcf a) { float x; asm("{ .reg .b32 t; mov.b64 {%0, t}, %1; }" : "=f"(x) : "l"(a.v)); return x; }
__device__ __forceinline__ float im(pcf a) { float y; asm("{ .reg .b32 t; mov.b64 {t, %0}, %1; }" : "=f"(y) : "l"(a.v)); return y; }
__device__ __forceinline__ pcf add(pcf a, pcf b) { pcf r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pcf sub(pcf a, pcf b) { pcf r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pcf mul(pcf a, pcf b) { pcf r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pcf fma(pcf a, pcf b, pcf c) { pcf r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

__device__ __forceinline__ pcf swap(pcf a) { return mk(im(a), re(a)); }
__device__ __forceinline__ pcf conj(pcf a) { return mk(re(a), -im(a)); }
__device__ __forceinline__ pcf neg(pcf a) { return mk(-re(a), -im(a)); }
// a * (-i) (forward) / a * (+i) (inverse)
template <bool INV> __device__ __forceinline__ pcf mul_neg_i(pcf a) { return INV ? mk(-im(a), re(a)) : mk(im(a), -re(a)); }
__device__ __forceinline__ pcf scale(pcf a, float s) { return mul(a, mk(s, s)); }
// a * w = (a.re, a.im) * wre + (-a.im, a.re) * wim.  Written with the sign on the swapped COPY OF a (not on the twiddle): ptxas folds
// the swap and the half-negation into the FFMA2 operand (`-R.F32x2.LO_HI.NP`) and broadcasts wim as a scalar, so a multiply by a
// run-time twiddle is exactly FMUL2 + FFMA2.  The earlier form fma(swap(a), (-wim, wim), ...) cost an FADD and a MOV per twiddle to
// build the (-wim, wim) pair -- and made them the first consumers of the twiddle's LDS.  Same roundings, bit-identical results.
__device__ __forceinline__ pcf cmul(pcf a, float wre, float wim) { return fma(mk(-im(a), re(a)), mk(wim, wim), mul(a, mk(wre, wre))); }
__device__ __forceinline__ pcf cmul(pcf a, pcf w) { return cmul(a, re(w), im(w)); }
// a * conj(w)
__device__ __forceinline__ pcf cmulc(pcf a, pcf w) { return cmul(a, re(w), -im(w)); }

template <bool INV> __device__ __forceinline__ void dft2(pcf *v) { const pcf a = v[0], b = v[1]; v[0] = add(a, b); v[1] = sub(a, b); }
template <bool INV> __device__ __forceinline__ void dft4(pcf *v)
{
    const pcf a0 = add(v[0], v[2]), a1 = sub(v[0], v[2]), a2 = add(v[1], v[3]), a3 = mul_neg_i<INV>(sub(v[1], v[3]));
    v[0] = add(a0, a2); v[2] = sub(a0, a2); v[1] = add(a1, a3); v[3] = sub(a1, a3);
}
template <bool INV> __device__ __forceinline__ void dft8(pcf *v)
{
    pcf e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    dft4<INV>(e); dft4<INV>(o);
    constexpr float h = 0.70710678118654752440f;
    // o1 * (h, -+h) = h (o1 + (-+i) o1) ; o3 * (-h, -+h) = h ((-+i) o3 - o3)
    o[1] = scale(add(o[1], mul_neg_i<INV>(o[1])), h);
    o[2] = mul_neg_i<INV>(o[2]);
    o[3] = scale(sub(mul_neg_i<INV>(o[3]), o[3]), h);
#pragma unroll
    for (int q = 0; q < 4; q++) { v[q] = add(e[q], o[q]); v[q + 4] = sub(e[q], o[q]); }
}

__device__ constexpr float w32c(int m)
{
    constexpr double t[32] = {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476, 0.5555702330196023, 0.38268343236508984, 0.19509032201612833, 0.0, -0.1950903220161282, -0.3826834323650897, -0.555570233019602, -0.7071067811865475, -0.8314696123025453, -0.9238795325112867, -0.9807852804032304, -1.0, -0.9807852804032304, -0.9238795325112868, -0.8314696123025455, -0.7071067811865477, -0.5555702330196022, -0.38268343236509034, -0.19509032201612866, 0.0, 0.1950903220161283, 0.38268343236509, 0.5555702330196018, 0.7071067811865474, 0.8314696123025452, 0.9238795325112865, 0.9807852804032303};
    return (float)t[m & 31];
}
__device__ constexpr float w32s(int m) { return w32c(m - 8); }    // sin(2 pi m / 32) = cos(2 pi (m - 8) / 32)

// multiply by W_32^m (forward) or its conjugate (inverse); m is a compile-time constant after unrolling
template <bool INV> __device__ __forceinline__ pcf tw32(pcf a, int m)
{
    m &= 31;
    if (m == 0) return a;
    if (m == 8) return mul_neg_i<INV>(a);
    if (m == 16) return mk(-re(a), -im(a));
    if (m == 24) return mul_neg_i<!INV>(a);
    constexpr float h = 0.70710678118654752440f;
    if (m == 4) return scale(add(a, mul_neg_i<INV>(a)), h);
    if (m == 12) return scale(sub(mul_neg_i<INV>(a), a), h);
    return cmul(a, w32c(m), INV ? w32s(m) : -w32s(m));
}

// Cooley-Tukey N = N1 * N2 in registers: n = N2 n1 + n2, k = k1 + N1 k2; natural order in and out
template <bool INV, int N1, int N2> __device__ __forceinline__ void dft_ct(pcf *v)
{
    constexpr int N = N1 * N2;
    pcf a[N2][N1];
#pragma unroll
    for (int n2 = 0; n2 < N2; n2++) {
#pragma unroll
        for (int n1 = 0; n1 < N1; n1++) a[n2][n1] = v[N2 * n1 + n2];
        if constexpr (N1 == 4) dft4<INV>(a[n2]); else if constexpr (N1 == 2) dft2<INV>(a[n2]); else dft8<INV>(a[n2]);
    }
#pragma unroll
    for (int k1 = 0; k1 < N1; k1++) {
        pcf b[N2];
#pragma unroll
        for (int n2 = 0; n2 < N2; n2++) b[n2] = tw32<INV>(a[n2][k1], n2 * k1 * (32 / N));
        if constexpr (N2 == 4) dft4<INV>(b); else if constexpr (N2 == 2) dft2<INV>(b); else dft8<INV>(b);
#pragma unroll
        for (int k2 = 0; k2 < N2; k2++) v[k1 + N1 * k2] = b[k2];
    }
}
template <bool INV, int RDX> __device__ __forceinline__ void dft(pcf *v)
{
    static_assert(RDX == 2 || RDX == 4 || RDX == 8 || RDX == 16 || RDX == 32, "packed DFT: power-of-two radix up to 32");
    if constexpr (RDX == 2) dft2<INV>(v);
    else if constexpr (RDX == 4) dft4<INV>(v);
    else if constexpr (RDX == 8) dft8<INV>(v);
    else if constexpr (RDX == 16) dft_ct<INV, 4, 4>(v);
    else dft_ct<INV, 4, 8>(v);
}

}  // namespace pk
}  // namespace ndc
#endif
