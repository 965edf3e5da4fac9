// kernels_fft_opt.cuh -- sm_100a fast path of the 2-D real conv_fft pipeline (device only).
//
// Tile = F0 x F1 = 1024 x 2048 real samples (overlap-save in both axes).  A row of 2048 reals is packed as
// L = 1024 complex z[j] = x[2j] + i x[2j+1]; the R2C post-processing and the C2R pre-processing of the
// half-length trick are NOT done per row: they are folded, together with the multiply by the cached kernel
// spectrum, into the column kernel, where column k and column L-k of the packed spectrum sit side by side.
//
//   row_fwd_packed   one warp per row: border-mapped (or plain vector) loads straight into registers, radix-32 x
//                    radix-32 Stockham with one warp-private shared-memory transpose, warp-shuffle pairing, 16-byte
//                    stores in the PAIRED layout: slot k in [0,512) holds (Z[k], Z[L-k]); slot 0 holds (Z[0], Z[L/2]).
//   col_pair_fmi     256 threads per 1024 x 8-column tile (= 4 slots): radix-32 x radix-32 forward along the strided
//                    axis, X = E + w^k O, Y = X .* K, re-packing, radix-32 x radix-32 inverse, in place.
//   row_inv_packed   one warp per (output row, tile): paired loads, inverse radix-32 x radix-32, crop [Kd-1, F) and
//                    stride decimation fused into the (vector) store.
//   kpair_repack     kernel spectrum [F0][L+1] (generic path, cached) -> Kpair[q][slot] = (K[q][k], K[-q][L-k]).
//
// Reference stages replaced: conv_fft/padding.rs:30-62, processor/real.rs:105-154, mod.rs:268, real.rs:233-280,
// mod.rs:282-289 (see kernels_fft.h for the stage-by-stage citations).
#pragma once
#include "kernels_fft.h"

#ifdef NDCONV_CUDA
namespace ndc {
namespace opt {

constexpr int kL = 1024;         // packed complex row length
constexpr int kF1 = 2048;        // real row tile
constexpr int kF0 = 1024;        // column tile
constexpr int kSlots = kL / 2;   // pair slots per row
constexpr int kKP = kSlots + 8;  // Kpair row pitch in float4 slots (slot kSlots = the self-paired middle column)

typedef cx<float> cf;

__device__ __forceinline__ cf ld_cf(const cf *p) { float2 t = *reinterpret_cast<const float2 *>(p); return cf{t.x, t.y}; }
__device__ __forceinline__ void st_cf(cf *p, cf v) { *reinterpret_cast<float2 *>(p) = make_float2(v.re, v.im); }

struct RowOptParams {
    // geometry (2-D)
    int64_t n[2], xstr[2], P[2], pf[2];
    const int32_t *map[2];
    float cfront[2], cback[2];
    int V[2], ntiles[2], Kd[2];
    int64_t s[2], O[2];
    const float *x;
    float *out;
    cf *ws;                 // [tile][kF0][kL] paired layout
    const cf *tw;           // exp(-2 pi i j / 1024), j < 1024
    int64_t nwork;
};

// s_tw[k1*32 + t] = W_1024^{t k1}
__device__ __forceinline__ void load_tw_table(cf *s_tw, const cf *tw, int tid, int nt)
{
    for (int idx = tid; idx < 1024; idx += nt) s_tw[idx] = tw[(idx >> 5) * (idx & 31)];
}

// ---- row forward -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4) row_fwd_packed(const __grid_constant__ RowOptParams p)
{
    __shared__ cf s_tw[1024];
    __shared__ cf s_buf[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tw_table(s_tw, p.tw, threadIdx.x, blockDim.x);
    __syncthreads();
    cf *sb = s_buf[warp];
    const int src_lane = (32 - lane) & 31;
    for (int64_t w = (int64_t)blockIdx.x * 4 + warp; w < p.nwork; w += (int64_t)gridDim.x * 4) {
        const int64_t tile = w / kF0;
        const int r = (int)(w % kF0);
        const int t0 = (int)(tile / p.ntiles[1]), t1 = (int)(tile % p.ntiles[1]);
        const int64_t c0 = (int64_t)t0 * p.V[0] + r;         // padded row
        const int64_t cl0 = (int64_t)t1 * p.V[1];            // first padded column of the tile
        float4 *dst = reinterpret_cast<float4 *>(p.ws + tile * ((int64_t)kF0 * kL) + (int64_t)r * kL);
        // resolve the row (axis 0)
        bool zero_row = false, row_const = false, row_init = false;
        float row_cval = 0.f;
        int64_t rowbase = 0;
        if (c0 >= p.P[0]) zero_row = true;
        else {
            const int32_t m0 = p.map[0][c0];
            if (m0 >= 0) rowbase = (int64_t)m0 * p.xstr[0];
            else if (m0 == NDC_MAP_INIT) row_init = true;
            else { row_const = true; row_cval = (m0 == NDC_MAP_CONST_FRONT) ? p.cfront[0] : p.cback[0]; }
        }
        if (zero_row) {
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) dst[lane + 32 * k2] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        cf v[32];
        const bool interior = !row_const && !row_init && p.xstr[1] == 1 && cl0 >= p.pf[1] && cl0 + kF1 <= p.pf[1] + p.n[1];
        if (interior) {
            const float *src = p.x + rowbase + (cl0 - p.pf[1]);
            if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
                const float2 *s2 = reinterpret_cast<const float2 *>(src);
#pragma unroll
                for (int j = 0; j < 32; j++) { float2 t = __ldg(s2 + lane + 32 * j); v[j] = cf{t.x, t.y}; }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) { const int e = 2 * (lane + 32 * j); v[j] = cf{__ldg(src + e), __ldg(src + e + 1)}; }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float q[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int64_t cl = cl0 + 2 * (lane + 32 * j) + h;
                    float val = 0.f;
                    if (cl < p.P[1]) {
                        const int32_t m = p.map[1][cl];
                        if (m == NDC_MAP_CONST_FRONT) val = p.cfront[1];
                        else if (m == NDC_MAP_CONST_BACK) val = p.cback[1];
                        else if (row_const) val = row_cval;
                        else if (m == NDC_MAP_INIT || row_init) val = 0.f;
                        else val = __ldg(p.x + rowbase + (int64_t)m * p.xstr[1]);
                    }
                    q[h] = val;
                }
                v[j] = cf{q[0], q[1]};
            }
        }
        dft32<float>(v, false);                                  // over j -> k1
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) sb[k1 * 33 + lane] = cmul(v[k1], s_tw[k1 * 32 + lane]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = sb[lane * 33 + i];
        __syncwarp();
        dft32<float>(v, false);                                  // over t -> k2: v[k2] = Z[lane + 32 k2]
        // pair (Z[k], Z[L-k]): the partner lives in lane (32-lane)&31, register 31-k2 (lane 0: register 32-k2; k = 0 pairs with L/2)
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            float px = __shfl_sync(0xffffffffu, v[31 - k2].re, src_lane);
            float py = __shfl_sync(0xffffffffu, v[31 - k2].im, src_lane);
            if (lane == 0) { const cf o = (k2 == 0) ? v[16] : v[(32 - k2) & 31]; px = o.re; py = o.im; }
            dst[lane + 32 * k2] = make_float4(v[k2].re, v[k2].im, px, py);
        }
    }
}

// ---- row inverse + crop + decimate ---------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4) row_inv_packed(const __grid_constant__ RowOptParams p)
{
    __shared__ cf s_tw[1024];
    __shared__ cf s_buf[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tw_table(s_tw, p.tw, threadIdx.x, blockDim.x);
    __syncthreads();
    cf *sb = s_buf[warp];
    const int src_lane = (32 - lane) & 31;
    const int ntl = p.ntiles[1];
    for (int64_t w = (int64_t)blockIdx.x * 4 + warp; w < p.nwork; w += (int64_t)gridDim.x * 4) {
        const int t1 = (int)(w % ntl);
        const int64_t o0 = w / ntl;
        const int64_t q0 = o0 * p.s[0];
        const int64_t t0 = q0 / p.V[0];
        const int r0 = (int)(q0 - t0 * p.V[0]) + p.Kd[0] - 1;
        const int64_t tile = t0 * ntl + t1;
        const float4 *src = reinterpret_cast<const float4 *>(p.ws + tile * ((int64_t)kF0 * kL) + (int64_t)r0 * kL);
        cf v[32], b[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const float4 t = src[lane + 32 * k2];
            v[k2] = cf{t.x, t.y};
            b[k2] = cf{t.z, t.w};
        }
        // v[j'] for j' >= 16 is Zy[lane + 32 j'] = the partner half loaded by lane (32-lane)&31 at index 31-j'
#pragma unroll
        for (int jp = 16; jp < 32; jp++) {
            float px = __shfl_sync(0xffffffffu, b[31 - jp].re, src_lane);
            float py = __shfl_sync(0xffffffffu, b[31 - jp].im, src_lane);
            if (lane == 0) { const cf o = (jp == 16) ? b[0] : b[32 - jp]; px = o.re; py = o.im; }
            v[jp] = cf{px, py};
        }
        dft32<float>(v, true);                                   // over j -> n1
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) sb[n1 * 33 + lane] = cmulc(v[n1], s_tw[n1 * 32 + lane]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = sb[lane * 33 + i];
        __syncwarp();
        dft32<float>(v, true);                                   // v[n2] = z[lane + 32 n2] = (y[2n], y[2n+1])
        // crop [Kd-1, F1) of the tile; global position m = t1*V + i; keep m < P and (m-Kd+1) % s == 0
        const int Kd1 = p.Kd[1];
        const int64_t orow = o0 * p.O[1];
        const int64_t mbase = (int64_t)t1 * p.V[1];
        if (p.s[1] == 1) {
            const int64_t obase = orow + mbase - (Kd1 - 1);       // output element of local sample 0
            const bool vec_ok = (obase & 1) == 0;
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) {
                const int i = 2 * (lane + 32 * n2);
                const int64_t o_lo = mbase + i - (Kd1 - 1);       // output column of sample i
                const bool ok0 = i >= Kd1 - 1 && o_lo < p.O[1];
                const bool ok1 = i + 1 >= Kd1 - 1 && o_lo + 1 < p.O[1];
                if (vec_ok && ok0 && ok1) *reinterpret_cast<float2 *>(p.out + orow + o_lo) = make_float2(v[n2].re, v[n2].im);
                else {
                    if (ok0) p.out[orow + o_lo] = v[n2].re;
                    if (ok1) p.out[orow + o_lo + 1] = v[n2].im;
                }
            }
        } else {
            const int64_t s1 = p.s[1];
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * (lane + 32 * n2) + h;
                    if (i < Kd1 - 1) continue;
                    const int64_t q = mbase + i - (Kd1 - 1);
                    if (q % s1) continue;
                    const int64_t o = q / s1;
                    if (o < p.O[1]) p.out[orow + o] = h ? v[n2].im : v[n2].re;
                }
            }
        }
    }
}

// ---- column forward * K * inverse ----------------------------------------------------------------------
struct ColOptParams {
    cf *ws;
    const float4 *kpair;    // [kF0][kKP]: (K[q][k], K[-q][L-k]); slot 0: (K[q][0], K[q][L]); slot kSlots: (K[q][L/2], K[-q][L/2])
    const cf *tw;           // exp(-2 pi i j / 1024)
    const cf *twr;          // exp(-2 pi i k / 2048), k <= 512
    int64_t ntiles_total;
    int64_t nwork;          // ntiles_total * (kL / 8)
};

constexpr int kColPitch = 32 * 8 + 8;   // padded k1-row stride of the exchange buffer (complex elements)
constexpr int kColSmem = (1024 + 32 * kColPitch) * 8;

__global__ void __launch_bounds__(256, 2) col_pair_fmi(const __grid_constant__ ColOptParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *s_tw = reinterpret_cast<cf *>(smem_raw);        // 1024
    cf *S = s_tw + 1024;                                // 32 * kColPitch  (>= 1024 * 8 linear)
    const int tid = threadIdx.x;
    const int c = tid & 7, i = tid >> 3;
    load_tw_table(s_tw, p.tw, tid, blockDim.x);
    __syncthreads();
    for (int64_t w = blockIdx.x; w < p.nwork; w += gridDim.x) {
        const int cb = (int)(w % (kL / 8));
        const int64_t tile = w / (kL / 8);
        cf *g = p.ws + tile * ((int64_t)kF0 * kL) + cb * 8 + c;
        cf v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = ld_cf(g + (int64_t)(i + 32 * j) * kL);
        // ---- forward: radix 32 over j, twiddle, exchange, radix 32 over i ----
        dft32<float>(v, false);
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) S[k1 * kColPitch + i * 8 + c] = cmul(v[k1], s_tw[k1 * 32 + i]);
        __syncthreads();
#pragma unroll
        for (int ii = 0; ii < 32; ii++) v[ii] = S[i * kColPitch + ii * 8 + c];
        dft32<float>(v, false);                            // v[k2] = Zhat[q = i + 32 k2][column]
        __syncthreads();
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) S[(i + 32 * k2) * 8 + c] = v[k2];
        __syncthreads();
        // ---- pair algebra: X = E + w^k O ; Y = X K ; re-pack ----
        const int slot = cb * 4 + (c >> 1);
        const int odd = c & 1;
        const int role = slot ? odd : 2 + odd;               // 0 primary, 1 secondary, 2 DC/Nyquist, 3 middle (self-paired)
        const int bcol = (role <= 1) ? (c ^ 1) : c;
        const cf wk = p.twr[role == 3 ? kSlots : slot];
        const int kslot = role == 3 ? kSlots : slot;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const int q = i + 32 * j, qm = (kF0 - q) & (kF0 - 1);
            const cf a = S[q * 8 + c], b = S[qm * 8 + bcol];
            const float4 kk = __ldg(p.kpair + (int64_t)(role == 1 ? qm : q) * kKP + kslot);
            const cf K0 = cf{kk.x, kk.y}, K1 = cf{kk.z, kk.w};
            const cf ap = (role == 1) ? b : a, bp = (role == 1) ? a : b;
            const cf E = cf{0.5f * (ap.re + bp.re), 0.5f * (ap.im - bp.im)};          // (ap + conj bp) / 2
            const cf O = cf{0.5f * (ap.im + bp.im), -0.5f * (ap.re - bp.re)};         // -i (ap - conj bp) / 2
            cf out;
            if (role == 2) {
                const cf Yd = cmul(cadd(E, O), K0), Yn = cmul(csub(E, O), K1);
                const cf sm = cadd(Yd, Yn), df = csub(Yd, Yn);
                out = cf{sm.re - df.im, sm.im + df.re};                               // sm + i df
            } else {
                const cf t = cmul(wk, O);
                const cf X1 = cadd(E, t), X2 = cconj(csub(E, t));
                const cf Y1 = cmul(X1, K0), Y2 = cmul(X2, K1);
                if (role == 1) {
                    const cf sm = cadd(Y2, cconj(Y1)), df = csub(Y2, cconj(Y1));
                    const cf u = cmul(wk, df);                                        // out = sm - i w df
                    out = cf{sm.re + u.im, sm.im - u.re};
                } else {
                    const cf sm = cadd(Y1, cconj(Y2)), df = csub(Y1, cconj(Y2));
                    const cf u = cmulc(df, wk);                                       // out = sm + i conj(w) df
                    out = cf{sm.re - u.im, sm.im + u.re};
                }
            }
            v[j] = out;
        }
        // ---- inverse: radix 32 over j, conj twiddle, exchange, radix 32 over i ----
        dft32<float>(v, true);
        __syncthreads();                                   // every thread has finished reading the linear buffer
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) S[n1 * kColPitch + i * 8 + c] = cmulc(v[n1], s_tw[n1 * 32 + i]);
        __syncthreads();
#pragma unroll
        for (int ii = 0; ii < 32; ii++) v[ii] = S[i * kColPitch + ii * 8 + c];
        dft32<float>(v, true);
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) st_cf(g + (int64_t)(i + 32 * n2) * kL, v[n2]);
        __syncthreads();                                   // S is rewritten by the next work item
    }
}

}  // namespace opt

// kernel spectrum [kF0][Hp] (bins 0..L) -> paired layout.  Block-stride body (host-emulable).
struct KpairParams {
    const cx<float> *kspec;
    float *kpair;     // float4 [F0][KP]
    int F0, L, Hp, KP;
};
struct KpairBody {
    static HD void run(const BlockCtx &c, const KpairParams &p)
    {
        const int S = p.L / 2;
        const int64_t total = (int64_t)p.F0 * (S + 1);
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int q = (int)(e / (S + 1)), slot = (int)(e % (S + 1));
            const int qm = (p.F0 - q) % p.F0;
            cx<float> a, b;
            if (slot == 0) { a = p.kspec[(int64_t)q * p.Hp]; b = p.kspec[(int64_t)q * p.Hp + p.L]; }
            else if (slot == S) { a = p.kspec[(int64_t)q * p.Hp + S]; b = p.kspec[(int64_t)qm * p.Hp + S]; }
            else { a = p.kspec[(int64_t)q * p.Hp + slot]; b = p.kspec[(int64_t)qm * p.Hp + (p.L - slot)]; }
            float *d = p.kpair + ((int64_t)q * p.KP + slot) * 4;
            d[0] = a.re; d[1] = a.im; d[2] = b.re; d[3] = b.im;
        }
    }
};
}  // namespace ndc
#endif  // NDCONV_CUDA
