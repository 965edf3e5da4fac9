// kernels_fft_fast_f64.cuh -- sm_100a fast path of conv_fft for real f64 problems of rank 2 and 3.
//
// The reference builds real and complex processors for every FftNum (src/conv_fft/processor/mod.rs:133-143) and gates f64 results at
// 1e-9 (src/conv_fft/tests.rs:15-16).  The f32 fast path (kernels_fft_fast.cuh) holds 32 complex values per thread and feeds the packed
// FP32 pipe; 32 complex doubles would be the whole register file, so these kernels keep the same pass structure -- one lane group per
// row with no block barrier, an in-place column pass on 8-column blocks with the spectral multiply fused between the forward and the
// inverse transform, crop + stride fused into the last store -- with SIXTEEN values per thread:
//   row_fwd_d<T,N>  rows of 2 L real samples, L = 16 T complex (T = 4 / 8 / 16 lanes per row -> tiles of 128 / 256 / 512 samples):
//                   border-mapped loads (16 bytes per lane where the row is plain array data), radix 16 -> twiddle -> exchange ->
//                   radix T, R2C post-processing through lane shuffles, paired (k, L-k) 32-byte stores
//   col_pass_d<E,Tc> columns of F = E Tc <= 256 rows in place, modes FWD / INV / FWD x kspec x INV, next item staged with cp.async
//   row_inv_d<T,N>  the mirror image of row_fwd_d: C2R pre-processing, radix T -> conj twiddle -> exchange -> radix 16, crop / stride
// Workspace rows: L + 8 complex doubles (column 0 <- bin 0, 1 <- bin L/2, (2s, 2s+1) <- bins (s, L-s), column L <- the Nyquist bin,
// then zeros), the layout of the f32 path with 16-byte elements.  Arithmetic is plain FP64 (DADD / DMUL / DFMA; B200: 64 per clock and SM).
#pragma once
#include "kernels_fft_fast.cuh"

#ifdef NDCONV_CUDA
namespace ndc {
namespace fast64 {

typedef cx<double> cd;
constexpr int R1 = 16;           // values per thread = radix of the in-register butterfly
constexpr int kPadD = 8;         // extra columns per workspace row (Nyquist + 7 zeros): rows stay 128-byte aligned and a multiple of the column pass's 8-column blocks

using fast::pdl_launch_dependents;
using fast::pdl_wait;
using fast::cp_async16;
using fast::cp_async_commit;
using fast::cp_async_wait_all;

template <bool INV, int RDX> __device__ __forceinline__ void dftd(cd *v)
{
    static_assert(RDX == 2 || RDX == 4 || RDX == 8 || RDX == 16, "f64 fast path: power-of-two radix up to 16");
    if constexpr (RDX == 2) dft2(v);
    else if constexpr (RDX == 4) dft4(v, INV);
    else if constexpr (RDX == 8) dft8(v, INV);
    else dft16(v, INV);
}
__device__ __forceinline__ cd ld_cd(const cd *p) { const double2 q = *reinterpret_cast<const double2 *>(p); return cd{q.x, q.y}; }
__device__ __forceinline__ void st_cd(cd *p, cd v) { *reinterpret_cast<double2 *>(p) = make_double2(v.re, v.im); }
// dst[k * dst_stride] = v[k] * tw[k * tw_stride] (CONJ: times the conjugate), k < N, both in shared memory, with the twiddles fetched B at
// a time ahead of the stores that follow them (fast::twiddle_store: the compiler cannot hoist a shared-memory load over a shared-memory store)
template <int N, bool CONJ, int B = 4>
__device__ __forceinline__ void twiddle_store_d(const cd *v, const cd *tw, int tw_stride, cd *dst, int dst_stride);
__device__ __forceinline__ double shfl_d(double x, int src) { return __shfl_sync(0xffffffffu, x, src); }

struct RowParamsD {
    int ndim;                                  // 2 or 3
    int64_t n[3], xstr[3], P[3], pf[3];
    const int32_t *map[3];
    double cfront[3], cback[3];
    int F[3], V[3], ntiles[3], Kd[3];          // F[ndim-1] = 2L
    int64_t s[3], O[3];
    const double *x;
    double *out;
    cd *ws;
    const cd *tw;                              // exp(-2 pi i j / L), j < L
    const cd *twr;                             // exp(-2 pi i k / 2L), k <= L/2
    int64_t nwork;                             // rows to transform (fwd) / (output row, tile) pairs (inv)
    int64_t rows_per_tile, tile_elems;         // prod of the outer F ; rows_per_tile * (L + 8)
    int64_t xstr_batch;                        // same-shape batch (see fast::RowParams)
};

template <int N, bool CONJ, int B>
__device__ __forceinline__ void twiddle_store_d(const cd *v, const cd *tw, int tw_stride, cd *dst, int dst_stride)
{
#pragma unroll
    for (int k0 = 0; k0 < N; k0 += B) {
        cd w[B];
#pragma unroll
        for (int e = 0; e < B; e++) if (k0 + e < N) w[e] = ld_cd(tw + (k0 + e) * tw_stride);
#pragma unroll
        for (int e = 0; e < B; e++) if (k0 + e < N) st_cd(dst + (k0 + e) * dst_stride, CONJ ? cmulc(v[k0 + e], w[e]) : cmul(v[k0 + e], w[e]));
    }
}

template <int T> struct RowCfgD {
    static constexpr int L = R1 * T, M = R1 / T, G = 32 / T;          // complex length, radix-T butterflies per lane, rows per warp
    static constexpr int gstride = R1 * (T + 1) + T;                  // exchange buffer of one lane group (complex elements)
    static constexpr int wstride = (G * gstride > (L + 2) * G ? G * gstride : (L + 2) * G);   // row_inv staging: L/2 + 1 pairs per group
    static constexpr int smem = (L + L / 2 + 4 * wstride) * 16;       // W_L table, w^k table, 4 warps of exchange buffers
};

struct RowSrcInfoD {
    int64_t base, cl0;
    double cval;
    bool zero, has_const, beyond, active;
    cd *dst;
};

// outer-axis resolution of one tile row (the f64 twin of fast::resolve_fwd_row)
template <int N> __device__ __forceinline__ RowSrcInfoD resolve_fwd_row_d(const RowParamsD &p, int64_t w, int pitch)
{
    RowSrcInfoD r; r.base = 0; r.cval = 0.0; r.zero = false; r.has_const = false; r.beyond = false; r.active = w < p.nwork; r.dst = nullptr; r.cl0 = 0;
    if (!r.active) return r;
    const uint32_t w32 = (uint32_t)w, rpt = (uint32_t)p.rows_per_tile;
    const uint32_t tile = w32 / rpt;
    uint32_t row = w32 - tile * rpt;
    r.dst = p.ws + (int64_t)tile * p.tile_elems + (int64_t)row * pitch;
    uint32_t tt = tile;
    const uint32_t tl = tt % (uint32_t)p.ntiles[N - 1]; tt /= (uint32_t)p.ntiles[N - 1];
    r.cl0 = (int64_t)tl * p.V[N - 1];
    int64_t c[2] = {0, 0};
#pragma unroll
    for (int a = N - 2; a >= 0; a--) {
        const uint32_t ta = tt % (uint32_t)p.ntiles[a]; tt /= (uint32_t)p.ntiles[a];
        const uint32_t ra = row % (uint32_t)p.F[a]; row /= (uint32_t)p.F[a];
        c[a] = (int64_t)ta * p.V[a] + ra;
        if (c[a] >= p.P[a]) r.beyond = true;
    }
    if (r.beyond) return r;
    r.base = (int64_t)tt * p.xstr_batch;
#pragma unroll
    for (int a = N - 2; a >= 0; a--) {
        if (r.has_const) continue;
        const int32_t m = p.map[a][c[a]];
        if (m >= 0) r.base += (int64_t)m * p.xstr[a];
        else if (m == NDC_MAP_INIT) r.zero = true;
        else { r.has_const = true; r.cval = (m == NDC_MAP_CONST_FRONT) ? p.cfront[a] : p.cback[a]; }
    }
    return r;
}

// two neighbouring samples (padded columns cl0, cl0 + 1 of the last axis) of a row that is not plain array data there
template <int N> __device__ __noinline__ cd border_pair_d(const RowParamsD &p, int64_t base, int flags, double cval, int64_t cl0)
{
    constexpr int al = N - 1;
    const bool active = flags & 1, beyond = flags & 2, zero = flags & 4, has_const = flags & 8;
    const bool plain = active && !beyond && !zero && !has_const;
    double q[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int64_t cl = cl0 + h;
        double val = 0.0;
        const int64_t cc = cl - p.pf[al];
        if (plain && cc >= 0 && cc < p.n[al]) val = __ldg(p.x + base + cc * p.xstr[al]);
        else if (active && !beyond && cl < p.P[al]) {
            const int32_t m = p.map[al][cl];
            if (m == NDC_MAP_CONST_FRONT) val = p.cfront[al];
            else if (m == NDC_MAP_CONST_BACK) val = p.cback[al];
            else if (has_const) val = cval;
            else if (m != NDC_MAP_INIT && !zero) val = __ldg(p.x + base + (int64_t)m * p.xstr[al]);
        }
        q[h] = val;
    }
    return cd{q[0], q[1]};
}

// ---- row forward ---------------------------------------------------------------------------------------------------------
template <int T, int N>
__global__ void __launch_bounds__(128, 4) row_fwd_d(const __grid_constant__ RowParamsD p)
{
    pdl_launch_dependents();
    constexpr int WPB = 4;
    constexpr int L = RowCfgD<T>::L, M = RowCfgD<T>::M, G = RowCfgD<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *s_tw = reinterpret_cast<cd *>(smem_raw);          // s_tw[k1 * T + t] = W_L^{t k1}
    cd *s_twr = s_tw + L;                                 // (-i/2) w^k, k < L/2
    cd *s_ex = s_twr + L / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) st_cd(s_tw + idx, ld_cd(p.tw + (idx / T) * (idx % T)));
    for (int idx = threadIdx.x; idx < L / 2; idx += blockDim.x) { const cd w = ld_cd(p.twr + idx); st_cd(s_twr + idx, cd{0.5 * w.im, -0.5 * w.re}); }
    __syncthreads();
    pdl_wait();
    cd *sb = s_ex + warp * RowCfgD<T>::wstride + g * RowCfgD<T>::gstride;
    constexpr int al = N - 1;
    const int src_lane = g * T + ((T - t) % T);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    const int64_t wstep = (int64_t)gridDim.x * WPB;
    int64_t wi = (int64_t)blockIdx.x * WPB + warp;
    RowSrcInfoD nxt = resolve_fwd_row_d<N>(p, wi * G + g, L + kPadD);
    for (; wi < nwarp_items; wi += wstep) {
        const RowSrcInfoD ri = nxt;
        if (wi + wstep < nwarp_items) nxt = resolve_fwd_row_d<N>(p, (wi + wstep) * G + g, L + kPadD);
        if (__all_sync(0xffffffffu, !ri.active || ri.beyond || ((ri.has_const ? ri.cval == 0.0 : ri.zero) && p.cfront[al] == 0.0 && p.cback[al] == 0.0))) {
            // rows beyond the padded extent of an outer axis, and all-zero rows (a Zeros border or a never-written plane of an outer axis, no
            // non-zero constant border on the last axis): zero spectrum, no transform
            if (ri.active) for (int q = t; q < L + kPadD; q += T) st_cd(ri.dst + q, cd{0.0, 0.0});
            continue;
        }
        cd v[R1];
        {
            const bool plain = ri.active && !ri.beyond && !ri.zero && !ri.has_const && p.xstr[al] == 1;
            const double *rowp = p.x + ri.base - p.pf[al];                // padded column cl of this row is rowp[cl] where it is an array sample
            const int64_t lo = p.pf[al], hi = p.pf[al] + p.n[al];
            const int flags = (ri.active ? 1 : 0) | (ri.beyond ? 2 : 0) | (ri.zero ? 4 : 0) | (ri.has_const ? 8 : 0);
            if (!plain && p.xstr[al] == 1) {
                // a row without array samples (an outer axis sits in a Zeros / Const border or a never-written plane): its value depends on
                // the column only through the last axis' own constant borders -- no loads from x, map lookups in the border columns only, no calls (these are most
                // rows of a small padded problem: every pair of them through border_pair cost c3 12 us)
                const double rowval = ri.has_const ? ri.cval : 0.0;
                const bool live = ri.active && !ri.beyond;
                if (!live || (rowval == 0.0 && p.cfront[al] == 0.0 && p.cback[al] == 0.0)) {
                    // identically zero: only reached when the row shares its warp with rows that do have samples, so it must cost nothing (see fast::row_fwd)
#pragma unroll
                    for (int j = 0; j < R1; j++) v[j] = cd{0.0, 0.0};
                } else
#pragma unroll
                for (int j = 0; j < R1; j++) {
                    const int64_t cl = ri.cl0 + 2 * (t + T * j);
                    double q[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int64_t c1 = cl + h;
                        double val = 0.0;
                        if (live && c1 < p.P[al]) {
                            val = rowval;
                            if (c1 < lo || c1 >= hi) {                 // the last axis' own border: a constant fill there wins
                                const int32_t m = p.map[al][c1];
                                if (m == NDC_MAP_CONST_FRONT) val = p.cfront[al];
                                else if (m == NDC_MAP_CONST_BACK) val = p.cback[al];
                            }
                        }
                        q[h] = val;
                    }
                    v[j] = cd{q[0], q[1]};
                }
            } else
#pragma unroll
            for (int j = 0; j < R1; j++) {
                const int64_t cl = ri.cl0 + 2 * (t + T * j);
                if (plain && cl >= lo && cl + 1 < hi) {
                    const double *s = rowp + cl;
                    if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) { const double2 q = __ldg(reinterpret_cast<const double2 *>(s)); v[j] = cd{q.x, q.y}; }
                    else v[j] = cd{__ldg(s), __ldg(s + 1)};
                } else v[j] = border_pair_d<N>(p, ri.base, flags, ri.cval, cl);
            }
        }
        dftd<false, R1>(v);                                               // over j -> k1
        twiddle_store_d<R1, false>(v, s_tw + t, T, sb + t, T + 1);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int i = 0; i < T; i++) v[m * T + i] = ld_cd(sb + (t + T * m) * (T + 1) + i);
        __syncwarp();
        // the next row's samples are requested into L2 now (after this row's exchange, as in fast::row_fwd): 4 L reals = 32 T doubles per row,
        // one 128-byte line per lane and half
        if (wi + wstep < nwarp_items && nxt.active && !nxt.beyond && !nxt.zero && !nxt.has_const && p.xstr[al] == 1) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int64_t cc = nxt.cl0 - p.pf[al] + 16 * (t + T * h);
                if (cc >= 0 && cc < p.n[al]) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + nxt.base + cc));
            }
        }
#pragma unroll
        for (int m = 0; m < M; m++) dftd<false, T>(v + m * T);            // v[m*T + k2] = Z[k], k = t + T m + 16 k2
        // partner Z[L-k] and R2C post-processing; primaries are k2 < T/2 (k < L/2):
        //   X[k] = E + w^k O,  X[L-k] = conj(E - w^k O),  2E = Z[k] + conj Z[L-k],  w^k O = (-i/2) w^k (Z[k] - conj Z[L-k])
#pragma unroll
        for (int m = 0; m < M; m++) {
#pragma unroll
            for (int k2 = 0; k2 < T / 2; k2++) {
                const int ia = (M - 1 - m) * T + (T - 1 - k2);                                    // lanes t > 0
                const int ib = ((M - m) % M) * T + (m > 0 ? T - 1 - k2 : (k2 > 0 ? T - k2 : T / 2)); // lane t == 0 (own registers)
                double px = shfl_d(v[ia].re, src_lane);
                double py = shfl_d(v[ia].im, src_lane);
                if (t == 0) { px = v[ib].re; py = v[ib].im; }
                const cd zk = v[m * T + k2];
                const int k = t + T * m + R1 * k2;
                cd o0, o1;
                if (k == 0) {
                    o0 = cd{zk.re + zk.im, 0.0};                                                   // X[0]
                    o1 = cd{px, -py};                                                              // X[L/2] = conj Z[L/2]
                    if (ri.active) { st_cd(ri.dst + L, cd{zk.re - zk.im, 0.0}); st_cd(ri.dst + L + 1, cd{0.0, 0.0}); }   // Nyquist X[L]
                } else {
                    const cd cp = cd{px, -py};
                    const cd e2 = cadd(zk, cp), d = csub(zk, cp);
                    const cd tw = cmul(d, ld_cd(s_twr + k));
                    o0 = cd{0.5 * e2.re + tw.re, 0.5 * e2.im + tw.im};
                    o1 = cd{0.5 * e2.re - tw.re, -(0.5 * e2.im - tw.im)};
                }
                if (ri.active) { st_cd(ri.dst + 2 * k, o0); st_cd(ri.dst + 2 * k + 1, o1); }
            }
        }
        if (ri.active && t > 0 && t < 4) { st_cd(ri.dst + L + 2 * t, cd{0.0, 0.0}); st_cd(ri.dst + L + 2 * t + 1, cd{0.0, 0.0}); }
    }
}

// ---- row inverse + crop + decimate -------------------------------------------------------------------------------------------
struct RowInvInfoD {
    const cd *src;
    int64_t orow;      // output element offset of the row
    int tl;
    bool active;
};
template <int N> __device__ __forceinline__ RowInvInfoD resolve_inv_row_d(const RowParamsD &p, int64_t w, int pitch)
{
    RowInvInfoD r; r.active = w < p.nwork; r.src = p.ws; r.orow = 0; r.tl = 0;
    if (!r.active) return r;
    const uint32_t ntl = (uint32_t)p.ntiles[N - 1];
    const uint32_t w32 = (uint32_t)w;
    r.tl = (int)(w32 % ntl);
    uint32_t orow = w32 / ntl;
    uint32_t o[2] = {0, 0};
#pragma unroll
    for (int a = N - 2; a >= 0; a--) { o[a] = orow % (uint32_t)p.O[a]; orow /= (uint32_t)p.O[a]; }
    int64_t tile = orow, row = 0, obase = orow;  // what is left of the row index is the batch index (0 without a batch)
#pragma unroll
    for (int a = 0; a < N - 1; a++) {
        const uint32_t q = o[a] * (uint32_t)p.s[a];
        const uint32_t ta = q / (uint32_t)p.V[a];
        const uint32_t ra = q - ta * (uint32_t)p.V[a] + (uint32_t)p.Kd[a] - 1;
        tile = tile * p.ntiles[a] + ta;
        row = row * p.F[a] + ra;
        obase = obase * p.O[a] + o[a];
    }
    tile = tile * ntl + r.tl;
    r.src = p.ws + tile * p.tile_elems + row * pitch;
    r.orow = obase * p.O[N - 1];
    return r;
}

template <int T, int N>
__global__ void __launch_bounds__(128, 4) row_inv_d(const __grid_constant__ RowParamsD p)
{
    pdl_launch_dependents();
    constexpr int WPB = 4;
    constexpr int L = RowCfgD<T>::L, M = RowCfgD<T>::M, G = RowCfgD<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *s_tw = reinterpret_cast<cd *>(smem_raw);          // TRANSPOSED for the inverse flow: s_tw[i * 16 + k1] = W_L^{i k1}
    cd *s_twr = s_tw + L;                                 // i conj(w^k), k < L/2
    cd *s_ex = s_twr + L / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) st_cd(s_tw + idx, ld_cd(p.tw + (idx / R1) * (idx % R1)));
    for (int idx = threadIdx.x; idx < L / 2; idx += blockDim.x) { const cd w = ld_cd(p.twr + idx); st_cd(s_twr + idx, cd{w.im, w.re}); }
    __syncthreads();
    pdl_wait();
    cd *sb = s_ex + warp * RowCfgD<T>::wstride + g * RowCfgD<T>::gstride;
    cd *sp = s_ex + warp * RowCfgD<T>::wstride + g * (L + 2);           // staging: L/2 pairs (2 complex each) + the Nyquist pair, linear
    constexpr int al = N - 1;
    const int src_lane = g * T + ((T - t) % T);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    const int64_t wstep = (int64_t)gridDim.x * WPB;
    int64_t wi = (int64_t)blockIdx.x * WPB + warp;
    if (wi >= nwarp_items) return;
    RowInvInfoD ri = resolve_inv_row_d<N>(p, wi * G + g, L + kPadD);
    auto stage = [&](const RowInvInfoD &r) {
        if (r.active) {
            // L complex = L 16-byte chunks: T lanes x 16
#pragma unroll
            for (int q = 0; q < R1; q++) cp_async16(sp + t + T * q, r.src + t + T * q);
            if (t < 2) cp_async16(sp + L + t, r.src + L + t);                // column L: the Nyquist bin (and the zero beside it)
        }
        cp_async_commit();
    };
    stage(ri);
    for (; wi < nwarp_items; wi += wstep) {
        cd v[R1], b[R1 / 2];
        cp_async_wait_all();
        __syncwarp();
        const double nyq = ld_cd(sp + L).re;
        // C2R pre-processing (x2): Z[k] = E + O', Z[L-k] = conj(E - O'), E = Y[k] + conj Y[L-k], O' = i conj(w^k) (Y[k] - conj Y[L-k])
#pragma unroll
        for (int m = 0; m < M; m++) {
#pragma unroll
            for (int k2 = 0; k2 < T / 2; k2++) {
                const int k = t + T * m + R1 * k2;
                const cd yk = ld_cd(sp + 2 * k), ym = ld_cd(sp + 2 * k + 1);
                cd zk, zp;
                if (k == 0) {
                    zk = cd{yk.re + nyq, yk.re - nyq};                         // Z[0] from the real DC / Nyquist bins
                    zp = cd{2.0 * ym.re, -2.0 * ym.im};                        // Z[L/2] = 2 conj Y[L/2]
                } else {
                    const cd cm = cconj(ym);
                    const cd E = cadd(yk, cm), D = csub(yk, cm);
                    const cd O = cmul(D, ld_cd(s_twr + k));
                    zk = cadd(E, O);
                    zp = cconj(csub(E, O));
                }
                v[m * T + k2] = zk;
                b[m * (T / 2) + k2] = zp;
            }
        }
        __syncwarp();
        // deliver Z[L-k] to its owner: register (mr, k2r >= T/2) of lane t comes from lane (T-t)%T
#pragma unroll
        for (int mr = 0; mr < M; mr++) {
#pragma unroll
            for (int k2r = T / 2; k2r < T; k2r++) {
                const int ia = (M - 1 - mr) * (T / 2) + (T - 1 - k2r);                                            // source lanes t' > 0
                const int ib = mr > 0 ? (M - mr) * (T / 2) + (T - 1 - k2r) : (k2r == T / 2 ? 0 : (T - k2r));      // lane 0: own b[]
                double px = shfl_d(b[ia].re, src_lane);
                double py = shfl_d(b[ia].im, src_lane);
                if (t == 0) { px = b[ib].re; py = b[ib].im; }
                v[mr * T + k2r] = cd{px, py};
            }
        }
        // inverse: radix T over k2, conj twiddle, exchange, radix 16 over k1
#pragma unroll
        for (int m = 0; m < M; m++) dftd<true, T>(v + m * T);
#pragma unroll
        for (int m = 0; m < M; m++)
            twiddle_store_d<T, true>(v + m * T, s_tw + (t + T * m), R1, sb + (t + T * m) * (T + 1), 1);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < R1; k1++) v[k1] = ld_cd(sb + k1 * (T + 1) + t);
        __syncwarp();
        const RowInvInfoD cur = ri;
        if (wi + wstep < nwarp_items) { ri = resolve_inv_row_d<N>(p, (wi + wstep) * G + g, L + kPadD); stage(ri); }
        dftd<true, R1>(v);                                               // v[j] = z[t + T j] = (y[2n], y[2n+1]), n = t + T j
        if (!cur.active) continue;
        const int Kd1 = p.Kd[al];
        const int64_t mbase = (int64_t)cur.tl * p.V[al];
        if (p.s[al] == 1) {
            double *outp = p.out + (cur.orow + mbase - (Kd1 - 1));                         // tile sample i -> outp[i] (fast::keep_range, kernels_fft_fast.cuh)
            const int64_t room = p.O[al] - (mbase - (Kd1 - 1));
            const int i0 = 2 * t;
            const fast::KeepRange kr = fast::keep_range<2 * T>(Kd1 - 1, room < (int64_t)(2 * L) ? (int)room : 2 * L, i0);
            const bool vec_ok = (reinterpret_cast<uintptr_t>(outp) & 15) == 0;
            double *o = outp + i0;
#pragma unroll
            for (int j = 0; j < R1; j++) {
                const bool ok0 = j >= kr.ja0 && j < kr.jb0, ok1 = j >= kr.ja1 && j < kr.jb1;
                if (vec_ok && ok0 && ok1) *reinterpret_cast<double2 *>(o + 2 * T * j) = make_double2(v[j].re, v[j].im);
                else {
                    if (ok0) o[2 * T * j] = v[j].re;
                    if (ok1) o[2 * T * j + 1] = v[j].im;
                }
            }
        } else {
            const uint32_t s1 = (uint32_t)(p.s[al] < 0x7fffffff ? p.s[al] : 0x7fffffff), O1 = (uint32_t)p.O[al];
#pragma unroll
            for (int j = 0; j < R1; j++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * (t + T * j) + h;
                    if (i < Kd1 - 1) continue;
                    const uint32_t q = (uint32_t)(mbase + i - (Kd1 - 1));
                    const uint32_t o = q / s1;
                    if (o * s1 != q) continue;
                    if (o < O1) p.out[cur.orow + o] = h ? v[j].im : v[j].re;
                }
            }
        }
    }
}

// ---- column pass: FWD / INV / FMI ---------------------------------------------------------------------------------------------
struct ColParamsD {
    cd *ws;
    const cd *kspec;        // kernel spectrum, same (outer, F, inner) geometry as one tile (FMI only)
    const cd *tw;           // exp(-2 pi i j / F)
    int mode;               // 0 forward, 1 inverse, 2 forward * kspec * inverse
    int64_t outer, inner;   // the tile is [outer][F][inner] complex, inner % 8 == 0
    int64_t tile_elems, ntiles_total;
    int64_t nwork;          // ntiles_total * outer * inner / 8
    int skip;               // modes 1 / 2: the first `skip` rows of the axis (the aliased head the crop discards) are not stored
};

template <int E, int Tc> struct ColCfgD {
    static constexpr int F = E * Tc, Mc = E / Tc;
    static constexpr int threads = Tc * 8;
    static constexpr int pitch = Tc * 8 + 8;                             // padded k1-row stride of the exchange buffer
    static constexpr int ex = (E * pitch > F * 8 ? E * pitch : F * 8);
    static constexpr int smem = ((Tc == E ? 1 : 2) * F + ex) * 16;       // forward table, transposed table for the inverse when Tc != E, exchange buffer
    static constexpr int min_blocks = 4;
};

template <int E, int Tc>
__global__ void __launch_bounds__(ColCfgD<E, Tc>::threads, ColCfgD<E, Tc>::min_blocks) col_pass_d(const __grid_constant__ ColParamsD p)
{
    pdl_launch_dependents();
    using C = ColCfgD<E, Tc>;
    constexpr int F = C::F, Mc = C::Mc, pitch = C::pitch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *s_tw = reinterpret_cast<cd *>(smem_raw);        // s_tw[k1 * Tc + i] = W_F^{i k1}, k1 < E, i < Tc   (forward: i is the thread index)
    cd *s_twT = s_tw + F;                               // s_twT[ii * E + k1] = W_F^{ii k1}   (inverse, Tc != E: k1 is the thread-dependent index)
    cd *S = s_tw + (Tc == E ? 1 : 2) * F;
    const int tid = threadIdx.x;
    const int c = tid & 7, i = tid >> 3;                // column of the block, thread index inside the column (< Tc)
    for (int idx = tid; idx < F; idx += C::threads) { st_cd(s_tw + idx, ld_cd(p.tw + (idx / Tc) * (idx % Tc))); if (Tc != E) st_cd(s_twT + idx, ld_cd(p.tw + (idx / E) * (idx % E))); }
    const uint32_t iblocks = (uint32_t)(p.inner / 8), outer = (uint32_t)p.outer;
    struct Item { int64_t off, rel; };           // element offset of the block's first column in ws ; the same inside one tile
    auto decode = [&](int64_t w) {
        const uint32_t w32 = (uint32_t)w, q = w32 / iblocks, ib = w32 - q * iblocks, tile = q / outer, o = q - tile * outer;
        Item it; it.rel = (int64_t)o * F * p.inner + ib * 8; it.off = (int64_t)tile * p.tile_elems + it.rel;
        return it;
    };
    constexpr int kChunks = (F * 8 + C::threads - 1) / C::threads;       // 16-byte chunks (one complex double) per thread
    auto prefetch = [&](const Item &it) {
        const cd *gn = p.ws + it.off;
#pragma unroll
        for (int m = 0; m < kChunks; m++) {
            const int id = tid + C::threads * m, row = id >> 3, part = id & 7;
            if (F * 8 % C::threads == 0 || id < F * 8) cp_async16(S + row * 8 + part, gn + (int64_t)row * p.inner + part);
        }
        cp_async_commit();
    };
    pdl_wait();
    Item nxt; nxt.off = 0; nxt.rel = 0;
    if ((int64_t)blockIdx.x < p.nwork) {
        nxt = decode(blockIdx.x);
        prefetch(nxt);
    }
    for (int64_t w = blockIdx.x; w < p.nwork; w += gridDim.x) {
        const Item cur = nxt;
        cd *gt = p.ws + cur.off + c;
        cd v[E];
        cp_async_wait_all();
        __syncthreads();
        if (p.mode != 1) {
            // ---- forward: rows i + Tc j -> radix E -> twiddle -> exchange -> radix Tc -> rows q = (i + Tc m) + E k2 ----
#pragma unroll
            for (int j = 0; j < E; j++) v[j] = ld_cd(S + (i + Tc * j) * 8 + c);
            __syncthreads();
            dftd<false, E>(v);
            twiddle_store_d<E, false>(v, s_tw + i, Tc, S + i * 8 + c, pitch);
            __syncthreads();
#pragma unroll
            for (int m = 0; m < Mc; m++)
#pragma unroll
                for (int ii = 0; ii < Tc; ii++) v[m * Tc + ii] = ld_cd(S + (i + Tc * m) * pitch + ii * 8 + c);
#pragma unroll
            for (int m = 0; m < Mc; m++) dftd<false, Tc>(v + m * Tc);
            if (p.mode == 2) {
                const cd *kp = p.kspec + cur.rel + c;                           // same offset inside the kernel spectrum tile
#pragma unroll
                for (int m = 0; m < Mc; m++)
#pragma unroll
                    for (int k2 = 0; k2 < Tc; k2++) v[m * Tc + k2] = cmul(v[m * Tc + k2], ld_cd(kp + (int64_t)(i + Tc * m + E * k2) * p.inner));
            }
        } else {
#pragma unroll
            for (int m = 0; m < Mc; m++)
#pragma unroll
                for (int k2 = 0; k2 < Tc; k2++) v[m * Tc + k2] = ld_cd(S + (i + Tc * m + E * k2) * 8 + c);
        }
        if (p.mode == 0) {
            __syncthreads();                               // all exchange reads done: S may be restaged
            if (w + gridDim.x < p.nwork) { nxt = decode(w + gridDim.x); prefetch(nxt); }
#pragma unroll
            for (int m = 0; m < Mc; m++)
#pragma unroll
                for (int k2 = 0; k2 < Tc; k2++) st_cd(gt + (int64_t)(i + Tc * m + E * k2) * p.inner, v[m * Tc + k2]);
            continue;
        }
        if constexpr (Tc == E) {
            // square case: the rows this thread holds (i + E k2) are also the rows the forward-structured flow starts from
            dftd<true, E>(v);
            __syncthreads();                               // every thread has finished reading S
            twiddle_store_d<E, true>(v, s_tw + i, Tc, S + i * 8 + c, pitch);
            __syncthreads();
#pragma unroll
            for (int ii = 0; ii < Tc; ii++) v[ii] = ld_cd(S + i * pitch + ii * 8 + c);
            __syncthreads();
        } else {
            // ---- inverse: radix Tc over k2 -> conj twiddle -> exchange -> radix E over k1 -> rows i + Tc j ----
#pragma unroll
            for (int m = 0; m < Mc; m++) dftd<true, Tc>(v + m * Tc);
            __syncthreads();                               // every thread has finished reading S
#pragma unroll
            for (int m = 0; m < Mc; m++)
                twiddle_store_d<Tc, true>(v + m * Tc, s_twT + (i + Tc * m), E, S + (i + Tc * m) * pitch + c, 8);
            __syncthreads();
#pragma unroll
            for (int k1 = 0; k1 < E; k1++) v[k1] = ld_cd(S + k1 * pitch + i * 8 + c);
            __syncthreads();
        }
        if (w + gridDim.x < p.nwork) { nxt = decode(w + gridDim.x); prefetch(nxt); }
        dftd<true, E>(v);
#pragma unroll
        for (int j = 0; j < E; j++) if (i + Tc * j >= p.skip) st_cd(gt + (int64_t)(i + Tc * j) * p.inner, v[j]);
    }
}

}  // namespace fast64

// kernel spectrum [rows][Hp] in natural bin order (bins 0..L, generic path) -> the f64 fast path's row layout (pitch L + 8), see KfastBody
struct KfastParamsD {
    const cx<double> *kspec;
    cx<double> *kfast;
    int64_t rows;
    int L, Hp;
};
struct KfastBodyD {
    static HD void run(const BlockCtx &c, const KfastParamsD &p)
    {
        const int pitch = p.L + fast64::kPadD;
        const int64_t total = p.rows * pitch;
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int64_t q = e / pitch;
            const int pc = (int)(e % pitch);
            cx<double> val = cx<double>{0.0, 0.0};
            if (pc <= p.L) {
                int bin;
                if (pc == p.L) bin = p.L;
                else if (pc == 0) bin = 0;
                else if (pc == 1) bin = p.L / 2;
                else bin = (pc & 1) ? p.L - (pc >> 1) : (pc >> 1);
                val = p.kspec[q * p.Hp + bin];
            }
            p.kfast[e] = val;
        }
    }
};
}  // namespace ndc
#endif  // NDCONV_CUDA
