// kernels_fft_opt.cuh -- sm_100a fast path of the 2-D real conv_fft pipeline (device only).
//
// Tile = F0 x F1 = 1024 x 2048 real samples (overlap-save in both axes).  A row of 2048 reals is packed as
// L = 1024 complex z[j] = x[2j] + i x[2j+1] (half-length real FFT).
//
//   row_fwd_packed   one warp per row: border-mapped (or plain vector) loads straight into registers, radix-32 x
//                    radix-32 Stockham with one warp-private shared-memory transpose; Z[L-k] is fetched from its owner
//                    lane by warp shuffle, the R2C post-processing X[k] = E + w^k O is done in registers, and the row
//                    is stored with 16-byte stores in the PAIRED layout: slot k in [1,512) holds (X[k], X[L-k]);
//                    slot 0 holds ((X[0], X[L]) packed as re/im -- both are real --, X[L/2]).
//   col_fmi          256 threads per 1024 x 8-column tile: radix-32 x radix-32 forward along the strided axis, multiply by
//                    the cached kernel spectrum (same physical layout), radix-32 x radix-32 inverse, in place.  Physical
//                    column 0 carries two real sequences (DC / Nyquist); it is separated through its q <-> -q mirror.
//   row_inv_packed   one warp per (output row, tile): paired loads, C2R pre-processing in registers, shuffle, inverse
//                    radix-32 x radix-32, crop [Kd-1, F) and stride decimation fused into the (vector) store.
//   kphys_repack     kernel spectrum [F0][L+1] (generic path, cached) -> the physical column order above.
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
    int F0;                 // column tile height (256 or 1024 rows)
    int64_t s[2], O[2];
    const float *x;
    float *out;
    cf *ws;                 // [tile][kF0][kL] paired layout
    const cf *tw;           // exp(-2 pi i j / 1024), j < 1024
    const cf *twr;          // exp(-2 pi i k / 2048), k <= 512
    int64_t nwork;
    // L2-resident batching: when batch != 0 the launch covers tiles (t0 = b_t0, t1 in [b_t1, b_t1 + b_nt1)) only and the
    // workspace holds just those tiles (tile slot = t1 - b_t1); row_inv covers output rows [b_o0, b_o0 + b_no0)
    int batch, b_t0, b_t1, b_nt1;
    int64_t b_o0, b_no0;
};

// s_tw[k1*32 + t] = W_1024^{t k1}
__device__ __forceinline__ void load_tw_table(cf *s_tw, const cf *tw, int tid, int nt)
{
    for (int idx = tid; idx < 1024; idx += nt) s_tw[idx] = tw[(idx >> 5) * (idx & 31)];
}

// ---- cp.async helpers (LDGSTS): global -> shared without staging in registers ---------------------------------
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- row forward -----------------------------------------------------------------------------------
struct RowFwdInfo {
    float4 *dst;
    const float *src;       // interior rows: first sample of the tile
    int64_t rowbase, cl0;
    float row_cval;
    bool zero_row, row_const, row_init, interior, vec;
};

__device__ __forceinline__ RowFwdInfo row_fwd_resolve(const RowOptParams &p, int64_t w)
{
    RowFwdInfo ri;
    int64_t tile = w / p.F0;
    const int r = (int)(w % p.F0);
    int t0, t1;
    if (p.batch) { t0 = p.b_t0; t1 = p.b_t1 + (int)tile; }
    else { t0 = (int)(tile / p.ntiles[1]); t1 = (int)(tile % p.ntiles[1]); }
    const int64_t c0 = (int64_t)t0 * p.V[0] + r;         // padded row
    ri.cl0 = (int64_t)t1 * p.V[1];                       // first padded column of the tile
    ri.dst = reinterpret_cast<float4 *>(p.ws + tile * ((int64_t)p.F0 * kL) + (int64_t)r * kL);
    ri.zero_row = false; ri.row_const = false; ri.row_init = false; ri.row_cval = 0.f; ri.rowbase = 0;
    if (c0 >= p.P[0]) ri.zero_row = true;
    else {
        const int32_t m0 = p.map[0][c0];
        if (m0 >= 0) ri.rowbase = (int64_t)m0 * p.xstr[0];
        else if (m0 == NDC_MAP_INIT) ri.row_init = true;
        else { ri.row_const = true; ri.row_cval = (m0 == NDC_MAP_CONST_FRONT) ? p.cfront[0] : p.cback[0]; }
    }
    ri.interior = !ri.zero_row && !ri.row_const && !ri.row_init && p.xstr[1] == 1 && ri.cl0 >= p.pf[1] && ri.cl0 + kF1 <= p.pf[1] + p.n[1];
    ri.src = p.x + ri.rowbase + (ri.cl0 - p.pf[1]);
    ri.vec = ri.interior && (reinterpret_cast<uintptr_t>(ri.src) & 7) == 0;
    return ri;
}

// Measured on B200 (c5): staging the next input row with 8-byte cp.async costs more issue slots than the latency it hides
// (1.97 -> 2.06 ms), while the 16-byte variant in row_inv_packed pays off (2.01 -> 1.83 ms).  Input rows are only 8-byte
// aligned in general ((cl0 - pf) * 4), so row_fwd keeps the direct register loads.
constexpr bool kRowFwdPrefetch = false;

__global__ void __launch_bounds__(128, 4) row_fwd_packed(const __grid_constant__ RowOptParams p)
{
    __shared__ cf s_tw[1024];
    __shared__ cf s_twr[512];
    __shared__ __align__(16) cf s_buf[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tw_table(s_tw, p.tw, threadIdx.x, blockDim.x);
    for (int idx = threadIdx.x; idx < 512; idx += blockDim.x) s_twr[idx] = p.twr[idx];
    __syncthreads();
    cf *sb = s_buf[warp];
    const int src_lane = (32 - lane) & 31;
    const int64_t wstride = (int64_t)gridDim.x * 4;
    int64_t w = (int64_t)blockIdx.x * 4 + warp;
    if (w >= p.nwork) return;
    RowFwdInfo cur = row_fwd_resolve(p, w);
    // software pipeline: the NEXT row's 8 KB are copied global -> shared (cp.async, no registers) while this row computes
    bool staged = false;
    if (kRowFwdPrefetch && cur.vec) {
#pragma unroll
        for (int j = 0; j < 32; j++) cp_async8(sb + lane + 32 * j, reinterpret_cast<const float2 *>(cur.src) + lane + 32 * j);
        cp_async_commit();
        staged = true;
    }
    for (; w < p.nwork; w += wstride) {
        const bool has_next = kRowFwdPrefetch && (w + wstride < p.nwork);
        RowFwdInfo nxt = cur;
        if (has_next) nxt = row_fwd_resolve(p, w + wstride);
        if (cur.zero_row) {
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) cur.dst[lane + 32 * k2] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kRowFwdPrefetch && has_next && nxt.vec) {          // nothing staged for a zero row: the buffer is free
#pragma unroll
                for (int j = 0; j < 32; j++) cp_async8(sb + lane + 32 * j, reinterpret_cast<const float2 *>(nxt.src) + lane + 32 * j);
                cp_async_commit();
                staged = true;
            }
            if (kRowFwdPrefetch) cur = nxt; else if (w + wstride < p.nwork) cur = row_fwd_resolve(p, w + wstride);
            continue;
        }
        cf v[32];
        if (staged) {
            cp_async_wait_all();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = sb[lane + 32 * j];
            __syncwarp();
            staged = false;
        } else if (cur.vec) {
            const float2 *s2 = reinterpret_cast<const float2 *>(cur.src);
#pragma unroll
            for (int j = 0; j < 32; j++) { float2 t = __ldg(s2 + lane + 32 * j); v[j] = cf{t.x, t.y}; }
        } else if (cur.interior) {
#pragma unroll
            for (int j = 0; j < 32; j++) { const int e = 2 * (lane + 32 * j); v[j] = cf{__ldg(cur.src + e), __ldg(cur.src + e + 1)}; }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float q[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int64_t cl = cur.cl0 + 2 * (lane + 32 * j) + h;
                    float val = 0.f;
                    if (cl < p.P[1]) {
                        const int32_t m = p.map[1][cl];
                        if (m == NDC_MAP_CONST_FRONT) val = p.cfront[1];
                        else if (m == NDC_MAP_CONST_BACK) val = p.cback[1];
                        else if (cur.row_const) val = cur.row_cval;
                        else if (m == NDC_MAP_INIT || cur.row_init) val = 0.f;
                        else val = __ldg(p.x + cur.rowbase + (int64_t)m * p.xstr[1]);
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
        if (kRowFwdPrefetch && has_next && nxt.vec) {            // the buffer is free again: prefetch the next row
#pragma unroll
            for (int j = 0; j < 32; j++) cp_async8(sb + lane + 32 * j, reinterpret_cast<const float2 *>(nxt.src) + lane + 32 * j);
            cp_async_commit();
            staged = true;
        }
        dft32<float>(v, false);                                  // over t -> k2: v[k2] = Z[lane + 32 k2]
        // Z[L-k] lives in lane (32-lane)&31, register 31-k2 (lane 0: register 32-k2; k = 0 pairs with L/2).
        // R2C post-processing: E = (Z[k] + conj Z[L-k])/2, O = -i (Z[k] - conj Z[L-k])/2, X[k] = E + w^k O, X[L-k] = conj(E - w^k O)
        float4 *dst = cur.dst;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            float px = __shfl_sync(0xffffffffu, v[31 - k2].re, src_lane);
            float py = __shfl_sync(0xffffffffu, v[31 - k2].im, src_lane);
            if (lane == 0) { const cf o = (k2 == 0) ? v[16] : v[(32 - k2) & 31]; px = o.re; py = o.im; }
            const cf zk = v[k2];
            float4 o4;
            if (lane == 0 && k2 == 0) {
                o4 = make_float4(zk.re + zk.im, zk.re - zk.im, px, -py);          // (X[0], X[L]) packed ; X[L/2] = conj Z[L/2]
            } else {
                const cf E = cf{0.5f * (zk.re + px), 0.5f * (zk.im - py)};
                const cf O = cf{0.5f * (zk.im + py), -0.5f * (zk.re - px)};
                const cf t = cmul(s_twr[lane + 32 * k2], O);
                o4 = make_float4(E.re + t.re, E.im + t.im, E.re - t.re, -(E.im - t.im));
            }
            dst[lane + 32 * k2] = o4;
        }
        if (kRowFwdPrefetch) cur = nxt; else if (w + wstride < p.nwork) cur = row_fwd_resolve(p, w + wstride);
    }
}

// ---- row inverse + crop + decimate ---------------------------------------------------------------------
__device__ __forceinline__ const float4 *row_inv_src(const RowOptParams &p, int64_t w, int &t1, int64_t &o0)
{
    const int ntl = p.ntiles[1];
    if (p.batch) { t1 = p.b_t1 + (int)(w % p.b_nt1); o0 = p.b_o0 + w / p.b_nt1; }
    else { t1 = (int)(w % ntl); o0 = w / ntl; }
    const int64_t q0 = o0 * p.s[0];
    const int64_t t0 = q0 / p.V[0];
    const int r0 = (int)(q0 - t0 * p.V[0]) + p.Kd[0] - 1;
    const int64_t tile = p.batch ? (int64_t)(t1 - p.b_t1) : t0 * ntl + t1;
    return reinterpret_cast<const float4 *>(p.ws + tile * ((int64_t)p.F0 * kL) + (int64_t)r0 * kL);
}

__global__ void __launch_bounds__(128, 4) row_inv_packed(const __grid_constant__ RowOptParams p)
{
    __shared__ cf s_tw[1024];
    __shared__ cf s_twr[512];
    __shared__ __align__(16) cf s_buf[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tw_table(s_tw, p.tw, threadIdx.x, blockDim.x);
    for (int idx = threadIdx.x; idx < 512; idx += blockDim.x) s_twr[idx] = p.twr[idx];
    __syncthreads();
    cf *sb = s_buf[warp];
    float4 *sb4 = reinterpret_cast<float4 *>(sb);
    const int src_lane = (32 - lane) & 31;
    const int64_t wstride = (int64_t)gridDim.x * 4;
    int64_t w = (int64_t)blockIdx.x * 4 + warp;
    if (w >= p.nwork) return;
    int t1; int64_t o0;
    const float4 *src = row_inv_src(p, w, t1, o0);
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) cp_async16(sb4 + lane + 32 * k2, src + lane + 32 * k2);
    cp_async_commit();
    for (; w < p.nwork; w += wstride) {
        cf v[32], b[16];
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const float4 t = sb4[lane + 32 * k2];
            // C2R pre-processing (x2): Z[k] = E + i O, Z[L-k] = conj(E) + i conj(O), E = Y[k] + conj Y[L-k], O = conj(w^k) (Y[k] - conj Y[L-k])
            if (lane == 0 && k2 == 0) {
                v[0] = cf{t.x + t.y, t.x - t.y};                                   // slot 0 = ((Y[0], Y[L]) packed, Y[L/2])
                b[0] = cf{2.f * t.z, -2.f * t.w};                                  // Z[L/2] = 2 conj Y[L/2]
            } else {
                const cf E = cf{t.x + t.z, t.y - t.w};
                const cf D = cf{t.x - t.z, t.y + t.w};
                const cf O = cmulc(D, s_twr[lane + 32 * k2]);
                v[k2] = cf{E.re - O.im, E.im + O.re};
                b[k2] = cf{E.re + O.im, -E.im + O.re};
            }
        }
        __syncwarp();
        // v[j'] for j' >= 16 is Z[lane + 32 j'] = the partner half held by lane (32-lane)&31 at index 31-j'
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
        // crop parameters of THIS row, then prefetch the next row's spectrum into the (free again) buffer
        const int cur_t1 = t1;
        const int64_t cur_o0 = o0;
        if (w + wstride < p.nwork) {
            src = row_inv_src(p, w + wstride, t1, o0);
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) cp_async16(sb4 + lane + 32 * k2, src + lane + 32 * k2);
            cp_async_commit();
        }
        dft32<float>(v, true);                                   // v[n2] = z[lane + 32 n2] = (y[2n], y[2n+1])
        // crop [Kd-1, F1) of the tile; global position m = t1*V + i; keep m < P and (m-Kd+1) % s == 0
        const int Kd1 = p.Kd[1];
        const int64_t orow = cur_o0 * p.O[1];
        const int64_t mbase = (int64_t)cur_t1 * p.V[1];
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
constexpr int kKphysPitch = kL + 8;      // kernel spectrum in physical column order, [kF0][kKphysPitch]; column kL holds K[q][L] (Nyquist)

struct ColOptParams {
    cf *ws;
    const cf *kphys;
    const cf *tw;           // exp(-2 pi i j / 1024)
    int64_t ntiles_total;
    int64_t nwork;          // ntiles_total * (kL / 8)
};

// R = 32: 1024-row tiles (256 threads); R = 16: 256-row tiles (128 threads).  F0 = R*R, radix R x radix R.
template <int R> struct ColCfg {
    static constexpr int F0 = R * R;
    static constexpr int threads = R * 8;
    static constexpr int pitch = R * 8 + 8;                         // padded k1-row stride of the exchange buffer (complex elements)
    static constexpr int smem = (F0 + R * pitch) * 8;
};

template <int R>
__global__ void __launch_bounds__(ColCfg<R>::threads, R == 32 ? 2 : 6) col_fmi(const __grid_constant__ ColOptParams p)
{
    constexpr int F0 = ColCfg<R>::F0, pitch = ColCfg<R>::pitch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *s_tw = reinterpret_cast<cf *>(smem_raw);        // F0 entries: s_tw[k1*R + i] = W_F0^{i k1}
    cf *S = s_tw + F0;                                  // R * pitch
    const int tid = threadIdx.x;
    const int c = tid & 7, i = tid >> 3;
    for (int idx = tid; idx < F0; idx += blockDim.x) s_tw[idx] = p.tw[(idx / R) * (idx % R)];
    // software pipeline: while the last inverse pass and the stores of one tile run, the exchange buffer is idle, so the next
    // tile is copied into it (linear [F0][8]) with 16-byte cp.async -- 4 threads per 64-byte row segment, no registers held
    constexpr int kChunks = F0 * 4 / ColCfg<R>::threads;       // 16-byte chunks per thread
    auto prefetch = [&](int64_t wn) {
        const int cbn = (int)(wn % (kL / 8));
        const cf *gn = p.ws + (wn / (kL / 8)) * ((int64_t)F0 * kL) + cbn * 8;
#pragma unroll
        for (int m = 0; m < kChunks; m++) {
            const int id = tid + ColCfg<R>::threads * m, row = id >> 2, part = id & 3;
            cp_async16(S + row * 8 + part * 2, gn + (int64_t)row * kL + part * 2);
        }
        cp_async_commit();
    };
    if ((int64_t)blockIdx.x < p.nwork) prefetch(blockIdx.x);
    for (int64_t w = blockIdx.x; w < p.nwork; w += gridDim.x) {
        const int cb = (int)(w % (kL / 8));
        const int64_t tile = w / (kL / 8);
        cf *g = p.ws + tile * ((int64_t)F0 * kL) + cb * 8 + c;
        cf v[R];
        cp_async_wait_all();
        __syncthreads();                                   // the staged tile (and, first time, the twiddle table) is visible
#pragma unroll
        for (int j = 0; j < R; j++) v[j] = S[(i + R * j) * 8 + c];
        __syncthreads();                                   // every thread has its rows: S becomes the exchange buffer
        // ---- forward: radix R over j, twiddle, exchange, radix R over i ----
        dft<float, R>(v, false);
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) S[k1 * pitch + i * 8 + c] = cmul(v[k1], s_tw[k1 * R + i]);
        __syncthreads();
#pragma unroll
        for (int ii = 0; ii < R; ii++) v[ii] = S[i * pitch + ii * 8 + c];
        dft<float, R>(v, false);                           // v[k2] = Xhat[q = i + R k2][column]
        // ---- multiply by the kernel spectrum (rows q = i + R k2 stay in this thread for the inverse) ----
        const cf *kp = p.kphys + cb * 8 + c;
        if (cb == 0) {
            // physical column 0 = DC + i Nyquist of every row: separate the two real sequences through the q <-> -q mirror
            __syncthreads();
            if (c == 0) {
#pragma unroll
                for (int k2 = 0; k2 < R; k2++) S[i + R * k2] = v[k2];
            }
            __syncthreads();
            if (c == 0) {
#pragma unroll
                for (int k2 = 0; k2 < R; k2++) {
                    const int q = i + R * k2, qm = (F0 - q) & (F0 - 1);
                    const cf a = v[k2], b = S[qm];
                    const cf A = cf{0.5f * (a.re + b.re), 0.5f * (a.im - b.im)};      // spectrum of the DC column
                    const cf B = cf{0.5f * (a.im + b.im), -0.5f * (a.re - b.re)};     // spectrum of the Nyquist column
                    const cf Yd = cmul(A, ld_cf(kp + (int64_t)q * kKphysPitch));
                    const cf Yn = cmul(B, ld_cf(p.kphys + (int64_t)q * kKphysPitch + kL));
                    v[k2] = cf{Yd.re - Yn.im, Yd.im + Yn.re};                         // Yd + i Yn
                }
            } else {
#pragma unroll
                for (int k2 = 0; k2 < R; k2++) v[k2] = cmul(v[k2], ld_cf(kp + (int64_t)(i + R * k2) * kKphysPitch));
            }
        } else {
#pragma unroll
            for (int k2 = 0; k2 < R; k2++) v[k2] = cmul(v[k2], ld_cf(kp + (int64_t)(i + R * k2) * kKphysPitch));
        }
        // ---- inverse: radix R over the register index, conj twiddle, exchange, radix R over i ----
        dft<float, R>(v, true);
        __syncthreads();                                   // every thread has finished reading S
#pragma unroll
        for (int n1 = 0; n1 < R; n1++) S[n1 * pitch + i * 8 + c] = cmulc(v[n1], s_tw[n1 * R + i]);
        __syncthreads();
#pragma unroll
        for (int ii = 0; ii < R; ii++) v[ii] = S[i * pitch + ii * 8 + c];
        __syncthreads();                                   // S is free: stage the next tile while this one finishes
        if (w + gridDim.x < p.nwork) prefetch(w + gridDim.x);
        dft<float, R>(v, true);
#pragma unroll
        for (int n2 = 0; n2 < R; n2++) st_cf(g + (int64_t)(i + R * n2) * kL, v[n2]);
    }
}

}  // namespace opt

// kernel spectrum [F0][Hp] (bins 0..L) -> physical column order of the paired layout:
// column 0 <- bin 0, column 1 <- bin L/2, columns (2s, 2s+1) <- bins (s, L-s), column L <- bin L (Nyquist).
struct KphysParams {
    const cx<float> *kspec;
    cx<float> *kphys;
    int F0, L, Hp, pitch;
};
struct KphysBody {
    static HD void run(const BlockCtx &c, const KphysParams &p)
    {
        const int64_t total = (int64_t)p.F0 * (p.L + 1);
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int q = (int)(e / (p.L + 1)), pc = (int)(e % (p.L + 1));
            int bin;
            if (pc == p.L) bin = p.L;
            else if (pc == 0) bin = 0;
            else if (pc == 1) bin = p.L / 2;
            else bin = (pc & 1) ? p.L - (pc >> 1) : (pc >> 1);
            p.kphys[(int64_t)q * p.pitch + pc] = p.kspec[(int64_t)q * p.Hp + bin];
        }
    }
};
}  // namespace ndc
#endif  // NDCONV_CUDA
