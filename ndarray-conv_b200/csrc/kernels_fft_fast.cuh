// kernels_fft_fast.cuh -- sm_100a fast path of the real f32 conv_fft pipeline, rank 2 and 3 (device only).
//
// Overlap-save tiles of F0 [x F1] x F_last real samples.  The last axis is a half-length real transform: F_last = 2L reals
// are packed as L = 32*T complex (T = 4, 8, 16, 32 lanes per row, 32 complex per lane).  Other axes: F = E*Tc, E = 8, 16, 32
// rows per thread, Tc = E, E/2 or E/4 threads per column (F = 16 .. 1024).
//
// Workspace layout per tile: rows of PITCH = L + 8 complex.  Columns [0, L) hold the half spectrum in PAIRED order -- slot k
// in [1, L/2) = (X[k], X[L-k]), slot 0 = (X[0], X[L/2]) -- so a row kernel emits 16-byte stores and the partner bins needed by
// the real-transform algebra are adjacent; column L holds the Nyquist bin X[L]; columns L+1 .. L+7 are zero.  Every column is an
// ordinary complex column for the other axes, which is what lets the same kernels serve rank 3.
//
//   row_fwd<T>      one lane-group of T lanes per row (32/T rows per warp, no block barrier): border-mapped or plain vector
//                   loads into registers, radix-32 over the register index, twiddle, warp-private shared-memory exchange,
//                   radix-T; Z[L-k] comes from its owner lane by __shfl_sync; R2C post-processing X[k] = E + w^k O in registers.
//   col_pass<E,Tc>  Tc*8 threads per F x 8-column tile (64-byte row segments): radix-E, twiddle, padded exchange, radix-Tc;
//                   mode FWD / INV / FMI (forward, multiply by the cached kernel spectrum, inverse -- in place).  The next tile is
//                   staged into the idle exchange buffer with 16-byte cp.async while the last pass and the stores run.
//   row_inv<T>      one lane-group per (output row, last-axis tile): cp.async-prefetched paired loads, C2R pre-processing in
//                   registers, shuffle, inverse radix-T, exchange, inverse radix-32, crop [Kd-1, F) and stride fused into the store.
//
// Reference stages replaced: conv_fft/padding.rs:30-62, processor/real.rs:105-154, mod.rs:268, real.rs:233-280, mod.rs:282-289.
#pragma once
#include "kernels_fft.h"
#include "packed_cf.cuh"

#ifdef NDCONV_CUDA
#include <cuda.h>          // CUtensorMap (col_pass_tma)
#include <type_traits>
namespace ndc {
namespace fast {

// Programmatic dependent launch: every kernel of the fast path lets the next kernel of the stream start early
// (launch_dependents at the top) and itself waits for its predecessors only after its prologue -- the twiddle tables staged
// into shared memory come from per-processor constants no kernel writes -- so launch latency and prologue of kernel n+1 overlap
// the tail of kernel n.  Nothing a predecessor writes (workspace, output) or reads (a workspace the successor overwrites) is
// touched before pdl_wait().  Launched without the attribute (NDCONV_DISABLE_PDL) both are no-ops.
DEV void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Experiments only (tools/exp/build_ring_lib.sh builds a second library with -DNDCONV_EXP_RING): every tile's spectra land in slot
// tile % ring of the workspace, so the workspace traffic of each pass stays in L2 -- RESULTS ARE GARBAGE; the timings bound what a
// fused, L2-resident pipeline could reach.  The product library is built without the macro.
#ifdef NDCONV_EXP_RING
__device__ int g_exp_ring = 1 << 30;
__device__ int g_exp_flags = 0;            // col_pass_tma_kres: 1 = no workspace loads (compute on what the landing buffer holds), 2 = no stores
#define NDC_RING_TILE(t) ((t) % g_exp_ring)
#define NDC_EXP_FLAG(b) ((g_exp_flags & (b)) != 0)
#else
#define NDC_RING_TILE(t) (t)
#define NDC_EXP_FLAG(b) false
#endif

typedef cx<float> cf;
typedef pk::pcf pc;              // the same 8 bytes as cf, held as one 64-bit register pair for FADD2 / FMUL2 / FFMA2 (packed_cf.cuh)
constexpr int kPad = 8;          // extra columns per row: Nyquist + 7 zeros (keeps rows 64-byte aligned; 128-byte rows measured no faster)

__device__ __forceinline__ pc ld_pc(const cf *p) { pc r; r.v = *reinterpret_cast<const unsigned long long *>(p); return r; }
__device__ __forceinline__ void st_pc(cf *p, pc v) { *reinterpret_cast<unsigned long long *>(p) = v.v; }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// The kernel spectrum (one tile, a few MB) is re-read by every tile while ~10 GB of workspace stream through L2: its loads carry an
// evict_last policy so the stream does not push it out (c5: col_fwd_mul_inv 3.08 -> 2.79 ms; DRAM re-reads of the spectrum gone).
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ pc ld_pc_keep(const cf *p, uint64_t pol)
{
    pc r; asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r.v) : "l"(p), "l"(pol)); return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Crop of a tile row at the store: tile samples [lo, hi) are output elements.  A lane holds samples i0 + STEP j and i0 + 1 + STEP j
// (STEP = 2 T, a power of two); the kept ones are the j of one interval per half -- [ja0, jb0) and [ja1, jb1) -- worked out once in 32-bit
// arithmetic, so a store costs two compares against constants and an immediate offset (the per-store 64-bit index arithmetic and bound
// checks were ~30 % of row_inv's executed instructions on c5).
struct KeepRange { int ja0, jb0, ja1, jb1; };
template <int STEP> __device__ __forceinline__ KeepRange keep_range(int lo, int hi, int i0)
{
    static_assert(STEP > 0 && (STEP & (STEP - 1)) == 0, "power of two");
    constexpr int SH = STEP == 128 ? 7 : STEP == 64 ? 6 : STEP == 32 ? 5 : STEP == 16 ? 4 : STEP == 8 ? 3 : STEP == 4 ? 2 : 1;
    static_assert((1 << SH) == STEP, "2 <= STEP <= 128");
    KeepRange k;                                                                   // (x + STEP - 1) >> SH = ceil(x / STEP) for either sign
    k.ja0 = (lo - i0 + STEP - 1) >> SH; k.jb0 = (hi - i0 + STEP - 1) >> SH;
    k.ja1 = (lo - i0 + STEP - 2) >> SH; k.jb1 = (hi - i0 + STEP - 2) >> SH;
    return k;
}

// dst[k * dst_stride] = v[k] * tw[k * tw_stride] (CONJ: times the conjugate), k < N, both in shared memory.  The twiddles are fetched B at a
// time BEFORE the stores that follow them: the compiler cannot move a shared-memory load above a shared-memory store it cannot
// disambiguate, and one LDS -> multiply -> STS chain per k left every twiddle's load latency exposed (ncu source page of c5, round 2e:
// the first consumer of each twiddle held 12 % of row_fwd's and 10 % of row_inv's stall samples).
template <int N, bool CONJ, int B = 8>
__device__ __forceinline__ void twiddle_store(const pc *v, const pc *tw, int tw_stride, pc *dst, int dst_stride)
{
#pragma unroll
    for (int k0 = 0; k0 < N; k0 += B) {
        pc w[B];
#pragma unroll
        for (int e = 0; e < B; e++) if (k0 + e < N) w[e] = tw[(k0 + e) * tw_stride];
#pragma unroll
        for (int e = 0; e < B; e++) if (k0 + e < N) dst[(k0 + e) * dst_stride] = CONJ ? pk::cmulc(v[k0 + e], w[e]) : pk::cmul(v[k0 + e], w[e]);
    }
}

struct RowParams {
    int ndim;                                  // 2 or 3
    int64_t n[3], xstr[3], P[3], pf[3];
    const int32_t *map[3];
    float cfront[3], cback[3];
    float cfront_im[3], cback_im[3];           // Complex<f32> problems: imaginary parts of the constant borders
    int F[3], V[3], ntiles[3], Kd[3];          // F[ndim-1] = 2L
    int64_t s[3], O[3];
    const float *x;
    float *out;
    cf *ws;
    const cf *tw;                              // exp(-2 pi i j / L), j < L
    const cf *twr;                             // exp(-2 pi i k / 2L), k <= L/2
    const cf *kfast;                           // rank 1 only: the kernel spectrum of one tile in the workspace row layout
    int64_t nwork;                             // rows to transform (fwd) / (output row, tile) pairs (inv) / tiles (rank 1)
    int64_t rows_per_tile, tile_elems;         // prod of the outer F ; rows_per_tile * (L + 8)
    // same-shape batch (a leading axis of kernel extent 1 folded away by the host): problem b reads x + b * xstr_batch, its tiles follow
    // those of problem b - 1 in the workspace and its output rows follow in `out`; the batch index is the outermost tile coordinate
    int64_t xstr_batch;
    FastDiv fd_rpt, fd_nt[3], fd_F[3];         // row_fwd: exact division by rows_per_tile / ntiles[a] / F[a] as multiply + shift (four run-time divisions per row were ~80 instructions)
    int pf_mode;                               // row_fwd, experiments: where the next row's samples are requested into L2 (0 never, 1 before this row's loads, 2 after its exchange)
};

// ---- row forward ---------------------------------------------------------------------------------------------------------
template <int T> struct RowCfg {
    static constexpr int L = 32 * T, M = 32 / T, G = 32 / T;          // complex length, radix-T butterflies per lane, rows per warp
    static constexpr int gstride = 32 * (T + 1) + T;                  // exchange buffer of one lane-group (complex elements)
    static constexpr int wstride = (G * gstride > (L + 2) * G ? G * gstride : (L + 2) * G);   // staging: L/2 + 1 float4 per group
    static constexpr int smem = (L + L / 2 + 4 * wstride) * 8;        // W_L table, w^k table, 4 warps of exchange buffers
};

struct RowSrcInfo {
    int64_t base, cl0;
    float cval, cval_im;
    bool zero, has_const, beyond, active;
    cf *dst;
};

// outer-axis resolution of one tile row: beyond the padded extent of an outer axis the FFT buffer is zero
// (conv_fft/padding.rs:47-59); otherwise the highest-numbered constant axis wins and never-written cells read 0
template <int N> __device__ __forceinline__ RowSrcInfo resolve_fwd_row(const RowParams &p, int64_t w, int pitch)
{
    RowSrcInfo r; r.base = 0; r.cval = 0.f; r.cval_im = 0.f; r.zero = false; r.has_const = false; r.beyond = false; r.active = w < p.nwork; r.dst = nullptr; r.cl0 = 0;
    if (!r.active) return r;
    // 32-bit index arithmetic (the host only takes this path when the work count fits; 64-bit division is ~5x dearer)
    const uint32_t w32 = (uint32_t)w, rpt = (uint32_t)p.rows_per_tile;
    const uint32_t tile = (uint32_t)fdiv((int)w32, p.fd_rpt);
    uint32_t row = w32 - tile * rpt;
    r.dst = p.ws + (int64_t)NDC_RING_TILE(tile) * p.tile_elems + (int64_t)row * pitch;
    uint32_t tt = tile;
    const uint32_t ql = (uint32_t)fdiv((int)tt, p.fd_nt[N - 1]);
    const uint32_t tl = tt - ql * (uint32_t)p.ntiles[N - 1]; tt = ql;
    r.cl0 = (int64_t)tl * p.V[N - 1];
    int64_t c[2] = {0, 0};
#pragma unroll
    for (int a = N - 2; a >= 0; a--) {
        const uint32_t qt = (uint32_t)fdiv((int)tt, p.fd_nt[a]), qr = (uint32_t)fdiv((int)row, p.fd_F[a]);
        const uint32_t ta = tt - qt * (uint32_t)p.ntiles[a]; tt = qt;
        const uint32_t ra = row - qr * (uint32_t)p.F[a]; row = qr;
        c[a] = (int64_t)ta * p.V[a] + ra;
        if (c[a] >= p.P[a]) r.beyond = true;
    }
    if (r.beyond) return r;
    r.base = (int64_t)tt * p.xstr_batch;         // what is left of the tile index is the batch index (0 without a batch)
#pragma unroll
    for (int a = N - 2; a >= 0; a--) {
        if (r.has_const) continue;
        const int32_t m = p.map[a][c[a]];
        if (m >= 0) r.base += (int64_t)m * p.xstr[a];
        else if (m == NDC_MAP_INIT) r.zero = true;
        else { r.has_const = true; r.cval = (m == NDC_MAP_CONST_FRONT) ? p.cfront[a] : p.cback[a]; r.cval_im = (m == NDC_MAP_CONST_FRONT) ? p.cfront_im[a] : p.cback_im[a]; }
    }
    return r;
}

// two neighbouring samples (padded columns cl, cl + 1 of the last axis) of a row that is not plain array data there: constant
// borders, index-mapped borders, cells no padding pass writes (0), columns beyond the padded extent (0)
template <int N> __device__ __noinline__ pc border_pair(const RowParams &p, int64_t base, int flags, float cval, int64_t cl0)
{
    constexpr int al = N - 1;
    const bool active = flags & 1, beyond = flags & 2, zero = flags & 4, has_const = flags & 8;
    const bool plain = active && !beyond && !zero && !has_const;
    float q[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int64_t cl = cl0 + h;
        float val = 0.f;
        const int64_t cc = cl - p.pf[al];
        if (plain && cc >= 0 && cc < p.n[al]) val = __ldg(p.x + base + cc * p.xstr[al]);      // in-array sample: no map lookup
        else if (active && !beyond && cl < p.P[al]) {
            const int32_t m = p.map[al][cl];
            if (m == NDC_MAP_CONST_FRONT) val = p.cfront[al];
            else if (m == NDC_MAP_CONST_BACK) val = p.cback[al];
            else if (has_const) val = cval;
            else if (m != NDC_MAP_INIT && !zero) val = __ldg(p.x + base + (int64_t)m * p.xstr[al]);
        }
        q[h] = val;
    }
    return pk::mk(q[0], q[1]);
}

// (Measured and rejected, round 2: ONE CTA of 16 warps per SM meeting at a barrier before every row, so that the warps of a scheduler
// fetch the same instruction lines together -- `no_instructions` is 21 % of this kernel's stall samples, 11 % of row_inv's,
// profiles/r02_ncu_full_c5s.md.  c5: row_fwd 1.94 -> 1.98 ms, row_inv 1.64 -> 1.69 ms, profiles/r02_variants_row_lockstep.log.)
template <int T, int N>
__global__ void __launch_bounds__(128, 4) row_fwd(const __grid_constant__ RowParams p)
{
    pdl_launch_dependents();
    constexpr int WPB = 4;
    constexpr int L = RowCfg<T>::L, M = RowCfg<T>::M, G = RowCfg<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_tw = reinterpret_cast<pc *>(smem_raw);          // s_tw[k1 * T + t] = W_L^{t k1}
    pc *s_twr = s_tw + L;                                 // (-i/2) w^k, k < L/2
    pc *s_ex = s_twr + L / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) s_tw[idx] = ld_pc(p.tw + (idx / T) * (idx % T));
    for (int idx = threadIdx.x; idx < L / 2; idx += blockDim.x) { const cf w = p.twr[idx]; s_twr[idx] = pk::mk(0.5f * w.im, -0.5f * w.re); }
    __syncthreads();
    pdl_wait();
    pc *sb = s_ex + warp * RowCfg<T>::wstride + g * RowCfg<T>::gstride;
    constexpr int al = N - 1;
    const int src_lane = g * T + ((T - t) % T);
    const pc half = pk::mk(0.5f, 0.5f);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    // The row AFTER this one is resolved (tile / row decode, outer-axis border maps: a chain of divisions and a dependent global load)
    // and its samples are requested into L2 (prefetch.global.L2: no registers) before this row's transform starts, so neither the
    // map lookup nor the DRAM latency of the samples sits in front of the next row's loads.
    const int64_t wstep = (int64_t)gridDim.x * WPB;
    int64_t wi = (int64_t)blockIdx.x * WPB + warp;
    RowSrcInfo nxt = resolve_fwd_row<N>(p, wi * G + g, L + kPad);
    auto prefetch_next = [&]() {
        if (nxt.active && !nxt.beyond && !nxt.zero && !nxt.has_const && p.xstr[al] == 1) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int64_t cc = nxt.cl0 - p.pf[al] + 32 * (t + T * h);          // one 128-byte line per lane and half
                if (cc >= 0 && cc < p.n[al]) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + nxt.base + cc));
            }
        }
    };
    for (; wi < nwarp_items; wi += wstep) {
        const RowSrcInfo ri = nxt;
        if (wi + wstep < nwarp_items) {
            nxt = resolve_fwd_row<N>(p, (wi + wstep) * G + g, L + kPad);
            if (p.pf_mode == 1) prefetch_next();
        }
        if (__all_sync(0xffffffffu, !ri.active || ri.beyond || ((ri.has_const ? ri.cval == 0.f : ri.zero) && p.cfront[al] == 0.f && p.cback[al] == 0.f))) {
            // rows beyond the padded extent of an outer axis, and rows that sit in a Zeros border (a constant fill of 0, half_dim.rs:30-49) or
            // a never-written plane of an outer axis while the last axis has no non-zero constant border of its own (which would win there):
            // zero spectrum, no transform -- most rows of a small Zeros-padded problem (c3: 33 -> 19 us for this kernel)
            if (ri.active) {
                float4 *z4 = reinterpret_cast<float4 *>(ri.dst);
                for (int q = t; q < (L + kPad) / 2; q += T) z4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            continue;
        }
        pc v[32];
        // a beyond-extent / zero row without a constant last axis is all zero -- but the last axis may still be a constant border
        const bool interior = ri.active && !ri.beyond && !ri.zero && !ri.has_const && p.xstr[al] == 1 && ri.cl0 >= p.pf[al] && ri.cl0 + 2 * L <= p.pf[al] + p.n[al];
        if (interior) {
            const float *src = p.x + ri.base + (ri.cl0 - p.pf[al]);
            if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
                const unsigned long long *s2 = reinterpret_cast<const unsigned long long *>(src);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j].v = __ldg(s2 + t + T * j);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) { const int e = 2 * (t + T * j); v[j] = pk::mk(__ldg(src + e), __ldg(src + e + 1)); }
            }
        } else {
            // a tile at the edge of the last axis (or a row with a constant / zero outer border): sample pairs that lie inside the array
            // are still fetched 8 bytes at a time; only the pairs that touch the border go through the index map, in a function of
            // its own -- unrolled inline 32 times it was 3 400 instructions that every edge-tile row (2 of c5's 17 tile columns)
            // walked through at 2.7 x the cost of an interior row
            const bool plain = ri.active && !ri.beyond && !ri.zero && !ri.has_const && p.xstr[al] == 1;
            const float *rowp = p.x + ri.base - p.pf[al];                 // padded column cl of this row is rowp[cl] where it is an array sample
            const int64_t lo = p.pf[al], hi = p.pf[al] + p.n[al];
            const int flags = (ri.active ? 1 : 0) | (ri.beyond ? 2 : 0) | (ri.zero ? 4 : 0) | (ri.has_const ? 8 : 0);
            if (!plain && p.xstr[al] == 1) {
                // a row without array samples (an outer axis sits in a Zeros / Const border or a never-written plane): its value depends on
                // the column only through the last axis' own constant borders -- no loads from x, map lookups in the border columns only, no calls (these are most
                // rows of a small padded problem: every pair of them through border_pair cost c3 12 us)
                const float rowval = ri.has_const ? ri.cval : 0.f;
                const bool live = ri.active && !ri.beyond;
                if (!live || (rowval == 0.f && p.cfront[al] == 0.f && p.cback[al] == 0.f)) {
                    // identically zero (a Zeros border row, a never-written plane, a row beyond the padded extent; no non-zero constant border
                    // of the last axis): such a row only comes here when it shares its warp with rows that do have samples (T < 32) -- the warp
                    // runs both branches one after the other, so this one must cost nothing (c2 with Zeros borders: 30 us against 27 with Reflect)
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = pk::mk(0.f, 0.f);
                } else
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int64_t cl = ri.cl0 + 2 * (t + T * j);
                    float q[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int64_t c1 = cl + h;
                        float val = 0.f;
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
                    v[j] = pk::mk(q[0], q[1]);
                }
            } else
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int64_t cl = ri.cl0 + 2 * (t + T * j);
                if (plain && cl >= lo && cl + 1 < hi) {
                    const float *s = rowp + cl;
                    if ((reinterpret_cast<uintptr_t>(s) & 7) == 0) v[j].v = __ldg(reinterpret_cast<const unsigned long long *>(s));
                    else v[j] = pk::mk(__ldg(s), __ldg(s + 1));
                } else v[j] = border_pair<N>(p, ri.base, flags, ri.cval, cl);
            }
        }
        pk::dft<false, 32>(v);                                           // over j -> k1
        twiddle_store<32, false>(v, s_tw + t, T, sb + t, T + 1);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int i = 0; i < T; i++) v[m * T + i] = sb[(t + T * m) * (T + 1) + i];
        __syncwarp();
        if (p.pf_mode == 2 && wi + wstep < nwarp_items) prefetch_next();
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<false, T>(v + m * T);         // v[m*T + k2] = Z[k], k = t + T m + 32 k2
        // partner Z[L-k] and R2C post-processing; primaries are k2 < T/2 (k < L/2):
        //   X[k] = E + w^k O,  X[L-k] = conj(E - w^k O),  2E = Z[k] + conj Z[L-k],  w^k O = (-i/2) w^k (Z[k] - conj Z[L-k])
        ulonglong2 *dst4 = reinterpret_cast<ulonglong2 *>(ri.dst);
#pragma unroll
        for (int m = 0; m < M; m++) {
#pragma unroll
            for (int k2 = 0; k2 < T / 2; k2++) {
                const int ia = (M - 1 - m) * T + (T - 1 - k2);                                    // lanes t > 0
                const int ib = ((M - m) % M) * T + (m > 0 ? T - 1 - k2 : (k2 > 0 ? T - k2 : T / 2)); // lane t == 0 (own registers)
                float px = __shfl_sync(0xffffffffu, pk::re(v[ia]), src_lane);
                float py = __shfl_sync(0xffffffffu, pk::im(v[ia]), src_lane);
                if (t == 0) { px = pk::re(v[ib]); py = pk::im(v[ib]); }
                const pc zk = v[m * T + k2];
                const int k = t + T * m + 32 * k2;
                ulonglong2 o4;
                if (k == 0) {
                    o4.x = pk::mk(pk::re(zk) + pk::im(zk), 0.f).v;                                 // X[0]
                    o4.y = pk::mk(px, -py).v;                                                      // X[L/2] = conj Z[L/2]
                    if (ri.active) *reinterpret_cast<float4 *>(ri.dst + L) = make_float4(pk::re(zk) - pk::im(zk), 0.f, 0.f, 0.f);   // Nyquist X[L]
                } else {
                    const pc cp = pk::mk(px, -py);
                    const pc e2 = pk::add(zk, cp), d = pk::sub(zk, cp);
                    const pc tw = pk::cmul(d, s_twr[k]);
                    o4.x = pk::fma(e2, half, tw).v;
                    o4.y = pk::conj(pk::fma(e2, half, pk::neg(tw))).v;
                }
                if (ri.active) dst4[k] = o4;
            }
        }
        if (ri.active && t > 0 && t < 4) *reinterpret_cast<float4 *>(ri.dst + L + 2 * t) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- row inverse + crop + decimate -------------------------------------------------------------------------------------------
struct RowInvInfo {
    const cf *src;
    int64_t orow;      // output element offset of the row
    int tl;
    bool active;
};
template <int N> __device__ __forceinline__ RowInvInfo resolve_inv_row(const RowParams &p, int64_t w, int pitch)
{
    RowInvInfo r; r.active = w < p.nwork; r.src = p.ws; r.orow = 0; r.tl = 0;
    if (!r.active) return r;
    const uint32_t ntl = (uint32_t)p.ntiles[N - 1];
    const uint32_t w32 = (uint32_t)w;
    r.tl = (int)(w32 % ntl);
    uint32_t orow = w32 / ntl;                   // row-major index over the outer output axes (fits 32 bits, checked on the host)
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
    r.src = p.ws + NDC_RING_TILE(tile) * p.tile_elems + row * pitch;
    r.orow = obase * p.O[N - 1];
    return r;
}

template <int T, int N>
__global__ void __launch_bounds__(128, 4) row_inv(const __grid_constant__ RowParams p)
{
    pdl_launch_dependents();
    constexpr int WPB = 4;
    constexpr int L = RowCfg<T>::L, M = RowCfg<T>::M, G = RowCfg<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_tw = reinterpret_cast<pc *>(smem_raw);          // TRANSPOSED for the inverse flow: s_tw[i * 32 + k1] = W_L^{i k1} (k1 is the lane-dependent index)
    pc *s_twr = s_tw + L;                                 // i conj(w^k), k < L/2
    pc *s_ex = s_twr + L / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) s_tw[idx] = ld_pc(p.tw + (idx >> 5) * (idx & 31));
    for (int idx = threadIdx.x; idx < L / 2; idx += blockDim.x) { const cf w = p.twr[idx]; s_twr[idx] = pk::mk(w.im, w.re); }
    __syncthreads();
    pdl_wait();
    pc *sb = s_ex + warp * RowCfg<T>::wstride + g * RowCfg<T>::gstride;
    ulonglong2 *sb4 = reinterpret_cast<ulonglong2 *>(s_ex + warp * RowCfg<T>::wstride) + g * (L / 2 + 1);  // staging: L/2 slots + the Nyquist column, linear
    constexpr int al = N - 1;
    const int src_lane = g * T + ((T - t) % T);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    const int64_t wstep = (int64_t)gridDim.x * WPB;
    int64_t wi = (int64_t)blockIdx.x * WPB + warp;
    if (wi >= nwarp_items) return;
    RowInvInfo ri = resolve_inv_row<N>(p, wi * G + g, L + kPad);
    auto stage = [&](const RowInvInfo &r) {
        if (r.active) {
            const ulonglong2 *s4 = reinterpret_cast<const ulonglong2 *>(r.src);
#pragma unroll
            for (int q = 0; q < 16; q++) cp_async16(sb4 + t + T * q, s4 + t + T * q);
            if (t == 0) cp_async16(sb4 + L / 2, s4 + L / 2);             // column L: the Nyquist bin
        }
        cp_async_commit();
    };
    stage(ri);
    for (; wi < nwarp_items; wi += wstep) {
        pc v[32], b[16];
        cp_async_wait_all();
        __syncwarp();
        pc nq; nq.v = sb4[L / 2].x;
        const float nyq = pk::re(nq);
        // C2R pre-processing (x2): Z[k] = E + O', Z[L-k] = conj(E - O'), E = Y[k] + conj Y[L-k], O' = i conj(w^k) (Y[k] - conj Y[L-k])
#pragma unroll
        for (int m = 0; m < M; m++) {
#pragma unroll
            for (int k2 = 0; k2 < T / 2; k2++) {
                const int k = t + T * m + 32 * k2;
                const ulonglong2 q = sb4[k];
                pc yk, ym; yk.v = q.x; ym.v = q.y;
                pc zk, zp;
                if (k == 0) {
                    zk = pk::mk(pk::re(yk) + nyq, pk::re(yk) - nyq);           // Z[0] from the real DC / Nyquist bins
                    zp = pk::mk(2.f * pk::re(ym), -2.f * pk::im(ym));          // Z[L/2] = 2 conj Y[L/2]
                } else {
                    const pc cm = pk::conj(ym);
                    const pc E = pk::add(yk, cm), D = pk::sub(yk, cm);
                    const pc O = pk::cmul(D, s_twr[k]);
                    zk = pk::add(E, O);
                    zp = pk::conj(pk::sub(E, O));
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
                float px = __shfl_sync(0xffffffffu, pk::re(b[ia]), src_lane);
                float py = __shfl_sync(0xffffffffu, pk::im(b[ia]), src_lane);
                if (t == 0) { px = pk::re(b[ib]); py = pk::im(b[ib]); }
                v[mr * T + k2r] = pk::mk(px, py);
            }
        }
        // inverse: radix T over k2, conj twiddle, exchange, radix 32 over k1
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<true, T>(v + m * T);
#pragma unroll
        for (int m = 0; m < M; m++) twiddle_store<T, true>(v + m * T, s_tw + (t + T * m), 32, sb + (t + T * m) * (T + 1), 1);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) v[k1] = sb[k1 * (T + 1) + t];
        __syncwarp();
        const RowInvInfo cur = ri;
        if (wi + wstep < nwarp_items) { ri = resolve_inv_row<N>(p, (wi + wstep) * G + g, L + kPad); stage(ri); }
        pk::dft<true, 32>(v);                                            // v[j] = z[t + T j] = (y[2n], y[2n+1]), n = t + T j
        if (!cur.active) continue;
        const int Kd1 = p.Kd[al];
        const int64_t mbase = (int64_t)cur.tl * p.V[al];
        if (p.s[al] == 1) {
            float *outp = p.out + (cur.orow + mbase - (Kd1 - 1));                          // tile sample i -> outp[i]
            const int64_t room = p.O[al] - (mbase - (Kd1 - 1));
            const int i0 = 2 * t;
            const KeepRange kr = keep_range<2 * T>(Kd1 - 1, room < (int64_t)(2 * L) ? (int)room : 2 * L, i0);
            const bool vec_ok = (reinterpret_cast<uintptr_t>(outp) & 7) == 0;
            float *o = outp + i0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const bool ok0 = j >= kr.ja0 && j < kr.jb0, ok1 = j >= kr.ja1 && j < kr.jb1;
                if (vec_ok && ok0 && ok1) *reinterpret_cast<unsigned long long *>(o + 2 * T * j) = v[j].v;
                else {
                    if (ok0) o[2 * T * j] = pk::re(v[j]);
                    if (ok1) o[2 * T * j + 1] = pk::im(v[j]);
                }
            }
        } else {
            // positions along the last axis fit 32 bits (checked on the host); a stride larger than the axis keeps output 0 only
            const uint32_t s1 = (uint32_t)(p.s[al] < 0x7fffffff ? p.s[al] : 0x7fffffff), O1 = (uint32_t)p.O[al];
#pragma unroll
            for (int j = 0; j < 32; j++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * (t + T * j) + h;
                    if (i < Kd1 - 1) continue;
                    const uint32_t q = (uint32_t)(mbase + i - (Kd1 - 1));
                    const uint32_t o = q / s1;
                    if (o * s1 != q) continue;
                    if (o < O1) p.out[cur.orow + o] = h ? pk::im(v[j]) : pk::re(v[j]);
                }
            }
        }
    }
}

// ---- Complex<f32> rows: C2C along the last axis (tiles of L = 32 T complex samples, workspace rows of exactly L columns) ----------
// Same lane-group structure as row_fwd / row_inv without the real-transform algebra: the column kernels do not care what a column
// holds, so rank 2 and 3 complex problems reuse col_pass unchanged (processor/complex.rs:33-145 in the reference).
template <int T> struct RowCxCfg { static constexpr int L = 32 * T, smem = (L + 4 * RowCfg<T>::wstride) * 8; };

template <int T, int N>
__global__ void __launch_bounds__(128, 4) row_fwd_c(const __grid_constant__ RowParams p)
{
    pdl_launch_dependents();
    constexpr int L = RowCfg<T>::L, M = RowCfg<T>::M, G = RowCfg<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_tw = reinterpret_cast<pc *>(smem_raw);          // s_tw[k1 * T + t] = W_L^{t k1}
    pc *s_ex = s_tw + L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) s_tw[idx] = ld_pc(p.tw + (idx / T) * (idx % T));
    __syncthreads();
    pdl_wait();
    pc *sb = s_ex + warp * RowCfg<T>::wstride + g * RowCfg<T>::gstride;
    constexpr int al = N - 1;
    const cf *xc = reinterpret_cast<const cf *>(p.x);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    for (int64_t wi = (int64_t)blockIdx.x * 4 + warp; wi < nwarp_items; wi += (int64_t)gridDim.x * 4) {
        const RowSrcInfo ri = resolve_fwd_row<N>(p, wi * G + g, L);
        if (__all_sync(0xffffffffu, !ri.active || ri.beyond)) {
            if (ri.active) for (int q = t; q < L; q += T) st_pc(ri.dst + q, pk::mk(0.f, 0.f));
            continue;
        }
        pc v[32];
        const bool interior = ri.active && !ri.beyond && !ri.zero && !ri.has_const && p.xstr[al] == 1 && ri.cl0 >= p.pf[al] && ri.cl0 + L <= p.pf[al] + p.n[al];
        if (interior) {
            const cf *src = xc + ri.base + (ri.cl0 - p.pf[al]);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j].v = __ldg(reinterpret_cast<const unsigned long long *>(src + t + T * j));
        } else {
            const bool plain = ri.active && !ri.beyond && !ri.zero && !ri.has_const;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int64_t cl = ri.cl0 + t + T * j;
                pc val = pk::mk(0.f, 0.f);
                const int64_t cc = cl - p.pf[al];
                if (plain && cc >= 0 && cc < p.n[al]) val.v = __ldg(reinterpret_cast<const unsigned long long *>(xc + ri.base + cc * p.xstr[al]));
                else if (ri.active && !ri.beyond && cl < p.P[al]) {
                    const int32_t m = p.map[al][cl];
                    if (m == NDC_MAP_CONST_FRONT) val = pk::mk(p.cfront[al], p.cfront_im[al]);
                    else if (m == NDC_MAP_CONST_BACK) val = pk::mk(p.cback[al], p.cback_im[al]);
                    else if (ri.has_const) val = pk::mk(ri.cval, ri.cval_im);
                    else if (m != NDC_MAP_INIT && !ri.zero) val.v = __ldg(reinterpret_cast<const unsigned long long *>(xc + ri.base + (int64_t)m * p.xstr[al]));
                }
                v[j] = val;
            }
        }
        pk::dft<false, 32>(v);
        twiddle_store<32, false>(v, s_tw + t, T, sb + t, T + 1);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int i = 0; i < T; i++) v[m * T + i] = sb[(t + T * m) * (T + 1) + i];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<false, T>(v + m * T);         // v[m*T + k2] = X[k], k = t + T m + 32 k2
        if (ri.active) {
#pragma unroll
            for (int m = 0; m < M; m++)
#pragma unroll
                for (int k2 = 0; k2 < T; k2++) st_pc(ri.dst + t + T * m + 32 * k2, v[m * T + k2]);
        }
    }
}

template <int T, int N>
__global__ void __launch_bounds__(128, 4) row_inv_c(const __grid_constant__ RowParams p)
{
    pdl_launch_dependents();
    constexpr int L = RowCfg<T>::L, M = RowCfg<T>::M, G = RowCfg<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_tw = reinterpret_cast<pc *>(smem_raw);          // transposed: s_tw[i * 32 + k1] = W_L^{i k1}
    pc *s_ex = s_tw + L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) s_tw[idx] = ld_pc(p.tw + (idx >> 5) * (idx & 31));
    __syncthreads();
    pdl_wait();
    pc *sb = s_ex + warp * RowCfg<T>::wstride + g * RowCfg<T>::gstride;
    constexpr int al = N - 1;
    cf *outc = reinterpret_cast<cf *>(p.out);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    for (int64_t wi = (int64_t)blockIdx.x * 4 + warp; wi < nwarp_items; wi += (int64_t)gridDim.x * 4) {
        const RowInvInfo ri = resolve_inv_row<N>(p, wi * G + g, L);
        pc v[32];
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int k2 = 0; k2 < T; k2++) v[m * T + k2] = ri.active ? ld_pc(ri.src + t + T * m + 32 * k2) : pk::mk(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<true, T>(v + m * T);
#pragma unroll
        for (int m = 0; m < M; m++) twiddle_store<T, true>(v + m * T, s_tw + (t + T * m), 32, sb + (t + T * m) * (T + 1), 1);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) v[k1] = sb[k1 * (T + 1) + t];
        __syncwarp();
        pk::dft<true, 32>(v);                                            // v[j] = y[t + T j]
        if (!ri.active) continue;
        const int Kd1 = p.Kd[al];
        const int64_t mbase = (int64_t)ri.tl * p.V[al];
        const uint32_t s1 = (uint32_t)(p.s[al] < 0x7fffffff ? p.s[al] : 0x7fffffff), O1 = (uint32_t)p.O[al];
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const int i = t + T * j;
            if (i < Kd1 - 1) continue;
            const uint32_t q = (uint32_t)(mbase + i - (Kd1 - 1));
            const uint32_t o = s1 == 1 ? q : q / s1;
            if (s1 != 1 && o * s1 != q) continue;
            if (o < O1) st_pc(outc + ri.orow + o, v[j]);
        }
    }
}

// ---- rank 1: the whole pipeline in one pass over memory ------------------------------------------------------------------------
// One lane group per overlap-save tile: border-mapped loads, forward transform, R2C post-processing, multiply by the kernel
// spectrum, C2R pre-processing, inverse transform, crop + stride -- all in registers and one warp-private exchange buffer.  The
// paired slot k = (X[k], X[L-k]) that row_fwd would store is exactly what row_inv loads on the same lane, so nothing is
// exchanged between the two halves and there is no workspace: the kernel reads x once and writes the output once.
template <int T> struct Row1dCfg {
    static constexpr int L = 32 * T;
    static constexpr int smem = (L + L + L / 2 + L / 2 + 4 * RowCfg<T>::wstride) * 8;     // W (forward order), W (inverse order), post / pre tables, exchange
};

template <int T>
__global__ void __launch_bounds__(128, 3) row1d(const __grid_constant__ RowParams p)
{
    pdl_launch_dependents();
    constexpr int L = RowCfg<T>::L, M = RowCfg<T>::M, G = RowCfg<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_twF = reinterpret_cast<pc *>(smem_raw);         // s_twF[k1 * T + t] = W_L^{t k1}
    pc *s_twI = s_twF + L;                                // s_twI[i * 32 + k1] = W_L^{i k1}
    pc *s_post = s_twI + L;                               // (-i/2) w^k
    pc *s_pre = s_post + L / 2;                           // i conj(w^k)
    pc *s_ex = s_pre + L / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) {
        s_twF[idx] = ld_pc(p.tw + (idx / T) * (idx % T));
        s_twI[idx] = ld_pc(p.tw + (idx >> 5) * (idx & 31));
    }
    for (int idx = threadIdx.x; idx < L / 2; idx += blockDim.x) {
        const cf w = p.twr[idx];
        s_post[idx] = pk::mk(0.5f * w.im, -0.5f * w.re);
        s_pre[idx] = pk::mk(w.im, w.re);
    }
    __syncthreads();
    pdl_wait();
    pc *sb = s_ex + warp * RowCfg<T>::wstride + g * RowCfg<T>::gstride;
    const int src_lane = g * T + ((T - t) % T);
    const pc half = pk::mk(0.5f, 0.5f);
    const ulonglong2 *k4 = reinterpret_cast<const ulonglong2 *>(p.kfast);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    for (int64_t wi = (int64_t)blockIdx.x * 4 + warp; wi < nwarp_items; wi += (int64_t)gridDim.x * 4) {
        const int64_t tl = wi * G + g;
        const bool active = tl < p.nwork;
        const int64_t cl0 = tl * p.V[0];
        pc v[32];
        // ---- load: padded samples [cl0, cl0 + 2L) ----
        const bool interior = active && p.xstr[0] == 1 && cl0 >= p.pf[0] && cl0 + 2 * L <= p.pf[0] + p.n[0];
        if (interior) {
            const float *src = p.x + (cl0 - p.pf[0]);
            if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
                const unsigned long long *s2 = reinterpret_cast<const unsigned long long *>(src);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j].v = __ldg(s2 + t + T * j);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) { const int e = 2 * (t + T * j); v[j] = pk::mk(__ldg(src + e), __ldg(src + e + 1)); }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float q[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int64_t cl = cl0 + 2 * (t + T * j) + h;
                    float val = 0.f;
                    if (active && cl < p.P[0]) {
                        const int64_t cc = cl - p.pf[0];
                        if (cc >= 0 && cc < p.n[0]) val = __ldg(p.x + cc * p.xstr[0]);
                        else {
                            const int32_t m = p.map[0][cl];
                            if (m == NDC_MAP_CONST_FRONT) val = p.cfront[0];
                            else if (m == NDC_MAP_CONST_BACK) val = p.cback[0];
                            else if (m != NDC_MAP_INIT) val = __ldg(p.x + (int64_t)m * p.xstr[0]);
                        }
                    }
                    q[h] = val;
                }
                v[j] = pk::mk(q[0], q[1]);
            }
        }
        // ---- forward: radix 32 over j, twiddle, exchange, radix T ----
        pk::dft<false, 32>(v);
        twiddle_store<32, false>(v, s_twF + t, T, sb + t, T + 1);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int i = 0; i < T; i++) v[m * T + i] = sb[(t + T * m) * (T + 1) + i];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<false, T>(v + m * T);         // v[m*T + k2] = Z[k], k = t + T m + 32 k2
        // ---- R2C post-processing, multiply by the kernel spectrum, C2R pre-processing: slot k = (X[k], X[L-k]) ----
        pc b[16];
        const float knyq = pk::re(ld_pc(p.kfast + L));                    // K[L] (real: the kernel is real)
#pragma unroll
        for (int m = 0; m < M; m++) {
#pragma unroll
            for (int k2 = 0; k2 < T / 2; k2++) {
                const int ia = (M - 1 - m) * T + (T - 1 - k2);                                    // lanes t > 0
                const int ib = ((M - m) % M) * T + (m > 0 ? T - 1 - k2 : (k2 > 0 ? T - k2 : T / 2)); // lane t == 0 (own registers)
                float px = __shfl_sync(0xffffffffu, pk::re(v[ia]), src_lane);
                float py = __shfl_sync(0xffffffffu, pk::im(v[ia]), src_lane);
                if (t == 0) { px = pk::re(v[ib]); py = pk::im(v[ib]); }
                const pc zk = v[m * T + k2];
                const int k = t + T * m + 32 * k2;
                const ulonglong2 kq = k4[k];                                                       // (K[k], K[L-k]) ; slot 0: (K[0], K[L/2])
                pc ka, kb; ka.v = kq.x; kb.v = kq.y;
                pc nzk, nzp;
                if (k == 0) {
                    const float x0 = pk::re(zk) + pk::im(zk), xl = pk::re(zk) - pk::im(zk);      // X[0], X[L]: real
                    const float y0 = x0 * pk::re(ka), yl = xl * knyq;
                    const pc ym = pk::cmul(pk::mk(px, -py), kb);                                   // Y[L/2] = conj Z[L/2] * K[L/2]
                    nzk = pk::mk(y0 + yl, y0 - yl);                                                // Z'[0]
                    nzp = pk::mk(2.f * pk::re(ym), -2.f * pk::im(ym));                             // Z'[L/2] = 2 conj Y[L/2]
                } else {
                    const pc cp = pk::mk(px, -py);
                    const pc e2 = pk::add(zk, cp), d = pk::sub(zk, cp);
                    const pc tw = pk::cmul(d, s_post[k]);
                    const pc yk = pk::cmul(pk::fma(e2, half, tw), ka);                             // Y[k] = X[k] K[k]
                    const pc ym = pk::cmul(pk::conj(pk::fma(e2, half, pk::neg(tw))), kb);          // Y[L-k] = X[L-k] K[L-k]
                    const pc cm = pk::conj(ym);
                    const pc E = pk::add(yk, cm), D = pk::sub(yk, cm);
                    const pc O = pk::cmul(D, s_pre[k]);
                    nzk = pk::add(E, O);
                    nzp = pk::conj(pk::sub(E, O));
                }
                b[m * (T / 2) + k2] = nzp;
                v[m * T + k2] = nzk;          // in place: partners are only ever read from the upper-half registers (k2 >= T/2)
            }
        }
        // deliver Z'[L-k] to its owner: register (mr, k2r >= T/2) of lane t comes from lane (T-t)%T
#pragma unroll
        for (int mr = 0; mr < M; mr++) {
#pragma unroll
            for (int k2r = T / 2; k2r < T; k2r++) {
                const int ia = (M - 1 - mr) * (T / 2) + (T - 1 - k2r);                                            // source lanes t' > 0
                const int ib = mr > 0 ? (M - mr) * (T / 2) + (T - 1 - k2r) : (k2r == T / 2 ? 0 : (T - k2r));      // lane 0: own b[]
                float px = __shfl_sync(0xffffffffu, pk::re(b[ia]), src_lane);
                float py = __shfl_sync(0xffffffffu, pk::im(b[ia]), src_lane);
                if (t == 0) { px = pk::re(b[ib]); py = pk::im(b[ib]); }
                v[mr * T + k2r] = pk::mk(px, py);
            }
        }
        // ---- inverse: radix T over k2, conj twiddle, exchange, radix 32 over k1 ----
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<true, T>(v + m * T);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++)
            twiddle_store<T, true>(v + m * T, s_twI + (t + T * m), 32, sb + (t + T * m) * (T + 1), 1);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) v[k1] = sb[k1 * (T + 1) + t];
        __syncwarp();
        pk::dft<true, 32>(v);                                            // v[j] = (y[2n], y[2n+1]), n = t + T j
        if (!active) continue;
        // ---- crop [Kd-1, 2L) of the tile, stride, store ----
        const int Kd1 = p.Kd[0];
        const int64_t mbase = tl * p.V[0];
        if (p.s[0] == 1) {
            float *outp = p.out + (mbase - (Kd1 - 1));                                     // tile sample i -> outp[i] (see row_inv)
            const int64_t room = p.O[0] - (mbase - (Kd1 - 1));
            const int i0 = 2 * t;
            const KeepRange kr = keep_range<2 * T>(Kd1 - 1, room < (int64_t)(2 * L) ? (int)room : 2 * L, i0);
            const bool vec_ok = (reinterpret_cast<uintptr_t>(outp) & 7) == 0;
            float *o = outp + i0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const bool ok0 = j >= kr.ja0 && j < kr.jb0, ok1 = j >= kr.ja1 && j < kr.jb1;
                if (vec_ok && ok0 && ok1) *reinterpret_cast<unsigned long long *>(o + 2 * T * j) = v[j].v;
                else {
                    if (ok0) o[2 * T * j] = pk::re(v[j]);
                    if (ok1) o[2 * T * j + 1] = pk::im(v[j]);
                }
            }
        } else {
            const int64_t s1 = p.s[0];
#pragma unroll
            for (int j = 0; j < 32; j++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * (t + T * j) + h;
                    if (i < Kd1 - 1) continue;
                    const int64_t q = mbase + i - (Kd1 - 1);
                    const int64_t o = q / s1;
                    if (o * s1 == q && o < p.O[0]) p.out[o] = h ? pk::im(v[j]) : pk::re(v[j]);
                }
            }
        }
    }
}

// ---- rank 1, Complex<f32>: forward C2C, multiply by the kernel spectrum, inverse C2C, crop -- one pass, no workspace ----------------
template <int T> struct Row1dCxCfg { static constexpr int L = 32 * T, smem = (L + L + 4 * RowCfg<T>::wstride) * 8; };

template <int T>
__global__ void __launch_bounds__(128, 4) row1d_c(const __grid_constant__ RowParams p)
{
    pdl_launch_dependents();
    constexpr int L = RowCfg<T>::L, M = RowCfg<T>::M, G = RowCfg<T>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_twF = reinterpret_cast<pc *>(smem_raw);         // s_twF[k1 * T + t] = W_L^{t k1}
    pc *s_twI = s_twF + L;                                // s_twI[i * 32 + k1] = W_L^{i k1}
    pc *s_ex = s_twI + L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / T, t = lane % T;
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) {
        s_twF[idx] = ld_pc(p.tw + (idx / T) * (idx % T));
        s_twI[idx] = ld_pc(p.tw + (idx >> 5) * (idx & 31));
    }
    __syncthreads();
    pdl_wait();
    pc *sb = s_ex + warp * RowCfg<T>::wstride + g * RowCfg<T>::gstride;
    const cf *xc = reinterpret_cast<const cf *>(p.x);
    cf *outc = reinterpret_cast<cf *>(p.out);
    const int64_t nwarp_items = (p.nwork + G - 1) / G;
    for (int64_t wi = (int64_t)blockIdx.x * 4 + warp; wi < nwarp_items; wi += (int64_t)gridDim.x * 4) {
        const int64_t tl = wi * G + g;
        const bool active = tl < p.nwork;
        const int64_t cl0 = tl * p.V[0];
        pc v[32];
        const bool interior = active && p.xstr[0] == 1 && cl0 >= p.pf[0] && cl0 + L <= p.pf[0] + p.n[0];
        if (interior) {
            const cf *src = xc + (cl0 - p.pf[0]);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j].v = __ldg(reinterpret_cast<const unsigned long long *>(src + t + T * j));
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int64_t cl = cl0 + t + T * j;
                pc val = pk::mk(0.f, 0.f);
                if (active && cl < p.P[0]) {
                    const int64_t cc = cl - p.pf[0];
                    if (cc >= 0 && cc < p.n[0]) val.v = __ldg(reinterpret_cast<const unsigned long long *>(xc + cc * p.xstr[0]));
                    else {
                        const int32_t m = p.map[0][cl];
                        if (m == NDC_MAP_CONST_FRONT) val = pk::mk(p.cfront[0], p.cfront_im[0]);
                        else if (m == NDC_MAP_CONST_BACK) val = pk::mk(p.cback[0], p.cback_im[0]);
                        else if (m != NDC_MAP_INIT) val.v = __ldg(reinterpret_cast<const unsigned long long *>(xc + (int64_t)m * p.xstr[0]));
                    }
                }
                v[j] = val;
            }
        }
        pk::dft<false, 32>(v);
        twiddle_store<32, false>(v, s_twF + t, T, sb + t, T + 1);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int i = 0; i < T; i++) v[m * T + i] = sb[(t + T * m) * (T + 1) + i];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<false, T>(v + m * T);         // v[m*T + k2] = X[k], k = t + T m + 32 k2
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int k2 = 0; k2 < T; k2++) v[m * T + k2] = pk::cmul(v[m * T + k2], ld_pc(p.kfast + t + T * m + 32 * k2));
#pragma unroll
        for (int m = 0; m < M; m++) pk::dft<true, T>(v + m * T);
#pragma unroll
        for (int m = 0; m < M; m++)
            twiddle_store<T, true>(v + m * T, s_twI + (t + T * m), 32, sb + (t + T * m) * (T + 1), 1);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) v[k1] = sb[k1 * (T + 1) + t];
        __syncwarp();
        pk::dft<true, 32>(v);                                            // v[j] = y[t + T j]
        if (!active) continue;
        const int Kd1 = p.Kd[0];
        const int64_t mbase = tl * p.V[0];
        const int64_t s1 = p.s[0];
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const int i = t + T * j;
            if (i < Kd1 - 1) continue;
            const int64_t q = mbase + i - (Kd1 - 1);
            const int64_t o = s1 == 1 ? q : q / s1;
            if (s1 != 1 && o * s1 != q) continue;
            if (o < p.O[0]) st_pc(outc + o, v[j]);
        }
    }
}

// ---- column pass: FWD / INV / FMI ---------------------------------------------------------------------------------------------
struct ColParams {
    cf *ws;
    const cf *kspec;        // kernel spectrum, same (outer, F, inner) geometry as one tile (FMI only)
    const cf *tw;           // exp(-2 pi i j / F)
    int mode;               // 0 forward, 1 inverse, 2 forward * kspec * inverse
    int64_t outer, inner;   // the tile is [outer][F][inner] complex, inner % 8 == 0
    int64_t tile_elems, ntiles_total;
    int64_t nwork;          // ntiles_total * outer * inner / 8
    int skip;               // modes 1 / 2: the first `skip` rows of the axis (= Kd - 1, the aliased head the crop discards) are not stored
    int store_rows;         // col_pass_tma: rows per store box (divides F - skip)
    // col_pass_tma_kres (mode 2, outer == 1): the kernel spectrum regrouped per 8-column block in the order Tensor Memory takes it,
    // and the bundle geometry of the work order (tiles of one block are taken `bundle` at a time; the last bundle has `bundle_last`)
    const cf *kres;
    int bundle, nbundles, bundle_last;
};

template <int E, int Tc> struct ColCfg {
    static constexpr int F = E * Tc, Mc = E / Tc;
    static constexpr int threads = Tc * 8;
    static constexpr int pitch = Tc * 8 + 8;                             // padded k1-row stride of the exchange buffer
    static constexpr int ex = (E * pitch > F * 8 ? E * pitch : F * 8);
    static constexpr int smem = ((Tc == E ? 1 : 2) * F + ex) * 8;   // forward table, transposed table for the inverse when Tc != E, exchange buffer
    static constexpr int min_blocks = (E == 32 && Tc == 32) ? 2 : 4;      // 512-row tiles: 4 CTAs of 128 threads at 126 registers beat 3 at 168 (s512: 52 -> 48 us)
};

template <int E, int Tc>
__global__ void __launch_bounds__(ColCfg<E, Tc>::threads, ColCfg<E, Tc>::min_blocks) col_pass(const __grid_constant__ ColParams p)
{
    pdl_launch_dependents();
    using C = ColCfg<E, Tc>;
    constexpr int F = C::F, Mc = C::Mc, pitch = C::pitch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pc *s_tw = reinterpret_cast<pc *>(smem_raw);        // s_tw[k1 * Tc + i] = W_F^{i k1}, k1 < E, i < Tc   (forward: i is the thread index)
    pc *s_twT = s_tw + F;                               // s_twT[ii * E + k1] = W_F^{ii k1}   (inverse, Tc != E: k1 is the thread-dependent index)
    pc *S = s_tw + (Tc == E ? 1 : 2) * F;
    const int tid = threadIdx.x;
    const int c = tid & 7, i = tid >> 3;                // column of the block, thread index inside the column (< Tc)
    for (int idx = tid; idx < F; idx += C::threads) { s_tw[idx] = ld_pc(p.tw + (idx / Tc) * (idx % Tc)); if (Tc != E) s_twT[idx] = ld_pc(p.tw + (idx / E) * (idx % E)); }
    // work item -> (tile, outer index, 8-column block): 32-bit arithmetic (the host only takes this path when nwork < 2^31;
    // a 64-bit division costs ~80 instructions), decoded ONCE per item: the prefetch of item w + gridDim.x hands its offsets on
    const uint32_t iblocks = (uint32_t)(p.inner / 8), outer = (uint32_t)p.outer;
    struct Item { int64_t off, rel; };           // element offset of the block's first column in ws ; the same inside one tile
    auto decode = [&](int64_t w) {
        const uint32_t w32 = (uint32_t)w, q = w32 / iblocks, ib = w32 - q * iblocks, tile = q / outer, o = q - tile * outer;
        Item it; it.rel = (int64_t)o * F * p.inner + ib * 8; it.off = (int64_t)tile * p.tile_elems + it.rel;
        return it;
    };
    constexpr int kChunks = (F * 4 + C::threads - 1) / C::threads;       // 16-byte chunks per thread
    const uint64_t pol_keep = l2_policy_evict_last();
    auto prefetch = [&](const Item &it) {
        const cf *gn = p.ws + it.off;
#pragma unroll
        for (int m = 0; m < kChunks; m++) {
            const int id = tid + C::threads * m, row = id >> 2, part = id & 3;
            if (F * 4 % C::threads == 0 || id < F * 4) cp_async16(S + row * 8 + part * 2, gn + (int64_t)row * p.inner + part * 2);
        }
        cp_async_commit();
    };
    pdl_wait();                                      // before the first access to the workspace (the table staging above reads constants)
    Item nxt; nxt.off = 0; nxt.rel = 0;
    if ((int64_t)blockIdx.x < p.nwork) {
        nxt = decode(blockIdx.x);
        prefetch(nxt);
    }
    for (int64_t w = blockIdx.x; w < p.nwork; w += gridDim.x) {
        const Item cur = nxt;
        cf *gt = p.ws + cur.off + c;
        pc v[E];
        cp_async_wait_all();
        __syncthreads();
        if (p.mode != 1) {
            // ---- forward: rows i + Tc j -> radix E -> twiddle -> exchange -> radix Tc -> rows q = (i + Tc m) + E k2 ----
#pragma unroll
            for (int j = 0; j < E; j++) v[j] = S[(i + Tc * j) * 8 + c];
            __syncthreads();
            pk::dft<false, E>(v);
            twiddle_store<E, false>(v, s_tw + i, Tc, S + i * 8 + c, pitch);
            __syncthreads();
#pragma unroll
            for (int m = 0; m < Mc; m++)
#pragma unroll
                for (int ii = 0; ii < Tc; ii++) v[m * Tc + ii] = S[(i + Tc * m) * pitch + ii * 8 + c];
#pragma unroll
            for (int m = 0; m < Mc; m++) pk::dft<false, Tc>(v + m * Tc);
            if (p.mode == 2) {
                const cf *kp = p.kspec + cur.rel + c;                           // same offset inside the kernel spectrum tile
#pragma unroll
                for (int m = 0; m < Mc; m++)
#pragma unroll
                    for (int k2 = 0; k2 < Tc; k2++) v[m * Tc + k2] = pk::cmul(v[m * Tc + k2], ld_pc_keep(kp + (int64_t)(i + Tc * m + E * k2) * p.inner, pol_keep));
            }
        } else {
#pragma unroll
            for (int m = 0; m < Mc; m++)
#pragma unroll
                for (int k2 = 0; k2 < Tc; k2++) v[m * Tc + k2] = S[(i + Tc * m + E * k2) * 8 + c];
        }
        if (p.mode == 0) {
            __syncthreads();                               // all exchange reads done: S may be restaged
            if (w + gridDim.x < p.nwork) { nxt = decode(w + gridDim.x); prefetch(nxt); }
#pragma unroll
            for (int m = 0; m < Mc; m++)
#pragma unroll
                for (int k2 = 0; k2 < Tc; k2++) st_pc(gt + (int64_t)(i + Tc * m + E * k2) * p.inner, v[m * Tc + k2]);
            continue;
        }
        if constexpr (Tc == E) {
            // square case: the rows this thread holds (i + E k2) are also the rows the forward-structured flow starts from,
            // so the inverse is the forward flow with conjugated twiddles (no transposed table needed)
            pk::dft<true, E>(v);
            __syncthreads();                               // every thread has finished reading S
            twiddle_store<E, true>(v, s_tw + i, Tc, S + i * 8 + c, pitch);
            __syncthreads();
#pragma unroll
            for (int ii = 0; ii < Tc; ii++) v[ii] = S[i * pitch + ii * 8 + c];
            __syncthreads();
        } else {
            // ---- inverse: radix Tc over k2 -> conj twiddle -> exchange -> radix E over k1 -> rows i + Tc j ----
#pragma unroll
            for (int m = 0; m < Mc; m++) pk::dft<true, Tc>(v + m * Tc);
            __syncthreads();                               // every thread has finished reading S
#pragma unroll
            for (int m = 0; m < Mc; m++)
                twiddle_store<Tc, true>(v + m * Tc, s_twT + (i + Tc * m), E, S + (i + Tc * m) * pitch + c, 8);
            __syncthreads();
            // after the exchange thread i owns "time" index i of every k1 row
#pragma unroll
            for (int k1 = 0; k1 < E; k1++) v[k1] = S[k1 * pitch + i * 8 + c];
            __syncthreads();
        }
        if (w + gridDim.x < p.nwork) { nxt = decode(w + gridDim.x); prefetch(nxt); }
        pk::dft<true, E>(v);
#pragma unroll
        for (int j = 0; j < E; j++) if (i + Tc * j >= p.skip) st_pc(gt + (int64_t)(i + Tc * j) * p.inner, v[j]);
    }
}

// ---- column pass, F = 1024, TMA-fed: one CTA per SM, two compute groups ----------------------------------------------------------
// col_pass above runs two CTAs per SM and stages the next item into its exchange buffer with cp.async, which is free only late in
// an item: the load latency of every item is exposed, 16 LDGSTS + 32 STG per thread carry 64-bit address arithmetic, and the
// register file (2 x 256 threads x 128) leaves no room for a third CTA.  col_pass_tma keeps the same butterflies but restructures
// the data movement around the TMA unit:
//   * ONE CTA of 512 threads per SM = two independent compute groups of 256 threads (named barriers 1 / 2), one twiddle table;
//   * a landing buffer P (64 KB) that `cp.async.bulk.tensor.2d` box loads (4 x [256 rows x 64 B]) fill while BOTH groups compute:
//     the group that has just drained P into registers immediately issues the load of the item the OTHER group takes next, so a full
//     item of look-ahead is always in flight and no thread ever waits for a load it has just issued;
//   * results leave through the group's (then idle) exchange buffer as dense [row][64 B] boxes and `cp.async.bulk.tensor.2d` stores
//     (rows [skip, F) only: the aliased head rows the crop discards are never written);
//   * no per-thread global address arithmetic for the workspace at all (SASS: UTMALDG / UTMASTG, mbarrier SYNCS).
// Shared memory: P 64 KB + 2 x 66 KB exchange + 8 KB twiddles = 204 KB.
struct ColTmaCfg {
    static constexpr int E = 32, Tc = 32, F = 1024, pitch = Tc * 8 + 8;
    static constexpr int group = 256, threads = 512;
    static constexpr int p_elems = F * 8, x_elems = E * pitch;                     // 8192, 8448 complex
    static constexpr int tw_pitch = 34;                                             // twiddle table rows of 32 + 2: 16-byte aligned, conflict-free 128-bit reads
    static constexpr int tw_elems = Tc * tw_pitch;
    static constexpr int smem = (p_elems + 2 * x_elems + tw_elems) * 8 + 64 + 128;  // + mbarriers + alignment slack
    static constexpr int box_rows = 256;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_group(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(ColTmaCfg::group) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ pc lds_pc(uint32_t a) { pc r; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r.v) : "r"(a)); return r; }
__device__ __forceinline__ void sts_pc(uint32_t a, pc v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v.v) : "memory"); }
__device__ __forceinline__ void lds_pc2(uint32_t a, pc &x, pc &y) { asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x.v), "=l"(y.v) : "r"(a)); }

// INNER > 0: the row pitch of the tile (complex elements) as a compile-time constant, so the kernel-spectrum loads address with
// immediates (1032 = the 2048-sample last-axis tile of the BASELINE workload); 0: read it from the parameters.
template <int INNER, bool TMA_STORE = true>
__global__ void __launch_bounds__(ColTmaCfg::threads, 1)
col_pass_tma(const __grid_constant__ ColParams p, const __grid_constant__ CUtensorMap tm_ld, const __grid_constant__ CUtensorMap tm_st)
{
    pdl_launch_dependents();
    using C = ColTmaCfg;
    constexpr int E = C::E, Tc = C::Tc, F = C::F, pitch = C::pitch;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, g = tid >> 8, lt = tid & 255;
    const int c = lt & 7, i = lt >> 3;                       // column of the block, thread index inside the column (< 32)
    const uint32_t sP = smem_addr(base);
    const uint32_t sX = sP + C::p_elems * 8 + g * (C::x_elems * 8);
    pc *s_tw = reinterpret_cast<pc *>(base + (C::p_elems + 2 * C::x_elems) * 8);      // s_tw[i * 34 + k1] = W_F^{i k1}: a thread reads its row two entries at a time
    const uint32_t sT = sP + (C::p_elems + 2 * C::x_elems) * 8 + i * (C::tw_pitch * 8);
    const uint32_t mb0 = sP + (C::p_elems + 2 * C::x_elems + C::tw_elems) * 8;         // full[0], full[1]: P holds an item for group 0 / 1
    const uint32_t mbg = mb0 + 8 * g;
    for (int idx = tid; idx < F; idx += C::threads) s_tw[(idx >> 5) * C::tw_pitch + (idx & 31)] = ld_pc(p.tw + (idx >> 5) * (idx & 31));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int64_t inner = INNER > 0 ? (int64_t)INNER : p.inner;
    const uint32_t iblocks = (uint32_t)(inner / 8), outer = (uint32_t)p.outer;
    // item w -> (q = tile * outer + o, 8-column block ib): rows [q F, q F + F) x columns [8 ib, 8 ib + 8) of the workspace seen as
    // one 2-D array of `inner` columns
    auto issue_load = [&](int64_t w, int gi) {
        const uint32_t w32 = (uint32_t)w, q = w32 / iblocks, ib = w32 - q * iblocks;
        const uint32_t mb = mb0 + 8 * gi;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(C::p_elems * 8)) : "memory");
#pragma unroll
        for (int b = 0; b < F / C::box_rows; b++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sP + b * C::box_rows * 64), "l"(reinterpret_cast<uint64_t>(&tm_ld)), "r"((int)(ib * 8)), "r"((int)(q * F + b * C::box_rows)), "r"(mb)
                         : "memory");
    };
    const uint64_t pol_keep = l2_policy_evict_last();
    pdl_wait();                                      // before the first access to the workspace
    if (tid == 0 && (int64_t)blockIdx.x < p.nwork) issue_load(blockIdx.x, 0);
    uint32_t parity = 0;
    const int nst = (F - p.skip) / p.store_rows;     // store boxes per item (host: store_rows divides F - skip; skip and store_rows even)
    for (int64_t n = g;; n += 2) {
        const int64_t w = (int64_t)blockIdx.x + n * gridDim.x;
        if (w >= p.nwork) break;
        const uint32_t w32 = (uint32_t)w, q = w32 / iblocks, ib = w32 - q * iblocks;
        pc v[32];
        mbar_wait(mbg, parity);
        parity ^= 1;
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = lds_pc(sP + ((i + 32 * j) * 8 + c) * 8);          // rows i + 32 j (both the forward's and, square case, the inverse's start)
        if (TMA_STORE && lt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous item's stores have read the exchange buffer
        // The first butterfly CONSUMES every loaded value before the barrier: a thread may arrive at a barrier with its shared-memory
        // loads still queued, and the TMA load issued after the barrier would overwrite P under them (seen on workspaces that fit
        // L2, where the box arrives within a few hundred cycles: the last rows of an item came out as the next item's).
        if (p.mode != 1) pk::dft<false, E>(v); else pk::dft<true, E>(v);
        bar_group(1 + g);
        if (lt == 0 && w + gridDim.x < p.nwork) issue_load(w + gridDim.x, g ^ 1);             // P is drained: the other group's next item
        if (p.mode != 1) {
#pragma unroll
            for (int k1 = 0; k1 < E; k1 += 2) {
                pc w0, w1; lds_pc2(sT + k1 * 8, w0, w1);
                sts_pc(sX + (k1 * pitch + i * 8 + c) * 8, pk::cmul(v[k1], w0));
                sts_pc(sX + ((k1 + 1) * pitch + i * 8 + c) * 8, pk::cmul(v[k1 + 1], w1));
            }
            bar_group(1 + g);
#pragma unroll
            for (int ii = 0; ii < Tc; ii++) v[ii] = lds_pc(sX + (i * pitch + ii * 8 + c) * 8);
            pk::dft<false, Tc>(v);                                                              // v[k2] = row i + 32 k2
            if (p.mode == 2) {
                const uint32_t o = q % outer;
                const cf *kp = p.kspec + ((int64_t)o * F * inner + ib * 8 + c) + (int64_t)i * inner;
#pragma unroll
                for (int k2 = 0; k2 < Tc; k2++) v[k2] = pk::cmul(v[k2], ld_pc_keep(kp + (int64_t)(E * k2) * inner, pol_keep));
                pk::dft<true, E>(v);
                bar_group(1 + g);                          // every thread has finished reading the exchange buffer
            }
        }
        if (p.mode != 0) {
            // square case: the rows this thread holds (i + 32 k2) are the rows the forward-structured flow starts from, so the
            // inverse is the forward flow with conjugated twiddles (mode INV: its first butterfly ran before the barrier above)
#pragma unroll
            for (int n1 = 0; n1 < E; n1 += 2) {
                pc w0, w1; lds_pc2(sT + n1 * 8, w0, w1);
                sts_pc(sX + (n1 * pitch + i * 8 + c) * 8, pk::cmulc(v[n1], w0));
                sts_pc(sX + ((n1 + 1) * pitch + i * 8 + c) * 8, pk::cmulc(v[n1 + 1], w1));
            }
            bar_group(1 + g);
#pragma unroll
            for (int ii = 0; ii < Tc; ii++) v[ii] = lds_pc(sX + (i * pitch + ii * 8 + c) * 8);
            pk::dft<true, E>(v);                           // v[j] = row i + 32 j
        }
        if constexpr (!TMA_STORE) {
            // A/B variant (NDCONV_COL_STG): results leave straight from the registers, 64-byte row segments per 8 lanes
            cf *gt = p.ws + ((int64_t)q * F + i) * inner + ib * 8 + c;
#pragma unroll
            for (int j = 0; j < 32; j++) if (i + 32 * j >= p.skip) st_pc(gt + (int64_t)(32 * j) * inner, v[j]);
            continue;                                      // (the exchange reads were consumed by the last butterfly; the next item's first barrier orders its writes)
        }
        bar_group(1 + g);                                  // exchange reads done: the buffer becomes the dense [row][8] store staging
#pragma unroll
        for (int j = 0; j < 32; j++) sts_pc(sX + ((i + 32 * j) * 8 + c) * 8, v[j]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_group(1 + g);
        if (lt == 0) {
            for (int b = 0; b < nst; b++) {
                const int r = p.skip + b * p.store_rows;
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(reinterpret_cast<uint64_t>(&tm_st)), "r"((int)(ib * 8)), "r"((int)(q * F + r)), "r"(sX + r * 64) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- column pass, F = 1024, mode FMI, kernel spectrum resident in Tensor Memory ----------------------------------------------------
// In col_pass_tma every item re-reads its 64 KB of kernel spectrum from L2 (32 LDG.64 per thread with no registers left to issue them
// early): measured 0.50 ms of the 2.5 ms pass on c5 (profiles/r02m_col_upper_bounds.log).  The spectrum of an 8-column block is the
// same for every tile, so this variant changes the work order -- a CTA stays on one block for a CHUNK of `bundle` (>= 10) tiles -- and
// keeps the block's spectrum in TENSOR MEMORY, which this FFT kernel has no other use for: each thread's 32 factors sit in its own
// TMEM lane (32x32b shape: lane = thread of the warp quadrant, 64 columns per thread; the two warps of a group that share a quadrant
// take columns [0, 64) / [64, 128) of a 128-column slot) and are read back with tcgen05.ld right where the multiply needs them:
// 41 cycles of latency instead of an L2 round trip, no shared-memory or L1 wavefronts, no address arithmetic.
//   * two slots (256 of the 512 columns): the current chunk's spectrum and the next chunk's, which arrives in eight 8 KB parts, one
//     per item: a 1-D bulk copy (cp.async.bulk, mbarrier-tracked) into the group's scratch buffer at the start of the item, then at
//     its end two LDS.128 + one tcgen05.st.x8 per thread.  `kres` holds the spectrum regrouped on the device once per kernel
//     (KresBody) so that a part is contiguous: kres[block][part d][half h][thread lt][2] = factors k2 = 4 d + 2 h + {0, 1} of thread lt.
//   * chunks of neighbouring CTAs are neighbouring blocks of the same tiles, so the DRAM page locality of the natural order is kept
//     (a CTA that stays on a block while its neighbours move on halves the bandwidth: tools/exp/colcopy_probe.cu).
//   * the two groups hand slots over through two counters in shared memory (`filled`: parts stored, `progress`: multiplies done);
//     by construction the waits never spin (a chunk is at least two items longer than the eight that carry parts).
// Work order: item j of CTA c (j = 0, 1, ...; group j & 1 takes it) is tile u * bundle + d of block kb, where chunk m = c + G * (j / len)
// = u * blocks + kb.  Everything else -- butterflies, TMA box loads into P, TMA stores from the exchange buffer -- is col_pass_tma.
struct ColKresCfg {
    static constexpr int parts = 8, part_elems = 1024;                                    // 8 parts of 256 threads x 4 factors
    static constexpr int scratch_off = (ColTmaCfg::p_elems + 2 * ColTmaCfg::x_elems + ColTmaCfg::tw_elems) * 8;
    static constexpr int mbar_off = scratch_off + 2 * part_elems * 8;                     // full[0], full[1], scr[0], scr[1], pro ; filled[2], progress[2], tmem base
    static constexpr int smem = mbar_off + 128 + 128;
    static constexpr int min_bundle = 10;
    static constexpr int tmem_cols = 512;            // two spectrum slots (2 x 128 columns) + the twiddle slot (128 columns), rounded up to a power of two
    static constexpr int tw_col = 256;
};
__device__ __forceinline__ void tm_ld32(uint32_t ta, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
                   "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(ta + 16) : "memory");
    // the wait names the registers so that no use of them can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                   "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
}
__device__ __forceinline__ void tm_ld16_issue(uint32_t ta, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tm_wait16(uint32_t *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                   "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t ta, const uint32_t *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t ta, uint4 a, uint4 b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(ta), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }
__device__ __forceinline__ int lds_acquire(uint32_t a) { int r; asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sts_release(uint32_t a, int v) { asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ pc pc_of(uint32_t lo, uint32_t hi) { pc r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(lo), "r"(hi)); return r; }

template <int INNER, bool TWT = true, bool TMA_STORE = true>
__global__ void __launch_bounds__(ColTmaCfg::threads, 1)
col_pass_tma_kres(const __grid_constant__ ColParams p, const __grid_constant__ CUtensorMap tm_ld, const __grid_constant__ CUtensorMap tm_st)
{
    pdl_launch_dependents();
    using C = ColTmaCfg;
    using K = ColKresCfg;
    constexpr int E = C::E, Tc = C::Tc, F = C::F, pitch = C::pitch;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, g = tid >> 8, lt = tid & 255, wq = (tid >> 5) & 3;
    const int c = lt & 7, i = lt >> 3;
    const uint32_t sP = smem_addr(base);
    const uint32_t sX = sP + C::p_elems * 8 + g * (C::x_elems * 8);
    pc *s_tw = reinterpret_cast<pc *>(base + (C::p_elems + 2 * C::x_elems) * 8);
    const uint32_t sT = sP + (C::p_elems + 2 * C::x_elems) * 8 + i * (C::tw_pitch * 8);
    const uint32_t sS = sP + K::scratch_off + g * (K::part_elems * 8);       // this group's scratch part
    const uint32_t mb0 = sP + K::mbar_off;                                    // full[0], full[1]
    const uint32_t mbg = mb0 + 8 * g, mbs = mb0 + 16 + 8 * g, mbp = mb0 + 32; // scr[g], prologue
    const uint32_t fl_filled = mb0 + 64, fl_progress = mb0 + 72, s_tbase = mb0 + 80;
    for (int idx = tid; idx < F; idx += C::threads) s_tw[(idx >> 5) * C::tw_pitch + (idx & 31)] = ld_pc(p.tw + (idx >> 5) * (idx & 31));
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 5; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb0 + 8 * b));
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(fl_filled), "r"(0) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if ((tid >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tbase), "r"((uint32_t)K::tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tbase; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tbase) : "r"(s_tbase));
    const uint32_t tq = tbase + ((uint32_t)(32 * wq) << 16) + (uint32_t)((lt >> 7) & 1) * 64;      // + slot * 128 + 2 * k2
    if constexpr (TWT) {
        // the thread's row of the twiddle table (W_F^{i k1}, k1 < 32) goes into Tensor Memory as well: 2 x 16 LDS.128 per item less
        // on the shared-memory pipe (group g stores k1 in [16 g, 16 g + 16); both groups read the same cells)
        uint32_t r[32];
#pragma unroll
        for (int e = 0; e < 16; e++) { const pc w = s_tw[i * C::tw_pitch + 16 * g + e]; r[2 * e] = (uint32_t)(w.v & 0xffffffffu); r[2 * e + 1] = (uint32_t)(w.v >> 32); }
        tm_st16(tq + K::tw_col + 32 * g, r);
        tm_st16(tq + K::tw_col + 32 * g + 16, r + 16);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // v[k1] * W^{i k1} (CONJ: its conjugate) -> exchange buffer row k1
    auto twiddle_exchange = [&](pc *v, auto conj_tag) {
        constexpr bool CONJ = decltype(conj_tag)::value;
        if constexpr (TWT) {
            uint32_t r[2][16];
            tm_ld16_issue(tq + K::tw_col, r[0]);
            tm_wait16(r[0]);
#pragma unroll
            for (int ch = 0; ch < 4; ch++) {
                if (ch < 3) tm_ld16_issue(tq + K::tw_col + 16 * (ch + 1), r[(ch + 1) & 1]);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int k1 = 8 * ch + e;
                    const pc w = pc_of(r[ch & 1][2 * e], r[ch & 1][2 * e + 1]);
                    sts_pc(sX + (k1 * pitch + i * 8 + c) * 8, CONJ ? pk::cmulc(v[k1], w) : pk::cmul(v[k1], w));
                }
                if (ch < 3) tm_wait16(r[(ch + 1) & 1]);
            }
        } else {
#pragma unroll
            for (int k1 = 0; k1 < E; k1 += 2) {
                pc w0, w1; lds_pc2(sT + k1 * 8, w0, w1);
                sts_pc(sX + (k1 * pitch + i * 8 + c) * 8, CONJ ? pk::cmulc(v[k1], w0) : pk::cmul(v[k1], w0));
                sts_pc(sX + ((k1 + 1) * pitch + i * 8 + c) * 8, CONJ ? pk::cmulc(v[k1 + 1], w1) : pk::cmul(v[k1 + 1], w1));
            }
        }
    };

    const int64_t inner = INNER > 0 ? (int64_t)INNER : p.inner;
    const uint32_t blocks = (uint32_t)(inner / 8), G = gridDim.x, cta = blockIdx.x;
    const int D = p.bundle, Dl = p.bundle_last;
    const uint32_t nq = (uint32_t)p.nbundles * blocks, nq_full = (uint32_t)(p.nbundles - 1) * blocks;     // chunks; chunks of full-length bundles
    const int Mc = cta < nq ? (int)((nq - cta + G - 1) / G) : 0, Mfull = cta < nq_full ? (int)((nq_full - cta + G - 1) / G) : 0;
    const int J = Mfull * D + (Mc - Mfull) * Dl;                               // items of this CTA
    auto chunk_len = [&](int m) { return m < Mfull ? D : Dl; };
    auto issue_load = [&](int m, int d, int gi) {
        if (NDC_EXP_FLAG(1)) return;
        const uint32_t q = cta + (uint32_t)m * G, u = q / blocks, ib = q - u * blocks, t = u * (uint32_t)D + (uint32_t)d;
        const uint32_t mb = mb0 + 8 * gi;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(C::p_elems * 8)) : "memory");
#pragma unroll
        for (int b = 0; b < F / C::box_rows; b++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sP + b * C::box_rows * 64), "l"(reinterpret_cast<uint64_t>(&tm_ld)), "r"((int)(ib * 8)), "r"((int)(NDC_RING_TILE(t) * F + b * C::box_rows)), "r"(mb)
                         : "memory");
    };
    pdl_wait();                                      // before the first access to the workspace (and to kres, which an earlier kernel of the stream wrote)
    if (Mc > 0) {
        // prologue: the first chunk's spectrum through P into slot 0 (all 512 threads: group g stores parts 4 g .. 4 g + 3)
        if (tid == 0) {
            const uint32_t kb = cta % blocks;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbp), "r"((uint32_t)(K::parts * K::part_elems * 8)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sP), "l"(p.kres + (size_t)kb * (K::parts * K::part_elems)), "r"((uint32_t)(K::parts * K::part_elems * 8)), "r"(mbp) : "memory");
        }
        mbar_wait(mbp, 0);
#pragma unroll
        for (int dd = 0; dd < 4; dd++) {
            const int d = 4 * g + dd;
            const uint4 a = lds128(sP + ((d * 2 + 0) * 256 + lt) * 16), b = lds128(sP + ((d * 2 + 1) * 256 + lt) * 16);
            tm_st8(tq + 8 * d, a, b);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                 // slot 0 complete; P read by everybody
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && J > 0) issue_load(0, 0, 0);
    uint32_t parity = 0, sparity = 0;
    const int nst = (F - p.skip) / p.store_rows;
    (void)nst;
    int m = 0, d = g, start = 0, m_checked = 0;      // chunk, position in it, CTA item index of the chunk's first item; the newest chunk whose slot this thread knows complete
    while (m < Mc && d >= chunk_len(m)) { d -= chunk_len(m); start += chunk_len(m); m++; }
    for (int j = g; j < J; j += 2) {
        const uint32_t q = cta + (uint32_t)m * G, u = q / blocks, ib = q - u * blocks, t = u * (uint32_t)D + (uint32_t)d;
        const bool carries_part = d < K::parts && m + 1 < Mc;      // group-uniform
        pc v[32];
        if (!NDC_EXP_FLAG(1)) mbar_wait(mbg, parity);
        parity ^= 1;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) v[jj] = lds_pc(sP + ((i + 32 * jj) * 8 + c) * 8);
        if (lt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        pk::dft<false, E>(v);                        // consumes every loaded value before the barrier (see col_pass_tma)
        bar_group(1 + g);
        if (lt == 0) {
            if (j + 1 < J) { int m1 = m, d1 = d + 1; if (d1 >= chunk_len(m)) { d1 = 0; m1++; } issue_load(m1, d1, g ^ 1); }
            if (carries_part) {                      // part d of the next chunk's spectrum -> this group's scratch
                const uint32_t kb = (q + G) % blocks;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbs), "r"((uint32_t)(K::part_elems * 8)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sS), "l"(p.kres + ((size_t)kb * K::parts + d) * K::part_elems), "r"((uint32_t)(K::part_elems * 8)), "r"(mbs) : "memory");
            }
            sts_release(fl_filled + 4 * g, j);       // every earlier item of this group has stored its part (all threads passed the barrier above)
        }
        twiddle_exchange(v, std::false_type{});
        bar_group(1 + g);
#pragma unroll
        for (int ii = 0; ii < Tc; ii++) v[ii] = lds_pc(sX + (i * pitch + ii * 8 + c) * 8);
        pk::dft<false, Tc>(v);                       // v[k2] = row i + 32 k2
        if (m > m_checked) {
            // first item of this group in chunk m: the other group's parts of it are stored once it has published an item index
            // above the chunk's last part-carrying item (start - len(m - 1) + 7)
            const int need = start - chunk_len(m - 1) + K::parts;
            while (lds_acquire(fl_filled + 4 * (g ^ 1)) < need) { }
            m_checked = m;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const uint32_t ts = tq + (uint32_t)(m & 1) * 128;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t r[32];
                tm_ld32(ts + 32 * h, r);
#pragma unroll
                for (int e = 0; e < 16; e++) v[16 * h + e] = pk::cmul(v[16 * h + e], pc_of(r[2 * e], r[2 * e + 1]));
            }
        }
        pk::dft<true, E>(v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        bar_group(1 + g);                            // every thread has finished reading the exchange buffer (and this item's slot)
        if (lt == 0) sts_release(fl_progress + 4 * g, j + 1);
        twiddle_exchange(v, std::true_type{});
        bar_group(1 + g);
#pragma unroll
        for (int ii = 0; ii < Tc; ii++) v[ii] = lds_pc(sX + (i * pitch + ii * 8 + c) * 8);
        pk::dft<true, E>(v);                         // v[jj] = row i + 32 jj
        if constexpr (!TMA_STORE) {
            // A/B variant (NDCONV_COL_STG): the results leave straight from the registers as 64-byte row segments per 8 lanes, so the TMA
            // unit moves the loads only (the exchange reads above were consumed by the last butterfly; the next item's first barrier orders
            // its writes to the exchange buffer behind them)
            cf *gt = p.ws + ((int64_t)NDC_RING_TILE(t) * F + i) * inner + ib * 8 + c;
#pragma unroll
            for (int jj = 0; jj < 32; jj++) if (i + 32 * jj >= p.skip) st_pc(gt + (int64_t)(32 * jj) * inner, v[jj]);
        } else {
        bar_group(1 + g);                            // exchange reads done: the buffer becomes the dense [row][8] store staging
#pragma unroll
        for (int jj = 0; jj < 32; jj++) sts_pc(sX + ((i + 32 * jj) * 8 + c) * 8, v[jj]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_group(1 + g);
        if (lt == 0 && !NDC_EXP_FLAG(2)) {
            for (int b = 0; b < nst; b++) {
                const int r = p.skip + b * p.store_rows;
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(reinterpret_cast<uint64_t>(&tm_st)), "r"((int)(ib * 8)), "r"((int)(NDC_RING_TILE(t) * F + r)), "r"(sX + r * 64) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        }
        if (carries_part) {
            // scratch -> the other slot, which chunk m - 1 was read from: the other group must be past the multiply of its last item there
            mbar_wait(mbs, sparity);
            sparity ^= 1;
            const uint4 a = lds128(sS + lt * 16), b = lds128(sS + (256 + lt) * 16);
            if (m > 0) while (lds_acquire(fl_progress + 4 * (g ^ 1)) < start - 1) { }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tm_st8(tq + (uint32_t)((m + 1) & 1) * 128 + 8 * d, a, b);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        d += 2;
        while (m < Mc && d >= chunk_len(m)) { d -= chunk_len(m); start += chunk_len(m); m++; }
    }
    if (lt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((uint32_t)K::tmem_cols) : "memory");
}

}  // namespace fast

// kernel spectrum [rows][Hp] in natural bin order (bins 0..L, generic path) -> the fast path's row layout (pitch L + 8):
// column 0 <- bin 0, column 1 <- bin L/2, columns (2s, 2s+1) <- bins (s, L-s), column L <- bin L (Nyquist), L+1.. <- 0.
struct KfastParams {
    const cx<float> *kspec;
    cx<float> *kfast;
    int64_t rows;
    int L, Hp;
    int is_cx;        // Complex<f32> problems: rows of exactly L columns in natural bin order (no pairing)
};
struct KfastBody {
    static HD void run(const BlockCtx &c, const KfastParams &p)
    {
        const int pitch = p.is_cx ? p.L : p.L + fast::kPad;
        const int64_t total = p.rows * pitch;
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int64_t q = e / pitch;
            const int pc = (int)(e % pitch);
            cx<float> val = cx<float>{0.f, 0.f};
            if (p.is_cx) val = p.kspec[q * p.Hp + pc];
            else if (pc <= p.L) {
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
// fast-path kernel spectrum of a 1024-row axis-0 tile [1024][inner] -> the order col_pass_tma_kres streams into Tensor Memory:
// kres[block][part d < 8][half h < 2][thread lt < 256][e < 2] = kfast[(i + 32 k2) * inner + 8 block + c], lt = 8 i + c, k2 = 4 d + 2 h + e
struct KresParams {
    const cx<float> *kfast;
    cx<float> *kres;
    int64_t inner;
};
struct KresBody {
    static HD void run(const BlockCtx &c, const KresParams &p)
    {
        const int64_t total = 1024 * p.inner;
        for (int64_t e = c.bid * c.nt + c.tid; e < total; e += c.nb * c.nt) {
            const int64_t blk = e >> 13;
            const int r = (int)(e & 8191), d = r >> 10, h = (r >> 9) & 1, lt = (r >> 1) & 255, ee = r & 1;
            const int i = lt >> 3, cc = lt & 7, k2 = 4 * d + 2 * h + ee;
            p.kres[e] = p.kfast[(int64_t)(i + 32 * k2) * p.inner + blk * 8 + cc];
        }
    }
};
}  // namespace ndc
#endif  // NDCONV_CUDA
