// kernels_direct_tile.cuh -- sm_100a direct convolution: TMA-staged shared-memory tile plus halo (device only).
//
// ConvExt::conv (src/conv/mod.rs:128-200) for rank <= 3 (lower ranks are embedded with leading extents of 1).
// One CTA computes a TO0 x TO1 x TO2 block of outputs.  Its input window ((TO-1)*stride + Kd per axis, last axis
// rounded up to 16 bytes) is staged in shared memory:
//   * interior tiles, and tiles whose overhang falls in a Zeros border: ONE cp.async.bulk.tensor.3d (TMA) box load,
//     out-of-bounds elements zero-filled by the hardware, completion on an mbarrier.  Measured on B200
//     (tools/exp/tma_probe.cu): with SWIZZLE_NONE the box must START on a 16-byte boundary of the innermost axis
//     (coordinate * sizeof(T) % 16 == 0) or UTMALDG raises an illegal-instruction fault, so the box start is aligned
//     down and the window is read with a per-tile shift;
//   * edge tiles with a non-zero border (Const / Reflect / Replicate / Circular): the same TMA load brings the in-bounds
//     part; only the halo elements that fall outside the array are then patched through the border index maps
//     (index arithmetic, a few % of the tile) -- the padded array of src/padding/mod.rs:84-117 is never materialised.
//     Without TMA (1/2/16-byte elements, unaligned rows) the whole window is gathered through the maps.
// The compacted tap list (gen_offset_list, src/dilation/mod.rs:34-60) lives in shared memory as (byte offset in the tile, weight);
// each thread keeps TO0 accumulators in registers and walks the taps in the reference's order with un-fused
// multiply/add, so results stay bit-identical (integers wrap, floats round identically).
#pragma once
#include "kernels_direct.h"
#include <climits>

#ifdef NDCONV_CUDA
#include <cuda.h>

namespace ndc {
namespace tile {

constexpr int kThreads = 256;
constexpr int kMaxTO0 = 4;
constexpr int kBlockR = 4;               // register-blocked variant: outputs per thread along the contiguous axis
constexpr int kRowTaps = 8;              //   ... and the longest kernel row it takes

struct TileParams {
    // geometry embedded in 3-D (axis 2 = contiguous)
    int64_t n[3], xstr[3], P[3], pf[3], Kd[3], s[3], O[3], ostr[3];
    const int32_t *map[3];
    unsigned char cfront[3][16], cback[3][16];
    int front_zero[3], back_zero[3];     // the border on that side reads as 0 everywhere (Zeros, or Const(0))
    int TO[3], IT[3], IT2p, ntile[3];
    int use_tma;
    int ntap;
    const int32_t *tap_off;              // [ntap][NDC_MAX_DIM] dilated offsets (axes of the ORIGINAL rank, right-aligned below)
    int axis_shift;                      // 3 - ndim
    const void *tap_w;
    const void *x;
    void *out;
    int tile_elems;
    // register-blocked variant (direct_tile_kernel<T, S2, D2>): kernel extents and dilations (embedded in 3-D) and the kernel rows
    int kk[3], dd[3];
    int nrow;                            // kk[0] * kk[1] kernel rows of up to kRowTaps taps along the contiguous axis
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile element at a 32-bit shared-memory address.  The tap loop addresses the tile as base + offset with the base converted ONCE:
// through a generic pointer ptxas re-derived the shared window (S2UR CgaCtaId / ULEA) for every access -- 9 instructions per
// multiply-add instead of 3 (ncu source page, c4).
template <int BYTES> struct LdsRaw;
template <> struct LdsRaw<1> { typedef uint8_t V; static __device__ __forceinline__ V ld(uint32_t a) { uint16_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=h"(v) : "r"(a)); return (V)v; } };
template <> struct LdsRaw<2> { typedef uint16_t V; static __device__ __forceinline__ V ld(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; } };
template <> struct LdsRaw<4> { typedef uint32_t V; static __device__ __forceinline__ V ld(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; } };
template <> struct LdsRaw<8> { typedef uint64_t V; static __device__ __forceinline__ V ld(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; } };
template <> struct LdsRaw<16> { typedef ulonglong2 V; static __device__ __forceinline__ V ld(uint32_t a) { ulonglong2 v; asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a)); return v; } };
template <class T> __device__ __forceinline__ T lds_elem(uint32_t a)
{
    typename LdsRaw<sizeof(T)>::V raw = LdsRaw<sizeof(T)>::ld(a);
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
}

// the tap loop: NQ accumulators (outputs along axis 0) per thread, taps in the reference's order, un-fused multiply/add
template <class T, int NQ>
__device__ __forceinline__ void tap_loop(uint32_t tile_addr, const T *s_w, const int32_t *s_off, int ntap, int step0_bytes, T *acc)
{
#pragma unroll
    for (int q = 0; q < NQ; q++) acc[q] = Elem<T>::zero();
#pragma unroll 4
    for (int t = 0; t < ntap; t++) {
        const T w = s_w[t];
        uint32_t a = tile_addr + (uint32_t)s_off[t];
#pragma unroll
        for (int q = 0; q < NQ; q++) { acc[q] = Elem<T>::mac(acc[q], lds_elem<T>(a), w); a += (uint32_t)step0_bytes; }
    }
}


// Register-blocked tap loop (stride S2 and dilation D2 of the contiguous axis are compile-time, 1 or 2): a thread owns NQ x R
// outputs (NQ along axis 0, R = 4 neighbours along the contiguous axis).  Per kernel ROW it loads, for each of its NQ planes, the
// (R - 1) S2 + (k2 - 1) D2 + 1 samples its R outputs share ONCE into registers, as 16-byte vectors (consecutive lanes read consecutive
// 16- or 32-byte chunks: no bank conflicts; scalar reads at that lane stride were 4- to 8-way conflicted and SLOWER than the plain
// tap loop), and the row's weights once -- instead of one shared-memory read per multiply-add.  SHIFT is the tile's alignment shift
// (the TMA box starts on a 16-byte boundary, kernels_direct_tile.cuh header), a compile-time constant here so that the samples keep
// static register indices.  Every output still sees its taps in the reference's order (rows ascending, taps of a row ascending, zero
// weights skipped: gen_offset_list, src/dilation/mod.rs:34-60) with un-fused multiply / add, so results stay bit-identical.
template <class T, int NQ, int S2, int D2, int SHIFT>
__device__ __forceinline__ void blocked_loop(uint32_t tile_addr16, const T *s_rw, const int32_t *s_roff, const uint32_t *s_rmask, int nrow, int k2, int step0_bytes,
                                             T (*acc)[kBlockR])
{
    constexpr int R = kBlockR, VN = 16 / (int)sizeof(T), SEG = (R - 1) * S2 + (kRowTaps - 1) * D2 + 1, NV = (SHIFT + SEG + VN - 1) / VN;
    const int nv = (SHIFT + (R - 1) * S2 + (k2 - 1) * D2 + 1 + VN - 1) / VN;
#pragma unroll
    for (int q = 0; q < NQ; q++)
#pragma unroll
        for (int j = 0; j < R; j++) acc[q][j] = Elem<T>::zero();
    for (int r = 0; r < nrow; r++) {
        const uint32_t mask = s_rmask[r];
        if (!mask) continue;                                  // a kernel row of zero weights
        T w[kRowTaps];
#pragma unroll
        for (int v = 0; v < kRowTaps; v++) w[v] = s_rw[r * kRowTaps + v];
        uint32_t a = tile_addr16 + (uint32_t)s_roff[r];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            T buf[NV * VN];
#pragma unroll
            for (int c = 0; c < NV; c++) {
                ulonglong2 raw = make_ulonglong2(0ull, 0ull);
                if (c < nv) raw = LdsRaw<16>::ld(a + 16u * (uint32_t)c);
                memcpy(&buf[c * VN], &raw, 16);
            }
            // taps in ascending order.  4-byte elements: the tail groups [4, 6) and [6, 8) sit behind uniform branches on the row length, so
            // a 5-tap row issues 24 multiply-adds instead of 32 with a third of them predicated off -- predicated-off instructions still
            // take their issue slots, and this loop is bound by issue: 256x1024x1024 i32 k = 3x5x5 2.41 -> 2.14 ms,
            // f32 k = 3x3x3 1.79 -> 1.34 ms, bit-identical; 7-tap rows pay the two branches (8192^2 k = 7x7 f32 616 -> 633 us).  Measured and
            // rejected: three straight-line blocks of 4 / 6 / 8 taps (i32 k = 5: 2.49 ms), and the branches for 8-byte elements (+3 %).
            auto tap = [&](int v) {
                if ((mask >> v) & 1u) {
#pragma unroll
                    for (int j = 0; j < R; j++) acc[q][j] = Elem<T>::mac(acc[q][j], buf[SHIFT + j * S2 + v * D2], w[v]);
                }
            };
            static_assert(kRowTaps == 8, "tap groups below");
            if constexpr (sizeof(T) == 4) {
                tap(0); tap(1); tap(2); tap(3);
                if (k2 > 4) { tap(4); tap(5); }
                if (k2 > 6) { tap(6); tap(7); }
            } else {
#pragma unroll
                for (int v = 0; v < kRowTaps; v++) tap(v);
            }
            a += (uint32_t)step0_bytes;
        }
    }
}
template <class T, int NQ, int S2, int D2>
__device__ __forceinline__ void blocked_loop_shift(int shift, uint32_t tile_addr16, const T *s_rw, const int32_t *s_roff, const uint32_t *s_rmask, int nrow, int k2,
                                                   int step0_bytes, T (*acc)[kBlockR])
{
    if constexpr (sizeof(T) == 8) {
        if (shift == 0) blocked_loop<T, NQ, S2, D2, 0>(tile_addr16, s_rw, s_roff, s_rmask, nrow, k2, step0_bytes, acc);
        else blocked_loop<T, NQ, S2, D2, 1>(tile_addr16, s_rw, s_roff, s_rmask, nrow, k2, step0_bytes, acc);
    } else {
        switch (shift) {
        case 0: blocked_loop<T, NQ, S2, D2, 0>(tile_addr16, s_rw, s_roff, s_rmask, nrow, k2, step0_bytes, acc); break;
        case 1: blocked_loop<T, NQ, S2, D2, 1>(tile_addr16, s_rw, s_roff, s_rmask, nrow, k2, step0_bytes, acc); break;
        case 2: blocked_loop<T, NQ, S2, D2, 2>(tile_addr16, s_rw, s_roff, s_rmask, nrow, k2, step0_bytes, acc); break;
        default: blocked_loop<T, NQ, S2, D2, 3>(tile_addr16, s_rw, s_roff, s_rmask, nrow, k2, step0_bytes, acc); break;
        }
    }
}

// S2 == 0: one output column per thread, the compacted tap list (any stride / dilation / kernel extent); S2 > 0: the register-blocked rows above
template <class T, int S2 = 0, int D2 = 0>
__global__ void __launch_bounds__(kThreads) direct_tile_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams p)
{
    // programmatic dependent launch (see kernels_fft_fast.cuh): the next kernel of the stream may start now; this one stages
    // its window maps (plan constants) and only then waits for its predecessors, before the first read of x / write of out
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);
    const int tile_bytes = (p.tile_elems * (int)sizeof(T) + 127) & ~127;
    T *s_w = reinterpret_cast<T *>(smem_raw + tile_bytes);
    int32_t *s_off = reinterpret_cast<int32_t *>(smem_raw + tile_bytes + ((p.ntap * (int)sizeof(T) + 15) & ~15));
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x;
    int tb = blockIdx.x;
    const int t2 = tb % p.ntile[2]; tb /= p.ntile[2];
    const int t1 = tb % p.ntile[1]; tb /= p.ntile[1];
    const int t0 = tb;
    const int tt[3] = {t0, t1, t2};
    int64_t c_lo[3];
    const bool tma_ok = p.use_tma != 0;
    bool need_patch = false;                                // the window overhangs a border that does not read as zero
#pragma unroll
    for (int a = 0; a < 3; a++) {
        c_lo[a] = (int64_t)tt[a] * p.TO[a] * p.s[a];
        const int64_t hi = c_lo[a] + p.IT[a];               // the window the outputs of this tile actually read
        if (c_lo[a] < p.pf[a] && !p.front_zero[a]) need_patch = true;
        if (hi > p.pf[a] + p.n[a] && !p.back_zero[a]) need_patch = true;
    }
    // 16-byte alignment of the box start along the contiguous axis
    constexpr int kAlign = sizeof(T) >= 16 ? 1 : 16 / (int)sizeof(T);
    const int64_t cx2_raw = c_lo[2] - p.pf[2];
    const int shift2 = (int)(((cx2_raw % kAlign) + kAlign) % kAlign);
    const int row_elems = p.IT2p, plane_elems = p.IT[1] * p.IT2p;
    const int64_t c2_lo = c_lo[2] - shift2;

    // window maps in shared memory: s_map[a][i] = border map of padded coordinate c_lo[a] + i (kBeyond outside [0, P[a]))
    constexpr int32_t kBeyond = INT32_MIN;
    // (staged after the TMA box has been issued, and only by the tiles that gather or patch through them: the dependent global loads of the
    // maps used to sit in front of every tile's box load)
    int32_t *s_map0 = s_off + p.ntap, *s_map1 = s_map0 + p.IT[0], *s_map2 = s_map1 + p.IT[1];
    auto stage_maps = [&]() {
        for (int e = tid; e < p.IT[0] + p.IT[1] + row_elems; e += kThreads) {
            const int a = e < p.IT[0] ? 0 : (e < p.IT[0] + p.IT[1] ? 1 : 2);
            const int i = e - (a == 0 ? 0 : (a == 1 ? p.IT[0] : p.IT[0] + p.IT[1]));
            const int64_t c = (a == 2 ? c2_lo : c_lo[a]) + i;
            s_map0[e] = (c < 0 || c >= p.P[a]) ? kBeyond : p.map[a][c];
        }
    };
    if (!tma_ok) stage_maps();
    // value of the padded array at window element (i0, i1, i2): the highest-numbered constant axis wins, never-written cells
    // and the 16-byte rounding slack read 0 (same precedence as the sequential padding of src/padding/mod.rs:119-153)
    auto resolve = [&](int i0, int i1, int i2) -> T {
        const int32_t m2 = s_map2[i2], m1 = s_map1[i1], m0 = s_map0[i0];
        if (m2 == kBeyond || m1 == kBeyond || m0 == kBeyond) return Elem<T>::zero();
        if (m2 == NDC_MAP_CONST_FRONT) return *(const T *)p.cfront[2];
        if (m2 == NDC_MAP_CONST_BACK) return *(const T *)p.cback[2];
        if (m1 == NDC_MAP_CONST_FRONT) return *(const T *)p.cfront[1];
        if (m1 == NDC_MAP_CONST_BACK) return *(const T *)p.cback[1];
        if (m0 == NDC_MAP_CONST_FRONT) return *(const T *)p.cfront[0];
        if (m0 == NDC_MAP_CONST_BACK) return *(const T *)p.cback[0];
        if (m2 == NDC_MAP_INIT || m1 == NDC_MAP_INIT || m0 == NDC_MAP_INIT) return Elem<T>::zero();
        return ((const T *)p.x)[(int64_t)m0 * p.xstr[0] + (int64_t)m1 * p.xstr[1] + (int64_t)m2 * p.xstr[2]];
    };
    // fill `nq * width` window rows, row (q, j) -> (i0, i1) by `rowmap`, elements h in [0, nsel) -> i2 by `colmap`.  The whole CTA
    // works on it, `rows_it` rows per iteration, one element per thread and iteration, so the gathers of an iteration are
    // independent loads (a warp per row with a map lookup feeding each load serialised ~20 us of latency on edge tiles of c4)
    auto fill_rows = [&](int nq, int width, int nsel, auto rowmap, auto colmap) {
        if (nq <= 0 || width <= 0 || nsel <= 0) return;
        int lanes = 1;
        while (lanes < nsel && lanes < kThreads) lanes <<= 1;
        const int rows_it = kThreads / lanes, sub = tid / lanes, ln = tid - sub * lanes;
        const int total = nq * width;
        int q = sub / width, j = sub - q * width;
#pragma unroll 4
        for (int r = sub; r < total; r += rows_it) {
            int i0, i1;
            rowmap(q, j, i0, i1);
            for (int h = ln; h < nsel; h += lanes) {
                const int i2 = colmap(h);
                tile[(i0 * p.IT[1] + i1) * row_elems + i2] = resolve(i0, i1, i2);
            }
            j += rows_it;
            while (j >= width) { j -= width; q++; }
        }
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tma_ok) {
        if (tid == 0) {
            const uint32_t mb = smem_u32(&mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = (uint32_t)(p.tile_elems * (int)sizeof(T));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            const int cx2 = (int)(cx2_raw - shift2), cx1 = (int)(c_lo[1] - p.pf[1]), cx0 = (int)(c_lo[0] - p.pf[0]);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(cx2), "r"(cx1), "r"(cx0), "r"(mb)
                : "memory");
        }
    }
    if (tma_ok && need_patch) stage_maps();
    // taps -> shared memory: (byte offset inside the tile, weight), reference order
    for (int t = tid; t < p.ntap; t += kThreads) {
        const int32_t *o = p.tap_off + t * NDC_MAX_DIM;
        int off = 0;
        if (p.axis_shift <= 0) off += o[0 - p.axis_shift] * plane_elems;
        if (p.axis_shift <= 1) off += o[1 - p.axis_shift] * row_elems;
        off += o[2 - p.axis_shift];
        s_off[t] = off * (int)sizeof(T);           // byte offset inside the tile
        s_w[t] = ((const T *)p.tap_w)[t];
    }
    // register-blocked variant: the kernel as dense rows (row = position on the two outer axes; tap v of a row sits at dilated offset v * D2)
    int32_t *s_roff = s_map2 + row_elems;
    uint32_t *s_rmask = reinterpret_cast<uint32_t *>(s_roff + p.nrow);
    T *s_rw = reinterpret_cast<T *>((reinterpret_cast<uintptr_t>(s_rmask + p.nrow) + 15) & ~(uintptr_t)15);
    if constexpr (S2 > 0) {
        for (int r = tid; r < p.nrow; r += kThreads) {
            s_roff[r] = ((r / p.kk[1]) * p.dd[0] * plane_elems + (r % p.kk[1]) * p.dd[1] * row_elems) * (int)sizeof(T);
            s_rmask[r] = 0u;
        }
        for (int e = tid; e < p.nrow * kRowTaps; e += kThreads) s_rw[e] = Elem<T>::zero();
        __syncthreads();
        for (int t = tid; t < p.ntap; t += kThreads) {
            const int32_t *o = p.tap_off + t * NDC_MAX_DIM;
            const int u0 = p.axis_shift <= 0 ? o[0 - p.axis_shift] / p.dd[0] : 0, u1 = p.axis_shift <= 1 ? o[1 - p.axis_shift] / p.dd[1] : 0, v = o[2 - p.axis_shift] / p.dd[2];
            const int r = u0 * p.kk[1] + u1;
            s_rw[r * kRowTaps + v] = ((const T *)p.tap_w)[t];
            atomicOr(&s_rmask[r], 1u << v);
        }
    }
    __syncthreads();
    auto ident_col = [&](int h) { return h; };
    if (!tma_ok) {
        // no TMA (1/2/16-byte elements, unaligned rows, strided views): gather the whole window through the maps
        fill_rows(p.IT[0], p.IT[1], row_elems, [&](int q, int j, int &i0, int &i1) { i0 = q; i1 = j; }, ident_col);
        __syncthreads();
    } else {
        const uint32_t mb = smem_u32(&mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(mb), "r"(0u) : "memory");
        }
        if (need_patch) {
            // halo patch: only the window elements that fall outside the array are rebuilt through the maps.  Rows
            // [lo0, hi0) x [lo1, hi1) lie inside the array on the outer axes and only need their axis-2 fringes [0, f2) and
            // [b2, row_elems); the rows above (i0 < lo0), below (i0 >= hi0) and beside them (i1 outside [lo1, hi1)) are rebuilt whole.
            const int lo0 = (int)min((int64_t)p.IT[0], max((int64_t)0, p.pf[0] - c_lo[0])), hi0 = (int)max((int64_t)lo0, min((int64_t)p.IT[0], p.pf[0] + p.n[0] - c_lo[0]));
            const int lo1 = (int)min((int64_t)p.IT[1], max((int64_t)0, p.pf[1] - c_lo[1])), hi1 = (int)max((int64_t)lo1, min((int64_t)p.IT[1], p.pf[1] + p.n[1] - c_lo[1]));
            const int f2 = (int)min((int64_t)row_elems, max((int64_t)0, p.pf[2] - c2_lo));                 // [0, f2) is front halo
            const int b2 = (int)max((int64_t)f2, min((int64_t)row_elems, max((int64_t)0, p.pf[2] + p.n[2] - c2_lo)));       // [b2, row_elems) is back halo
            const int wside = lo1 + (p.IT[1] - hi1);
            fill_rows(lo0, p.IT[1], row_elems, [&](int q, int j, int &i0, int &i1) { i0 = q; i1 = j; }, ident_col);
            fill_rows(p.IT[0] - hi0, p.IT[1], row_elems, [&](int q, int j, int &i0, int &i1) { i0 = hi0 + q; i1 = j; }, ident_col);
            fill_rows(hi0 - lo0, wside, row_elems, [&](int q, int j, int &i0, int &i1) { i0 = lo0 + q; i1 = j < lo1 ? j : hi1 + (j - lo1); }, ident_col);
            fill_rows(hi0 - lo0, hi1 - lo1, f2 + (row_elems - b2), [&](int q, int j, int &i0, int &i1) { i0 = lo0 + q; i1 = lo1 + j; },
                      [&](int h) { return h < f2 ? h : b2 + (h - f2); });
            __syncthreads();
        }
    }
    // compute
    if constexpr (S2 > 0) {
        // thread -> (o1, four neighbouring o2) of the tile, TO0 x 4 accumulators in registers
        const int T2 = p.TO[2] / kBlockR;
        const int o2l = (tid % T2) * kBlockR, o1l = tid / T2;
        if (o1l < p.TO[1]) {
            const int64_t o1 = (int64_t)t1 * p.TO[1] + o1l, o2 = (int64_t)t2 * p.TO[2] + o2l;
            T acc[kMaxTO0][kBlockR];
            const int base = o1l * (int)p.s[1] * row_elems + o2l * S2;          // 16-byte aligned; the tile's alignment shift is a template constant of the loop
            const uint32_t tile_addr16 = smem_u32(tile) + (uint32_t)base * (uint32_t)sizeof(T);
            const int step0_bytes = (int)p.s[0] * plane_elems * (int)sizeof(T);
            switch (p.TO[0]) {                                                   // the host picks 1, 2 or 4 outputs along axis 0 for this variant
            case 1: blocked_loop_shift<T, 1, S2, D2>(shift2, tile_addr16, s_rw, s_roff, s_rmask, p.nrow, p.kk[2], step0_bytes, acc); break;
            case 2: blocked_loop_shift<T, 2, S2, D2>(shift2, tile_addr16, s_rw, s_roff, s_rmask, p.nrow, p.kk[2], step0_bytes, acc); break;
            default: blocked_loop_shift<T, 4, S2, D2>(shift2, tile_addr16, s_rw, s_roff, s_rmask, p.nrow, p.kk[2], step0_bytes, acc); break;
            }
            if (o1 < p.O[1]) {
                T *out = (T *)p.out;
#pragma unroll
                for (int q = 0; q < kMaxTO0; q++) {
                    const int64_t o0 = (int64_t)t0 * p.TO[0] + q;
                    if (q < p.TO[0] && o0 < p.O[0]) {
                        T *orow = out + o0 * p.ostr[0] + o1 * p.ostr[1] + o2;
#pragma unroll
                        for (int j = 0; j < kBlockR; j++) if (o2 + j < p.O[2]) orow[j] = acc[q][j];
                    }
                }
            }
        }
    } else {
        // thread -> (o1, o2) of the tile, TO0 accumulators in registers
        const int o2l = tid % p.TO[2], o1l = tid / p.TO[2];
        if (o1l < p.TO[1]) {
            const int64_t o1 = (int64_t)t1 * p.TO[1] + o1l, o2 = (int64_t)t2 * p.TO[2] + o2l;
            T acc[kMaxTO0];
            const int base = o1l * (int)p.s[1] * row_elems + o2l * (int)p.s[2] + shift2;
            const uint32_t tile_addr = smem_u32(tile) + (uint32_t)base * (uint32_t)sizeof(T);
            const int step0_bytes = (int)p.s[0] * plane_elems * (int)sizeof(T);
            switch (p.TO[0]) {
            case 1: tap_loop<T, 1>(tile_addr, s_w, s_off, p.ntap, step0_bytes, acc); break;
            case 2: tap_loop<T, 2>(tile_addr, s_w, s_off, p.ntap, step0_bytes, acc); break;
            case 3: tap_loop<T, 3>(tile_addr, s_w, s_off, p.ntap, step0_bytes, acc); break;
            default: tap_loop<T, 4>(tile_addr, s_w, s_off, p.ntap, step0_bytes, acc); break;
            }
            if (o1 < p.O[1] && o2 < p.O[2]) {
                T *out = (T *)p.out;
#pragma unroll
                for (int q = 0; q < kMaxTO0; q++) {
                    const int64_t o0 = (int64_t)t0 * p.TO[0] + q;
                    if (q < p.TO[0] && o0 < p.O[0]) out[o0 * p.ostr[0] + o1 * p.ostr[1] + o2] = acc[q];
                }
            }
        }
    }
}


// ---- persistent variant: one CTA walks many tiles, two window buffers ---------------------------------------------------------------
// In the throughput regime (65 536 tiles of a 256 x 1024 x 1024 array) one CTA per tile pays the whole latency chain per tile -- window
// maps, TMA box, tap staging, barrier, compute, store -- with four CTAs per SM to overlap it: the large 3-D shapes ran at 0.7-1.2 TB/s of
// compulsory bytes while neither HBM nor the multiply-adds were busy.  Here a CTA stages the kernel rows ONCE, then loops over tiles
// n = blockIdx.x, + gridDim.x, ...: the TMA box of tile n + gridDim.x is issued into the other buffer before tile n is computed
// (mbarrier per buffer), the window's border maps are staged only for the edge tiles that need a halo patch.  Register-blocked compute
// (blocked_loop) only; everything else about a tile is direct_tile_kernel's.
template <class T, int S2, int D2>
__global__ void __launch_bounds__(kThreads) direct_tile_persistent(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams p, int ntiles_total)
{
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_bytes = (p.tile_elems * (int)sizeof(T) + 127) & ~127;
    unsigned char *tables = smem_raw + 2 * tile_bytes;
    int32_t *s_map0 = reinterpret_cast<int32_t *>(tables), *s_map1 = s_map0 + p.IT[0], *s_map2 = s_map1 + p.IT[1];
    const int row_elems = p.IT2p, plane_elems = p.IT[1] * p.IT2p;
    int32_t *s_roff = s_map2 + row_elems;
    uint32_t *s_rmask = reinterpret_cast<uint32_t *>(s_roff + p.nrow);
    T *s_rw = reinterpret_cast<T *>((reinterpret_cast<uintptr_t>(s_rmask + p.nrow) + 15) & ~(uintptr_t)15);
    __shared__ __align__(8) uint64_t mbar[2];
    const int tid = threadIdx.x;
    constexpr int32_t kBeyond = INT32_MIN;
    constexpr int kAlign = 16 / (int)sizeof(T);

    // kernel rows (plan constants: before the wait on the predecessors)
    for (int r = tid; r < p.nrow; r += kThreads) {
        s_roff[r] = ((r / p.kk[1]) * p.dd[0] * plane_elems + (r % p.kk[1]) * p.dd[1] * row_elems) * (int)sizeof(T);
        s_rmask[r] = 0u;
    }
    for (int e = tid; e < p.nrow * kRowTaps; e += kThreads) s_rw[e] = Elem<T>::zero();
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    for (int t = tid; t < p.ntap; t += kThreads) {
        const int32_t *o = p.tap_off + t * NDC_MAX_DIM;
        const int u0 = p.axis_shift <= 0 ? o[0 - p.axis_shift] / p.dd[0] : 0, u1 = p.axis_shift <= 1 ? o[1 - p.axis_shift] / p.dd[1] : 0, v = o[2 - p.axis_shift] / p.dd[2];
        const int r = u0 * p.kk[1] + u1;
        s_rw[r * kRowTaps + v] = ((const T *)p.tap_w)[t];
        atomicOr(&s_rmask[r], 1u << v);
    }
    struct TileGeo { int t0, t1, t2, shift2; int64_t c0, c1, c2, c2_lo; bool need_patch; };
    auto geo = [&](int n) {
        TileGeo g;
        int tb = n;
        g.t2 = tb % p.ntile[2]; tb /= p.ntile[2];
        g.t1 = tb % p.ntile[1]; tb /= p.ntile[1];
        g.t0 = tb;
        g.c0 = (int64_t)g.t0 * p.TO[0] * p.s[0]; g.c1 = (int64_t)g.t1 * p.TO[1] * p.s[1]; g.c2 = (int64_t)g.t2 * p.TO[2] * p.s[2];
        const int64_t c[3] = {g.c0, g.c1, g.c2};
        g.need_patch = false;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (c[a] < p.pf[a] && !p.front_zero[a]) g.need_patch = true;
            if (c[a] + p.IT[a] > p.pf[a] + p.n[a] && !p.back_zero[a]) g.need_patch = true;
        }
        const int64_t cx2_raw = g.c2 - p.pf[2];
        g.shift2 = (int)(((cx2_raw % kAlign) + kAlign) % kAlign);
        g.c2_lo = g.c2 - g.shift2;
        return g;
    };
    auto issue = [&](const TileGeo &g, int buf) {      // thread 0 only
        const uint32_t mb = smem_u32(&mbar[buf]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // a halo patch may have written this buffer through the generic proxy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(p.tile_elems * (int)sizeof(T))) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(smem_raw + buf * tile_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((int)(g.c2_lo - p.pf[2])), "r"((int)(g.c1 - p.pf[1])), "r"((int)(g.c0 - p.pf[0])), "r"(mb)
            : "memory");
    };
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0 && (int)blockIdx.x < ntiles_total) issue(geo(blockIdx.x), 0);
    __syncthreads();                                    // kernel rows complete
    int it = 0;
    for (int n = blockIdx.x; n < ntiles_total; n += gridDim.x, it++) {
        const int buf = it & 1;
        const TileGeo g = geo(n);
        T *tile = reinterpret_cast<T *>(smem_raw + buf * tile_bytes);
        // the next tile's box into the other buffer (every thread left it at the barrier that ended the previous iteration)
        if (tid == 0 && n + (int)gridDim.x < ntiles_total) issue(geo(n + gridDim.x), buf ^ 1);
        if (g.need_patch) {
            const int64_t c_lo[3] = {g.c0, g.c1, g.c2};
            for (int e = tid; e < p.IT[0] + p.IT[1] + row_elems; e += kThreads) {
                const int a = e < p.IT[0] ? 0 : (e < p.IT[0] + p.IT[1] ? 1 : 2);
                const int i = e - (a == 0 ? 0 : (a == 1 ? p.IT[0] : p.IT[0] + p.IT[1]));
                const int64_t c = (a == 2 ? g.c2_lo : c_lo[a]) + i;
                s_map0[e] = (c < 0 || c >= p.P[a]) ? kBeyond : p.map[a][c];
            }
            __syncthreads();
        }
        {
            const uint32_t mb = smem_u32(&mbar[buf]), parity = (uint32_t)((it >> 1) & 1);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(mb), "r"(parity) : "memory");
            }
        }
        if (g.need_patch) {
            auto resolve = [&](int i0, int i1, int i2) -> T {
                const int32_t m2 = s_map2[i2], m1 = s_map1[i1], m0 = s_map0[i0];
                if (m2 == kBeyond || m1 == kBeyond || m0 == kBeyond) return Elem<T>::zero();
                if (m2 == NDC_MAP_CONST_FRONT) return *(const T *)p.cfront[2];
                if (m2 == NDC_MAP_CONST_BACK) return *(const T *)p.cback[2];
                if (m1 == NDC_MAP_CONST_FRONT) return *(const T *)p.cfront[1];
                if (m1 == NDC_MAP_CONST_BACK) return *(const T *)p.cback[1];
                if (m0 == NDC_MAP_CONST_FRONT) return *(const T *)p.cfront[0];
                if (m0 == NDC_MAP_CONST_BACK) return *(const T *)p.cback[0];
                if (m2 == NDC_MAP_INIT || m1 == NDC_MAP_INIT || m0 == NDC_MAP_INIT) return Elem<T>::zero();
                return ((const T *)p.x)[(int64_t)m0 * p.xstr[0] + (int64_t)m1 * p.xstr[1] + (int64_t)m2 * p.xstr[2]];
            };
            auto fill_rows = [&](int nq, int width, int nsel, auto rowmap, auto colmap) {
                if (nq <= 0 || width <= 0 || nsel <= 0) return;
                int lanes = 1;
                while (lanes < nsel && lanes < kThreads) lanes <<= 1;
                const int rows_it = kThreads / lanes, sub = tid / lanes, ln = tid - sub * lanes;
                const int total = nq * width;
                int q = sub / width, j = sub - q * width;
#pragma unroll 4
                for (int r = sub; r < total; r += rows_it) {
                    int i0, i1;
                    rowmap(q, j, i0, i1);
                    for (int h = ln; h < nsel; h += lanes) {
                        const int i2 = colmap(h);
                        tile[(i0 * p.IT[1] + i1) * row_elems + i2] = resolve(i0, i1, i2);
                    }
                    j += rows_it;
                    while (j >= width) { j -= width; q++; }
                }
            };
            auto ident_col = [&](int h) { return h; };
            const int lo0 = (int)min((int64_t)p.IT[0], max((int64_t)0, p.pf[0] - g.c0)), hi0 = (int)max((int64_t)lo0, min((int64_t)p.IT[0], p.pf[0] + p.n[0] - g.c0));
            const int lo1 = (int)min((int64_t)p.IT[1], max((int64_t)0, p.pf[1] - g.c1)), hi1 = (int)max((int64_t)lo1, min((int64_t)p.IT[1], p.pf[1] + p.n[1] - g.c1));
            const int f2 = (int)min((int64_t)row_elems, max((int64_t)0, p.pf[2] - g.c2_lo));
            const int b2 = (int)max((int64_t)f2, min((int64_t)row_elems, max((int64_t)0, p.pf[2] + p.n[2] - g.c2_lo)));
            const int wside = lo1 + (p.IT[1] - hi1);
            fill_rows(lo0, p.IT[1], row_elems, [&](int q, int j, int &i0, int &i1) { i0 = q; i1 = j; }, ident_col);
            fill_rows(p.IT[0] - hi0, p.IT[1], row_elems, [&](int q, int j, int &i0, int &i1) { i0 = hi0 + q; i1 = j; }, ident_col);
            fill_rows(hi0 - lo0, wside, row_elems, [&](int q, int j, int &i0, int &i1) { i0 = lo0 + q; i1 = j < lo1 ? j : hi1 + (j - lo1); }, ident_col);
            fill_rows(hi0 - lo0, hi1 - lo1, f2 + (row_elems - b2), [&](int q, int j, int &i0, int &i1) { i0 = lo0 + q; i1 = lo1 + j; },
                      [&](int h) { return h < f2 ? h : b2 + (h - f2); });
            __syncthreads();
        }
        // compute: thread -> (o1, four neighbouring o2) of the tile, TO0 x 4 accumulators in registers
        {
            const int T2 = p.TO[2] / kBlockR;
            const int o2l = (tid % T2) * kBlockR, o1l = tid / T2;
            if (o1l < p.TO[1]) {
                const int64_t o1 = (int64_t)g.t1 * p.TO[1] + o1l, o2 = (int64_t)g.t2 * p.TO[2] + o2l;
                T acc[kMaxTO0][kBlockR];
                const int base = o1l * (int)p.s[1] * row_elems + o2l * S2;
                const uint32_t tile_addr16 = smem_u32(tile) + (uint32_t)base * (uint32_t)sizeof(T);
                const int step0_bytes = (int)p.s[0] * plane_elems * (int)sizeof(T);
                switch (p.TO[0]) {
                case 1: blocked_loop_shift<T, 1, S2, D2>(g.shift2, tile_addr16, s_rw, s_roff, s_rmask, p.nrow, p.kk[2], step0_bytes, acc); break;
                case 2: blocked_loop_shift<T, 2, S2, D2>(g.shift2, tile_addr16, s_rw, s_roff, s_rmask, p.nrow, p.kk[2], step0_bytes, acc); break;
                default: blocked_loop_shift<T, 4, S2, D2>(g.shift2, tile_addr16, s_rw, s_roff, s_rmask, p.nrow, p.kk[2], step0_bytes, acc); break;
                }
                if (o1 < p.O[1]) {
                    T *out = (T *)p.out;
#pragma unroll
                    for (int q = 0; q < kMaxTO0; q++) {
                        const int64_t o0 = (int64_t)g.t0 * p.TO[0] + q;
                        if (q < p.TO[0] && o0 < p.O[0]) {
                            T *orow = out + o0 * p.ostr[0] + o1 * p.ostr[1] + o2;
#pragma unroll
                            for (int j = 0; j < kBlockR; j++) if (o2 + j < p.O[2]) orow[j] = acc[q][j];
                        }
                    }
                }
            }
        }
        __syncthreads();                                // this buffer (and the window maps) may be overwritten from the next iteration on
    }
}

}  // namespace tile
}  // namespace ndc
#endif
