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
// The compacted tap list (gen_offset_list, src/dilation/mod.rs:34-60) lives in shared memory as (tile offset, weight);
// each thread keeps TO0 accumulators in registers and walks the taps in the reference's order with un-fused
// multiply/add, so results stay bit-identical (integers wrap, floats round identically).
#pragma once
#include "kernels_direct.h"

#ifdef NDCONV_CUDA
#include <cuda.h>

namespace ndc {
namespace tile {

constexpr int kThreads = 256;
constexpr int kMaxTO0 = 4;

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
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <class T> __device__ __forceinline__ T tile_padded_at(const TileParams &p, const int64_t *c)
{
    const T *x = (const T *)p.x;
    int64_t src = 0;
    bool init = false;
#pragma unroll
    for (int a = 2; a >= 0; a--) {
        if (c[a] >= p.P[a]) return Elem<T>::zero();     // only the 16-byte rounding of the window can reach past the padded extent
        const int32_t m = p.map[a][c[a]];
        if (m >= 0) src += (int64_t)m * p.xstr[a];
        else if (m == NDC_MAP_INIT) init = true;
        else return *(const T *)(m == NDC_MAP_CONST_FRONT ? p.cfront[a] : p.cback[a]);
    }
    return init ? Elem<T>::zero() : x[src];
}

template <class T>
__global__ void __launch_bounds__(kThreads) direct_tile_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams p)
{
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
    const int lane = tid & 31, warp = tid >> 5;
    const int nrows = p.IT[0] * p.IT[1];
    const int64_t c2_lo = c_lo[2] - shift2;

    // value of the padded array along one tile row for i2 in [lo, hi): the two outer axes are resolved once per row
    // (highest-numbered constant axis wins, never-written cells read 0), lanes sweep the contiguous axis
    auto fill_row = [&](int row, int lo, int hi, int lo2, int hi2) {      // fills [lo,hi) and [lo2,hi2) of the row
        const T *x = (const T *)p.x;
        const int i0 = row / p.IT[1], i1 = row - i0 * p.IT[1];
        const int64_t c0 = c_lo[0] + i0, c1 = c_lo[1] + i1;
        T *trow = tile + row * row_elems;
        bool zero = c0 >= p.P[0] || c1 >= p.P[1], has_const = false;
        const unsigned char *cval = nullptr;
        int64_t base = 0;
        if (!zero) {
            const int32_t m1 = p.map[1][c1];
            if (m1 >= 0) base += (int64_t)m1 * p.xstr[1];
            else if (m1 == NDC_MAP_INIT) zero = true;
            else { has_const = true; cval = (m1 == NDC_MAP_CONST_FRONT) ? p.cfront[1] : p.cback[1]; }
            if (!has_const) {
                const int32_t m0 = p.map[0][c0];
                if (m0 >= 0) base += (int64_t)m0 * p.xstr[0];
                else if (m0 == NDC_MAP_INIT) zero = true;
                else { has_const = true; cval = (m0 == NDC_MAP_CONST_FRONT) ? p.cfront[0] : p.cback[0]; }
            }
        }
        const int n1 = hi - lo, ntot = n1 + (hi2 - lo2);
#pragma unroll 2
        for (int e = lane; e < ntot; e += 32) {
            const int i2 = e < n1 ? lo + e : lo2 + (e - n1);
            const int64_t c2 = c2_lo + i2;
            T val = Elem<T>::zero();
            if (c2 >= 0 && c2 < p.P[2] && !(c0 >= p.P[0] || c1 >= p.P[1])) {
                const int32_t m2 = p.map[2][c2];
                if (m2 == NDC_MAP_CONST_FRONT) val = *(const T *)p.cfront[2];
                else if (m2 == NDC_MAP_CONST_BACK) val = *(const T *)p.cback[2];
                else if (has_const) val = *(const T *)cval;
                else if (m2 != NDC_MAP_INIT && !zero) val = x[base + (int64_t)m2 * p.xstr[2]];
            }
            trow[i2] = val;
        }
    };

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
    } else {
        for (int row = warp; row < nrows; row += kThreads / 32) fill_row(row, 0, row_elems, 0, 0);
    }
    // taps -> shared memory: (offset inside the tile, weight), reference order
    for (int t = tid; t < p.ntap; t += kThreads) {
        const int32_t *o = p.tap_off + t * NDC_MAX_DIM;
        int off = 0;
        if (p.axis_shift <= 0) off += o[0 - p.axis_shift] * plane_elems;
        if (p.axis_shift <= 1) off += o[1 - p.axis_shift] * row_elems;
        off += o[2 - p.axis_shift];
        s_off[t] = off;
        s_w[t] = ((const T *)p.tap_w)[t];
    }
    __syncthreads();
    if (tma_ok) {
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
            // halo patch.  (1) fringes: rows inside the array on the outer axes only leave it along axis 2 -- one thread per
            // (row, fringe element), so the per-row map lookups run 32 rows at a time; (2) rows that lie outside the array on
            // an outer axis are rebuilt whole, one warp per row.
            const int f2 = (int)min((int64_t)row_elems, max((int64_t)0, p.pf[2] - c2_lo));                 // [0, f2) is front halo
            const int b2 = (int)min((int64_t)row_elems, max((int64_t)0, p.pf[2] + p.n[2] - c2_lo));       // [b2, row_elems) is back halo
            const int nh = f2 + (row_elems - b2);
            for (int e = tid; e < nrows * nh; e += kThreads) {
                const int row = e / nh, h = e - row * nh;
                const int i2 = h < f2 ? h : b2 + (h - f2);
                const int i0 = row / p.IT[1], i1 = row - i0 * p.IT[1];
                const int64_t c[3] = {c_lo[0] + i0, c_lo[1] + i1, c2_lo + i2};
                const bool outer = c[0] < p.pf[0] || c[0] >= p.pf[0] + p.n[0] || c[1] < p.pf[1] || c[1] >= p.pf[1] + p.n[1];
                if (!outer) tile[row * row_elems + i2] = c[2] < 0 ? Elem<T>::zero() : tile_padded_at<T>(p, c);
            }
            for (int row = warp; row < nrows; row += kThreads / 32) {
                const int i0 = row / p.IT[1], i1 = row - i0 * p.IT[1];
                const int64_t c0 = c_lo[0] + i0, c1 = c_lo[1] + i1;
                const bool outer = c0 < p.pf[0] || c0 >= p.pf[0] + p.n[0] || c1 < p.pf[1] || c1 >= p.pf[1] + p.n[1];
                if (outer) fill_row(row, 0, row_elems, 0, 0);
            }
            __syncthreads();
        }
    }
    // compute: thread -> (o1, o2) of the tile, TO0 accumulators in registers
    const int o2l = tid % p.TO[2], o1l = tid / p.TO[2];
    if (o1l < p.TO[1]) {
        const int64_t o1 = (int64_t)t1 * p.TO[1] + o1l, o2 = (int64_t)t2 * p.TO[2] + o2l;
        T acc[kMaxTO0];
#pragma unroll
        for (int q = 0; q < kMaxTO0; q++) acc[q] = Elem<T>::zero();
        const int base = o1l * (int)p.s[1] * row_elems + o2l * (int)p.s[2] + shift2;
        const int step0 = (int)p.s[0] * plane_elems;
        for (int t = 0; t < p.ntap; t++) {
            const T w = s_w[t];
            const int off = base + s_off[t];
#pragma unroll
            for (int q = 0; q < kMaxTO0; q++)
                if (q < p.TO[0]) acc[q] = Elem<T>::mac(acc[q], tile[off + q * step0], w);
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

}  // namespace tile
}  // namespace ndc
#endif
