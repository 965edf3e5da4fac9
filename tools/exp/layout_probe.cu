// experiment: a column-blocked workspace layout  ws[tile][col block][row][8 complex]  (64 KB contiguous per work item of the
// column kernel) against today's row-major tiles (64-byte segments at row pitch).  Pure data movement, no arithmetic:
//   col   : in-place read+write of one 64 KB item per CTA iteration   (row-major: 1024 segments of 64 B at pitch 8256)
//   rowW  : warp per row, 8 KB sequential read from x, one row of the tile written (row-major: 8256 B contiguous;
//           blocked: 129 segments of 64 B at stride 64 KB)
//   rowR  : the reverse (read a tile row, write 8 KB sequentially)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 layout_probe.cu -o layout_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int F = 1024, NCB = 129;                       // rows per tile, 64-byte column blocks per row
constexpr int64_t TILE_BYTES = (int64_t)F * NCB * 64;

__device__ __forceinline__ int64_t seg_off(bool blocked, int row, int cb) { return blocked ? ((int64_t)cb * F + row) * 64 : ((int64_t)row * NCB + cb) * 64; }

template <bool BLOCKED> __global__ void __launch_bounds__(256) col_k(char *ws, int64_t nwork)
{
    for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
        const uint32_t tile = (uint32_t)w / NCB, cb = (uint32_t)w % NCB;
        char *base = ws + tile * TILE_BYTES;
        uint4 v[16];
#pragma unroll
        for (int m = 0; m < 16; m++) { const int id = threadIdx.x + 256 * m; v[m] = *reinterpret_cast<const uint4 *>(base + seg_off(BLOCKED, id >> 2, cb) + (id & 3) * 16); }
#pragma unroll
        for (int m = 0; m < 16; m++) { const int id = threadIdx.x + 256 * m; v[m].x += 1; *reinterpret_cast<uint4 *>(base + seg_off(BLOCKED, id >> 2, cb) + (id & 3) * 16) = v[m]; }
    }
}
// MODE 0: x -> tile row (row_fwd-like), 1: tile row -> out (row_inv-like)
template <bool BLOCKED, int MODE> __global__ void __launch_bounds__(128, 4) row_k(char *ws, char *lin, int64_t nrows)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * 4 + warp; r < nrows; r += (int64_t)gridDim.x * 4) {
        const uint32_t tile = (uint32_t)r / F, row = (uint32_t)r % F;
        char *tb = ws + tile * TILE_BYTES, *lr = lin + r * 8192;
        uint4 v[16];
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = *reinterpret_cast<const uint4 *>(lr + (lane + 32 * j) * 16);
#pragma unroll
            for (int j = 0; j < 16; j++) { const int k = lane + 32 * j; *reinterpret_cast<uint4 *>(tb + seg_off(BLOCKED, row, k >> 2) + (k & 3) * 16) = v[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) { const int k = lane + 32 * j; v[j] = *reinterpret_cast<const uint4 *>(tb + seg_off(BLOCKED, row, k >> 2) + (k & 3) * 16); }
#pragma unroll
            for (int j = 0; j < 16; j++) *reinterpret_cast<uint4 *>(lr + (lane + 32 * j) * 16) = v[j];
        }
    }
}
template <class L> static float best_ms(L f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int nsm = pr.multiProcessorCount, ntiles = 300;
    char *ws, *lin; cudaMalloc(&ws, ntiles * TILE_BYTES); cudaMalloc(&lin, (size_t)ntiles * F * 8192);
    cudaMemset(ws, 0, ntiles * TILE_BYTES); cudaMemset(lin, 0, (size_t)ntiles * F * 8192);
    const int64_t nwork = (int64_t)ntiles * NCB, nrows = (int64_t)ntiles * F;
    const double colb = (double)nwork * 65536 * 2, rowb = (double)nrows * 2 * 8192;
    float t;
    t = best_ms([&] { col_k<false><<<nsm * 2, 256>>>(ws, nwork); });        printf("col  row-major tiles : %.3f ms  %.0f GB/s\n", t, colb / t / 1e6);
    t = best_ms([&] { col_k<true><<<nsm * 2, 256>>>(ws, nwork); });         printf("col  column-blocked  : %.3f ms  %.0f GB/s\n", t, colb / t / 1e6);
    t = best_ms([&] { row_k<false, 0><<<nsm * 32, 128>>>(ws, lin, nrows); }); printf("rowW row-major tiles : %.3f ms  %.0f GB/s\n", t, rowb / t / 1e6);
    t = best_ms([&] { row_k<true, 0><<<nsm * 32, 128>>>(ws, lin, nrows); });  printf("rowW column-blocked  : %.3f ms  %.0f GB/s\n", t, rowb / t / 1e6);
    t = best_ms([&] { row_k<false, 1><<<nsm * 32, 128>>>(ws, lin, nrows); }); printf("rowR row-major tiles : %.3f ms  %.0f GB/s\n", t, rowb / t / 1e6);
    t = best_ms([&] { row_k<true, 1><<<nsm * 32, 128>>>(ws, lin, nrows); });  printf("rowR column-blocked  : %.3f ms  %.0f GB/s\n", t, rowb / t / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
