// experiment: what bandwidth does the column kernel's ACCESS PATTERN allow?  Each work item is F = 1024 rows x W bytes (one row
// segment per row, row pitch 8256 bytes) of a tile; items are dealt to a persistent grid exactly like col_pass does.  The kernel
// only copies (load 16-byte chunks to registers, store them back in place), so the number is the ceiling for any column kernel
// with that segment width and that many bytes in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 colcopy_probe.cu -o colcopy_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// ORDER 0: items dealt round-robin, column block fastest (neighbouring CTAs touch neighbouring segments of the same rows at the same time)
// ORDER 1: every CTA owns a contiguous range of (column block, tile) items, column block slowest (a CTA stays on one column block)
template <int W, int MODE, int ORDER = 0>    // W: segment bytes; MODE 0 read+write, 1 read only, 2 write only
__global__ void __launch_bounds__(256) colcopy(uint4 *ws, int64_t tile_bytes, int64_t pitch_bytes, int iblocks, int64_t nwork, uint4 *sink)
{
    constexpr int F = 1024, CPR = W / 16, CH = F * CPR / 256;     // chunks per row, chunks per thread
    uint4 acc = make_uint4(0, 0, 0, 0);
    const int64_t w_begin = ORDER ? nwork * blockIdx.x / gridDim.x : blockIdx.x, w_end = ORDER ? nwork * (blockIdx.x + 1) / gridDim.x : nwork;
    const uint32_t ntiles = (uint32_t)(nwork / iblocks);
    for (int64_t w = w_begin; w < w_end; w += ORDER ? 1 : gridDim.x) {
        const uint32_t w32 = (uint32_t)w;
        const uint32_t tile = ORDER ? w32 % ntiles : w32 / iblocks, ib = ORDER ? w32 / ntiles : w32 - tile * iblocks;
        char *base = reinterpret_cast<char *>(ws) + tile * tile_bytes + (int64_t)ib * W;
        constexpr int B = CH < 16 ? CH : 16;                       // 16 chunks (256 bytes) in flight per thread, like the real kernel
        for (int b0 = 0; b0 < CH; b0 += B) {
            uint4 v[B];
#pragma unroll
            for (int m = 0; m < B; m++) {
                const int id = threadIdx.x + 256 * (b0 + m), row = id / CPR, part = id % CPR;
                if (MODE != 2) v[m] = *reinterpret_cast<const uint4 *>(base + row * pitch_bytes + part * 16);
                else v[m] = make_uint4(id, w32, 0, 0);
            }
#pragma unroll
            for (int m = 0; m < B; m++) {
                const int id = threadIdx.x + 256 * (b0 + m), row = id / CPR, part = id % CPR;
                if (MODE != 1) { v[m].x += 1; *reinterpret_cast<uint4 *>(base + row * pitch_bytes + part * 16) = v[m]; }
                else { acc.x ^= v[m].x; acc.y ^= v[m].y; acc.z ^= v[m].z; acc.w ^= v[m].w; }
            }
        }
    }
    if (MODE == 1 && acc.x == 0x12345678u) sink[threadIdx.x] = acc;
}

template <int W, int MODE, int ORDER = 0> static void run(uint4 *ws, int ntiles, int64_t tile_bytes, int64_t pitch_bytes, int nsm, uint4 *sink)
{
    const int iblocks = (int)(pitch_bytes / W);
    const int64_t nwork = (int64_t)ntiles * iblocks;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int per_sm : {1, 2, 4}) {
        const int grid = nsm * per_sm;
        colcopy<W, MODE, ORDER><<<grid, 256>>>(ws, tile_bytes, pitch_bytes, iblocks, nwork, sink);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            colcopy<W, MODE, ORDER><<<grid, 256>>>(ws, tile_bytes, pitch_bytes, iblocks, nwork, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        const double bytes = (double)nwork * 1024 * W * (MODE == 0 ? 2 : 1);
        printf("order %d segment %3d B  %s  %d CTAs/SM: %.3f ms  %.0f GB/s\n", ORDER, W, MODE == 0 ? "read+write" : MODE == 1 ? "read only " : "write only", per_sm, best,
               bytes / (best * 1e-3) / 1e9);
    }
}

int main(int argc, char **argv)
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int nsm = pr.multiProcessorCount;
    const int ntiles = 300;                                   // ~2.5 GB
    uint4 *ws, *sink; cudaMalloc(&ws, (size_t)ntiles * 1024 * 16384); cudaMalloc(&sink, 4096);
    cudaMemset(ws, 0, (size_t)ntiles * 1024 * 16384);
    if (argc > 1) {
        // row pitch sweep: does the mapping of the 64-byte segments of consecutive rows onto DRAM channels / banks matter?
        const int64_t pitches[] = {8256, 8320, 8448, 8704, 9216, 10240, 12288, 16384, 8192};
        for (int64_t pb : pitches) {
            printf("pitch %lld bytes: ", (long long)pb);
            const int iblocks = 129;
            const int64_t nwork = (int64_t)ntiles * iblocks, tile_bytes = 1024 * pb;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            colcopy<64, 0, 0><<<nsm * 2, 256>>>(ws, tile_bytes, pb, iblocks, nwork, sink);
            cudaDeviceSynchronize();
            float best = 1e30f;
            for (int r = 0; r < 3; r++) {
                cudaEventRecord(e0);
                colcopy<64, 0, 0><<<nsm * 2, 256>>>(ws, tile_bytes, pb, iblocks, nwork, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
            }
            printf("64-byte segments, read+write, 2 CTAs/SM: %.3f ms  %.0f GB/s\n", best, (double)nwork * 1024 * 64 * 2 / (best * 1e-3) / 1e9);
        }
        return 0;
    }
    const int64_t pitch_bytes = 1032 * 8, tile_bytes = 1024 * pitch_bytes;
    printf("%s, %d SMs; %d tiles of 1024 rows x %lld bytes\n", pr.name, nsm, ntiles, (long long)pitch_bytes);
    run<64, 0>(ws, ntiles, tile_bytes, pitch_bytes, nsm, sink);
    run<64, 0, 1>(ws, ntiles, tile_bytes, pitch_bytes, nsm, sink);
    run<128, 0, 1>(ws, ntiles, tile_bytes, pitch_bytes, nsm, sink);
    run<64, 1, 1>(ws, ntiles, tile_bytes, pitch_bytes, nsm, sink);
    run<64, 2, 1>(ws, ntiles, tile_bytes, pitch_bytes, nsm, sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
