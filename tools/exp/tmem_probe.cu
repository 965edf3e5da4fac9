// experiment: Tensor Memory as a thread-private constant store for a non-MMA kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_probe.cu -o tmem_probe
// (1) correctness of the addressing the column pass would use: 512 threads = 2 groups of 256; group g fills "slot" g (128 columns:
//     warp quadrant = lanes, the two warps of a group that share a quadrant take 64 columns each) with tcgen05.st.32x32b; after a
//     fence + barrier EVERY thread reads BOTH slots with tcgen05.ld.32x32b (a warp of the other group reads what its twin wrote).
// (2) cost: cycles per 64 x 32-bit values per thread, all 16 warps busy: tcgen05.ld.x16 x 4 versus 32 x ld.shared.b64.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void tm_st16(uint32_t ta, const uint32_t *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t ta, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t expect(int slot, int lt, int col) { return 0x9e3779b9u * (uint32_t)(slot * 65536 + lt * 64 + col) + 12345u; }

__global__ void __launch_bounds__(512, 1) probe(unsigned *errs, long long *cyc, int iters, float *sink)
{
    __shared__ uint32_t tbase;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int tid = threadIdx.x, w = tid >> 5, g = tid >> 8, lt = tid & 255;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tbase;
    const uint32_t lane_base = (uint32_t)(32 * (w & 3)) << 16;
    const uint32_t colw = (uint32_t)((lt >> 7) & 1) * 64;           // which of the two warps of the group on this quadrant
    // fill slot g
    for (int ch = 0; ch < 4; ch++) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = expect(g, lt, ch * 16 + j);
        tm_st16(base + lane_base + g * 128 + colw + ch * 16, r);
    }
    tm_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned bad = 0;
    for (int slot = 0; slot < 2; slot++)
        for (int ch = 0; ch < 4; ch++) {
            uint32_t r[16];
            tm_ld16(base + lane_base + slot * 128 + colw + ch * 16, r);
            tm_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j++) bad += (r[j] != expect(slot, lt, ch * 16 + j));
        }
    if (bad) atomicAdd(errs, bad);
    if (tid == 0 && blockIdx.x == 0) errs[1] = base;
    __syncthreads();
    // ---- timing: TMEM reads
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int ch = 0; ch < 4; ch += 2) {
            uint32_t r[32];
            tm_ld16(base + lane_base + (it & 1) * 128 + colw + ch * 16, r);
            tm_ld16(base + lane_base + (it & 1) * 128 + colw + ch * 16 + 16, r + 16);
            tm_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; j++) acc += __uint_as_float(r[j] & 0x3fffffffu);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    // ---- timing: the same volume from shared memory (32 x ld.shared.b64 per thread, conflict-free)
    uint2 *sm = reinterpret_cast<uint2 *>(dyn);
    for (int j = 0; j < 32; j++) sm[j * 512 + tid] = make_uint2(tid + j, tid - j);
    __syncthreads();
    long long t2 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
            uint2 v = sm[((j + it) & 31) * 512 + tid];
            acc += __uint_as_float(v.x & 0x3fffffffu) + __uint_as_float(v.y & 0x3fffffffu);
        }
    }
    long long t3 = clock64();
    __syncthreads();
    // ---- both at once: warps of group 0 read TMEM while group 1 reads shared memory (do the two paths overlap?)
    long long t4 = clock64();
    for (int it = 0; it < iters; it++) {
        if (g == 0) {
#pragma unroll
            for (int ch = 0; ch < 4; ch += 2) {
                uint32_t r[32];
                tm_ld16(base + lane_base + (it & 1) * 128 + colw + ch * 16, r);
                tm_ld16(base + lane_base + (it & 1) * 128 + colw + ch * 16 + 16, r + 16);
                tm_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j++) acc += __uint_as_float(r[j] & 0x3fffffffu);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                uint2 v = sm[((j + it) & 31) * 512 + tid];
                acc += __uint_as_float(v.x & 0x3fffffffu) + __uint_as_float(v.y & 0x3fffffffu);
            }
        }
    }
    long long t5 = clock64();
    // ---- latency of one dependent tcgen05.ld.x16 + wait (single warp's view)
    long long t6 = clock64();
    uint32_t a = base + lane_base + colw;
    for (int it = 0; it < 64; it++) {
        uint32_t r[16];
        tm_ld16(a, r);
        tm_wait_ld();
        a = base + lane_base + colw + (r[0] & 16);
    }
    long long t7 = clock64();
    if (tid == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; cyc[3] = t7 - t6; }
    if (acc == 123.456f) *sink = acc + (float)a;
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

int main()
{
    unsigned *errs; long long *cyc; float *sink;
    cudaMalloc(&errs, 8); cudaMalloc(&cyc, 64); cudaMalloc(&sink, 4);
    cudaMemset(errs, 0, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 512 * 8);
    const int iters = 2000;
    probe<<<148, 512, 32 * 512 * 8>>>(errs, cyc, iters, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned h[2]; long long c[4];
    cudaMemcpy(h, errs, 8, cudaMemcpyDeviceToHost); cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
    printf("tmem_probe: %s, mismatches=%u, tmem base=0x%x\n", cudaGetErrorString(e), h[0], h[1]);
    printf("per iteration (64 x 32-bit per thread, 16 warps): tcgen05.ld %.1f cycles, ld.shared %.1f cycles, half/half %.1f cycles; dependent ld.x16+wait latency %.1f cycles\n",
           (double)c[0] / iters, (double)c[1] / iters, (double)c[2] / iters, (double)c[3] / 64);
    printf("  -> per SM: TMEM %.1f B/clk, shared %.1f B/clk\n", 512.0 * 256 / ((double)c[0] / iters), 512.0 * 256 / ((double)c[1] / iters));
    return 0;
}
