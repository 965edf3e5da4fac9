// experiment: what does the column pass's access pattern (64-byte row segments at an 8256-byte pitch, 1024 rows per item, read + written in place)
// sustain through the TMA unit as a function of the BYTES IN FLIGHT per SM?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 coltma_probe.cu -o coltma_probe
// One CTA per SM, one thread drives: a ring of NBUF 64 KB buffers; item n is loaded (4 boxes of 256 rows x 64 B) into buffer n % NBUF, and once it
// has landed it is stored back (one bulk-tensor store group); the load of item n + NBUF waits for that store to have read the buffer.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdint>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb), "r"(parity) : "memory");
}
template <int NBUF, int SEGC>      // SEGC complex columns per item (8 = 64-byte segments, 16 = 128-byte segments with 512-row items)
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tl, const __grid_constant__ CUtensorMap ts, int nwork, int iblocks, int rows_item, int store)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t mbar[NBUF];
    if (threadIdx.x != 0) return;
    for (int b = 0; b < NBUF; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
    const int boxes = rows_item / 256, box_bytes = 256 * SEGC * 8;
    auto issue = [&](int n, int buf) {
        const int w = blockIdx.x + n * gridDim.x, q = w / iblocks, ib = w - q * iblocks;
        const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(boxes * box_bytes)) : "memory");
        for (int b = 0; b < boxes; b++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(s0 + buf * 65536 + b * box_bytes), "l"(reinterpret_cast<uint64_t>(&tl)), "r"(ib * SEGC), "r"(q * rows_item + b * 256), "r"(mb) : "memory");
    };
    int nitems = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) nitems++;
    for (int n = 0; n < NBUF && n < nitems; n++) issue(n, n);
    for (int n = 0; n < nitems; n++) {
        const int buf = n % NBUF;
        mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar[buf]), (uint32_t)((n / NBUF) & 1));
        if (store) {
            const int w = blockIdx.x + n * gridDim.x, q = w / iblocks, ib = w - q * iblocks;
            for (int b = 0; b < boxes; b++)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(reinterpret_cast<uint64_t>(&ts)), "r"(ib * SEGC), "r"(q * rows_item + b * 256), "r"(s0 + buf * 65536 + b * box_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // (serialises the store's read of the buffer with the next load into it)
        }
        if (n + NBUF < nitems) issue(n + NBUF, buf);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int NBUF, int SEGC> static void run(PFN enc, void *ws, int64_t rows_total, int inner, int rows_item, int store, const char *tag)
{
    CUtensorMap tl; memset(&tl, 0, sizeof(tl));
    cuuint64_t gd[2] = {(cuuint64_t)inner, (cuuint64_t)rows_total}, gs[1] = {(cuuint64_t)inner * 8};
    cuuint32_t box[2] = {(cuuint32_t)SEGC, 256}, es[2] = {1, 1};
    if (enc(&tl, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, ws, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return; }
    const int iblocks = inner / SEGC, nwork = (int)(rows_total / rows_item) * iblocks;
    cudaFuncSetAttribute(k<NBUF, SEGC>, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * 65536 + 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<NBUF, SEGC><<<148, 128, NBUF * 65536 + 256>>>(tl, tl, nwork, iblocks, rows_item, store);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) {
            const double bytes = (double)rows_total * iblocks * SEGC * 8.0 * (store ? 2.0 : 1.0);
            printf("%-34s NBUF=%d (%3d KB in flight per SM) %s: %.3f ms, %.0f GB/s  %s\n", tag, NBUF, NBUF * 64, store ? "read+write" : "read only ", ms, bytes / ms / 1e6, cudaGetErrorString(e));
        }
    }
}
int main()
{
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    PFN enc = (PFN)f;
    const int inner = 1032;
    const int64_t rows_total = 578ll * 1024;       // c5: 34 x 17 tiles of 1024 rows
    void *ws; cudaMalloc(&ws, (size_t)rows_total * inner * 8);
    cudaMemset(ws, 0, (size_t)rows_total * inner * 8);
    for (int store = 0; store < 2; store++) {
        run<1, 8>(enc, ws, rows_total, inner, 1024, store, "1024 rows x 64 B");
        run<2, 8>(enc, ws, rows_total, inner, 1024, store, "1024 rows x 64 B");
        run<3, 8>(enc, ws, rows_total, inner, 1024, store, "1024 rows x 64 B");
        run<1, 16>(enc, ws, rows_total, 1024, 512, store, "512 rows x 128 B (pitch 8192)");
        run<2, 16>(enc, ws, rows_total, 1024, 512, store, "512 rows x 128 B (pitch 8192)");
        run<3, 16>(enc, ws, rows_total, 1024, 512, store, "512 rows x 128 B (pitch 8192)");
    }
    return 0;
}
