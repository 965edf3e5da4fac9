// experiment: issue rate of the packed FP32 instructions (FADD2 / FFMA2) against scalar FADD / FFMA on B200, and a register
// radix-32 butterfly written both ways.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DNDCONV_CUDA -I../../ndarray-conv_b200/csrc
//                                              -I../../include packed_probe.cu -o packed_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cmath>
#include "kernels_fft.h"
#include "packed_cf.cuh"
using namespace ndc;

template <int MODE> __global__ void __launch_bounds__(256) rate(float *out, int iters, float s)
{
    // 8 independent chains per thread
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    pk::pcf p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pk::mk(a[2 * i], a[2 * i + 1]);
    const pk::pcf ps = pk::mk(s, s), pt = pk::mk(0.5f, 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], s, 0.5f);            // 16 FFMA
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = pk::fma(p[i], ps, pt);           // 8 FFMA2 = 16 lanes
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] = a[i] + s;                       // 16 FADD
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = pk::add(p[i], ps);               // 8 FADD2
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) r += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++) r += pk::re(p[i]) + pk::im(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> __global__ void __launch_bounds__(128, 4) bfly(const float2 *in, float2 *out, int iters)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0) {
        cx<float> v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) { float2 q = in[(size_t)j * gridDim.x * blockDim.x + tid]; v[j] = cx<float>{q.x, q.y}; }
        for (int it = 0; it < iters; it++) {
            dft32<float>(v, (it & 1) != 0);
            if (it & 1) {
#pragma unroll
                for (int j = 0; j < 32; j++) { v[j].re *= 0.03125f; v[j].im *= 0.03125f; }
            }
        }
#pragma unroll
        for (int j = 0; j < 32; j++) out[(size_t)j * gridDim.x * blockDim.x + tid] = make_float2(v[j].re, v[j].im);
    } else {
        pk::pcf v[32];
        const uint64_t *in8 = reinterpret_cast<const uint64_t *>(in);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j].v = in8[(size_t)j * gridDim.x * blockDim.x + tid];
        for (int it = 0; it < iters; it += 2) {
            pk::dft<false, 32>(v);
            pk::dft<true, 32>(v);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = pk::scale(v[j], 0.03125f);
        }
        uint64_t *o8 = reinterpret_cast<uint64_t *>(out);
#pragma unroll
        for (int j = 0; j < 32; j++) o8[(size_t)j * gridDim.x * blockDim.x + tid] = v[j].v;
    }
}

template <class F> static float time_ms(F f, int reps = 5)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    return best;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s  SMs %d  clock %d MHz\n", pr.name, pr.multiProcessorCount, clk_khz / 1000);
    const int nsm = pr.multiProcessorCount;
    float *out; cudaMalloc(&out, (size_t)nsm * 8 * 256 * 4);
    const int iters = 4096;
    const char *names[4] = {"FFMA  (scalar)", "FFMA2 (packed)", "FADD  (scalar)", "FADD2 (packed)"};
    for (int m = 0; m < 4; m++) {
        float ms = 0;
        for (int bps : {2, 4, 8}) {
            auto f = [&] {
                if (m == 0) rate<0><<<nsm * bps, 256>>>(out, iters, 1.0001f);
                if (m == 1) rate<1><<<nsm * bps, 256>>>(out, iters, 1.0001f);
                if (m == 2) rate<2><<<nsm * bps, 256>>>(out, iters, 1.0001f);
                if (m == 3) rate<3><<<nsm * bps, 256>>>(out, iters, 1.0001f);
            };
            ms = time_ms(f);
            const double lanes = (double)nsm * bps * 256 * iters * 4 * 16;     // f32 lane-operations
            printf("%s  %d CTAs/SM x 256 thr: %.3f ms  -> %.1f lane-ops/clk/SM (at %d MHz)  %.2f T lane-ops/s\n", names[m], bps, ms,
                   lanes / (ms * 1e-3) / nsm / (clk_khz * 1e3), clk_khz / 1000, lanes / (ms * 1e-3) / 1e12);
        }
    }
    // butterflies
    const int nthreads = nsm * 4 * 128 * 4;
    float2 *in, *o0, *o1;
    cudaMalloc(&in, (size_t)nthreads * 32 * 8); cudaMalloc(&o0, (size_t)nthreads * 32 * 8); cudaMalloc(&o1, (size_t)nthreads * 32 * 8);
    float2 *h = (float2 *)malloc((size_t)nthreads * 32 * 8);
    for (size_t i = 0; i < (size_t)nthreads * 32; i++) h[i] = make_float2((float)((i * 2654435761u) % 1000) / 1000.f, (float)((i * 40503u) % 977) / 977.f);
    cudaMemcpy(in, h, (size_t)nthreads * 32 * 8, cudaMemcpyHostToDevice);
    const int bit = 64;
    float t0 = time_ms([&] { bfly<0><<<nthreads / 128, 128>>>(in, o0, bit); });
    float t1 = time_ms([&] { bfly<1><<<nthreads / 128, 128>>>(in, o1, bit); });
    printf("radix-32 butterfly x%d per thread, %d threads: scalar %.3f ms, packed %.3f ms (%.2fx)\n", bit, nthreads, t0, t1, t0 / t1);
    float2 *h0 = (float2 *)malloc((size_t)nthreads * 32 * 8), *h1 = (float2 *)malloc((size_t)nthreads * 32 * 8);
    cudaMemcpy(h0, o0, (size_t)nthreads * 32 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h1, o1, (size_t)nthreads * 32 * 8, cudaMemcpyDeviceToHost);
    double md = 0, mx = 0;
    for (size_t i = 0; i < (size_t)nthreads * 32; i++) {
        md = fmax(md, fmax(fabs(h0[i].x - h1[i].x), fabs(h0[i].y - h1[i].y)));
        mx = fmax(mx, fmax(fabs(h0[i].x), fabs(h0[i].y)));
    }
    printf("scalar vs packed butterflies: max|diff| %.3e  max|val| %.3e   (round trip vs input: ", md, mx);
    double mr = 0;
    for (size_t i = 0; i < (size_t)nthreads * 32; i++) mr = fmax(mr, fmax(fabs(h1[i].x - h[i].x), fabs(h1[i].y - h[i].y)));
    printf("%.3e)\n", mr);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
