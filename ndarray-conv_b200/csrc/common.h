// common.h -- shared definitions for the host logic and the kernel bodies.
//
// Kernel bodies in kernels_*.h are written against `BlockCtx` in block-stride style: every
// phase is `for (i = c.tid; i < n; i += c.nt)` and phases are separated by c.sync().  nvcc
// compiles them into __global__ entry points for sm_100a.  The same bodies compile as plain C++
// when NDCONV_HOST_EMUL is defined (one "thread" per block, blocks run in a loop): that build
// exists ONLY under tests/emul/ so index logic can be checked against the oracle on a box with
// no GPU.  The product library is never built with NDCONV_HOST_EMUL and has no CPU path.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__) && !defined(NDCONV_HOST_EMUL)
#define NDCONV_CUDA 1
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

#define NDC_MAX_DIM 6
#define NDC_MAX_PASS 16

// border index map codes (ndconv_border_index_map)
#define NDC_MAP_CONST_FRONT (-1)
#define NDC_MAP_CONST_BACK (-2)
#define NDC_MAP_INIT (-3)

namespace ndc {

template <class R> struct cx { R re, im; };

template <class R> HD cx<R> cmul(cx<R> a, cx<R> b) { return cx<R>{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
template <class R> HD cx<R> cmulc(cx<R> a, cx<R> b) { return cx<R>{a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im}; }  // a * conj(b)
template <class R> HD cx<R> cadd(cx<R> a, cx<R> b) { return cx<R>{a.re + b.re, a.im + b.im}; }
template <class R> HD cx<R> csub(cx<R> a, cx<R> b) { return cx<R>{a.re - b.re, a.im - b.im}; }
template <class R> HD cx<R> cconj(cx<R> a) { return cx<R>{a.re, -a.im}; }

struct BlockCtx {
    int tid, nt;          // thread index / threads per block
    int64_t bid, nb;      // block index / number of blocks
    char *smem;           // dynamic shared memory
    HD void sync() const {
#ifdef NDCONV_CUDA
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
#endif
    }
};

}  // namespace ndc
