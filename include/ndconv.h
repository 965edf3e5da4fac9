/*
 * ndconv.h -- C ABI of libndconv_cuda.so: the B200 (sm_100a) implementation of the
 * convolution hot path of the Rust crate TYPEmber/ndarray-conv v0.6.1.
 *
 * The reference has no FFI of its own: its boundary is the Rust trait surface re-exported at
 * src/lib.rs:74-78.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference root).  A `cuda` Cargo feature forwards the same traits
 * to these symbols (see INTEGRATION.md for the Rust `extern "C"` block and trait impls, and
 * ndarray-conv_b200/include/ndconv.hpp for the C++ mirror of the trait API used in this image,
 * which has no Rust toolchain).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / CUDA types in any signature (a CUDA stream is
 *    passed as void*).
 *  - every function returns an ndconv_status (0 = ok) unless stated otherwise; a message for
 *    the last failure on the calling thread is available from ndconv_last_error_string().
 *  - 1..3 mirror the reference's `Error<N>` variants (src/lib.rs:148-159); everything the
 *    reference answers with a panic is reported as NDCONV_ERR_PANIC instead of unwinding.
 *  - there is NO CPU fallback: compute entry points fail with NDCONV_ERR_CUDA when no sm_100
 *    device / driver is present.
 *  - a processor handle is not thread-safe (one in-flight call per handle, like `&mut self` in
 *    src/conv_fft/processor/mod.rs:91-118); distinct handles may be used from distinct threads.
 */
#ifndef NDCONV_H
#define NDCONV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDCONV_MAX_DIM 6 /* ndarray's fixed-rank Dim<[Ix; N]>, N = 1..6 */

typedef enum ndconv_status {
    NDCONV_OK = 0,
    NDCONV_ERR_DATA_SHAPE = 1,     /* Error::DataShape      src/lib.rs:150-152 */
    NDCONV_ERR_KERNEL_SHAPE = 2,   /* Error::KernelShape    src/lib.rs:153-155 */
    NDCONV_ERR_MISMATCH_SHAPE = 3, /* Error::MismatchShape  src/lib.rs:156-158 */
    NDCONV_ERR_PANIC = 4,          /* the reference would panic (index out of bounds in half_dim.rs, stride 0, ...) */
    NDCONV_ERR_BAD_ARG = 5,        /* malformed descriptor (ndim, dtype, null pointer, ...) */
    NDCONV_ERR_UNSUPPORTED = 6,    /* valid for the reference, outside this build's envelope */
    NDCONV_ERR_CUDA = 100,         /* CUDA runtime / driver failure, or no device */
    NDCONV_ERR_INTERNAL = 101
} ndconv_status;

/* element types: `T: NumAssign + Copy` of ConvExt (src/conv/mod.rs:118-121), `InElem` of ConvFFTExt */
typedef enum ndconv_dtype {
    NDCONV_I32 = 0, NDCONV_I64 = 1, NDCONV_F32 = 2, NDCONV_F64 = 3,
    NDCONV_C32 = 4, /* num::Complex<f32>, interleaved re,im */
    NDCONV_C64 = 5, /* num::Complex<f64> */
    NDCONV_I8 = 6, NDCONV_I16 = 7, NDCONV_U8 = 8, NDCONV_U16 = 9, NDCONV_U32 = 10, NDCONV_U64 = 11,
    NDCONV_I128 = 12, NDCONV_U128 = 13 /* Rust i128 / u128: 16 bytes, little-endian, direct conv only (wrapping, as in release builds) */
} ndconv_dtype;

/* BorderType<T>, src/lib.rs:131-143 */
typedef enum ndconv_border_type {
    NDCONV_BORDER_ZEROS = 0, NDCONV_BORDER_CONST = 1, NDCONV_BORDER_REFLECT = 2,
    NDCONV_BORDER_REPLICATE = 3, NDCONV_BORDER_CIRCULAR = 4
} ndconv_border_type;

/* ConvMode<N>, src/lib.rs:80-105 */
typedef enum ndconv_conv_mode {
    NDCONV_MODE_FULL = 0, NDCONV_MODE_SAME = 1, NDCONV_MODE_VALID = 2,
    NDCONV_MODE_CUSTOM = 3, NDCONV_MODE_EXPLICIT = 4
} ndconv_conv_mode;

typedef enum ndconv_memory {
    NDCONV_MEM_HOST = 0,  /* data / out are host pointers; the call copies in, computes, copies out, and has completed on return */
    NDCONV_MEM_DEVICE = 1 /* data / out are device pointers on the processor's device; the call is enqueued on the processor's stream */
} ndconv_memory;

typedef enum ndconv_path { NDCONV_PATH_DIRECT = 0, NDCONV_PATH_FFT = 1 } ndconv_path;

/* one side of one axis; `value` holds the Const payload in the problem's own dtype (first sizeof(T) bytes) */
typedef struct ndconv_border {
    int32_t type; /* ndconv_border_type */
    int32_t reserved;
    unsigned char value[16];
} ndconv_border;

/*
 * One convolution, fully lowered: what ConvExt::conv / ConvFFTExt::conv_fft see after
 * `kernel.into_kernel_with_dilation()` (src/dilation/mod.rs:8-12,192-212),
 * `conv_mode.unfold(&kwd)` (src/conv/mod.rs:28-66) and the PaddingMode -> per-side BorderType
 * lowering of the Custom / Explicit drivers (src/padding/mod.rs:346-452).
 * Strides are in ELEMENTS and may be zero or negative (ndarray views).  The kernel pointer is
 * always a HOST pointer (it is small; taps and its spectrum are derived from it on the host side).
 */
typedef struct ndconv_problem {
    int32_t dtype;  /* ndconv_dtype */
    int32_t ndim;   /* 1..NDCONV_MAX_DIM */
    int32_t memory; /* ndconv_memory: where `data` and the output live */
    int32_t reverse;/* 1 = flip the kernel (true convolution, default of with_dilation, src/dilation/mod.rs:118-125); 0 = no_reverse() */
    const void *data;
    int64_t data_shape[NDCONV_MAX_DIM];
    int64_t data_strides[NDCONV_MAX_DIM];
    const void *kernel;
    int64_t kernel_shape[NDCONV_MAX_DIM];
    int64_t kernel_strides[NDCONV_MAX_DIM];
    int64_t dilation[NDCONV_MAX_DIM];
    int64_t pad[NDCONV_MAX_DIM][2];   /* ExplicitConv::padding, src/conv/mod.rs:23-26 */
    int64_t stride[NDCONV_MAX_DIM];   /* ExplicitConv::strides */
    ndconv_border border[NDCONV_MAX_DIM][2];
} ndconv_problem;

typedef struct ndconv_processor ndconv_processor; /* opaque: device, stream, plans, twiddles, cached kernel spectrum, workspaces */

/* ---- library -------------------------------------------------------------------------- */
const char *ndconv_version(void);
/* 0 for the CUDA product.  (tests build a host emulation of the kernel bodies that answers 1; the product loader rejects it.) */
int ndconv_is_emulation(void);
const char *ndconv_last_error_string(void);
const char *ndconv_status_string(int status);
size_t ndconv_dtype_size(int dtype);
/* number of sm_100 devices visible, or a negative ndconv_status */
int ndconv_device_count(void);

/* ---- host-side lowering helpers (no GPU needed) ----------------------------------------- */
/* ConvMode::unfold, src/conv/mod.rs:28-66.  `padding`: [ndim] for CUSTOM, [ndim][2] for EXPLICIT, ignored otherwise. */
int ndconv_unfold_conv_mode(int mode, int ndim, const int64_t *kernel_shape, const int64_t *dilation,
                            const int64_t *padding, const int64_t *strides,
                            int64_t out_pad[][2], int64_t *out_stride);
/* good_size_cc, src/conv_fft/good_size.rs:6-31 (kept for API parity; the GPU path picks its own lengths) */
int64_t ndconv_good_fft_size(int64_t n);
/* the {2,3,5,7}-smooth transform length >= n this build would use for one un-tiled axis (even when real_axis != 0) */
int64_t ndconv_plan_fft_size(int64_t n, int real_axis);
/* shape checks in the reference's order + output shape.  conv: src/conv/mod.rs:136-168; conv_fft: src/conv_fft/mod.rs:205-227,282-289 */
int ndconv_out_shape(const ndconv_problem *problem, int path, int64_t *out_shape);
/* the 1-D border index map of one axis: out_map[i], i in [0, n+pf+pb): source index >= 0, -1 front constant,
 * -2 back constant, -3 never written (reads the zero-initialised buffer).  Literal symbolic replay of
 * src/padding/half_dim.rs:30-343.  Returns NDCONV_ERR_PANIC where the reference panics. */
int ndconv_border_index_map(int64_t n, int64_t pad_front, int64_t pad_back, int border_front, int border_back, int32_t *out_map);

/* ---- processors: get_fft_processor / GetProcessor::get_processor, src/conv_fft/processor/mod.rs:71-73,125-143 -- */
int ndconv_processor_create(int device, ndconv_processor **out);
int ndconv_processor_destroy(ndconv_processor *p);
/* use an existing CUDA stream (cudaStream_t as void*); NULL restores the processor's own stream */
int ndconv_processor_set_stream(ndconv_processor *p, void *cuda_stream);
int ndconv_processor_synchronize(ndconv_processor *p);
/* kernels launched by this processor so far (bench.py's gpu_launches) */
int64_t ndconv_processor_launch_count(const ndconv_processor *p);
/* bytes of device workspace currently held */
int64_t ndconv_processor_workspace_bytes(const ndconv_processor *p);

/* per-kernel device timing (CUDA events on the processor's stream around every launch) for bench.py's roofline:
 * enable, run, then read the totals per kernel name since the last read.  alg_bytes = sum of the algorithmic bytes of the
 * recorded launches.  Returns the number of entries written (<= max_entries). */
int ndconv_processor_set_profiling(ndconv_processor *p, int enable);
int ndconv_processor_get_profile(ndconv_processor *p, int max_entries, char (*names)[64], double *total_ms, int64_t *launches, double *alg_bytes);

/* ---- the hot path ---------------------------------------------------------------------- */
/* ConvExt::conv, src/conv/mod.rs:110-115,128-200.  `out`: contiguous standard-layout buffer of ndconv_out_shape()
 * elements, in the memory space named by problem->memory.  `p` may be NULL for host problems (a transient processor on device 0). */
int ndconv_conv_direct(ndconv_processor *p, const ndconv_problem *problem, void *out);
/* ConvFFTExt::conv_fft_with_processor, src/conv_fft/mod.rs:404-412 (p == NULL: conv_fft, :394-402, fresh processor per call).
 * dtype must be F32 / F64 / C32 / C64. */
int ndconv_conv_fft(ndconv_processor *p, const ndconv_problem *problem, void *out);
/* ConvFFTExt::conv_fft_par, src/conv_fft/mod.rs:414-423: same GPU call (the parallelism is the device's). */
int ndconv_conv_fft_par(ndconv_processor *p, const ndconv_problem *problem, void *out);
/* conv_fft_par with more than one GPU configured: the reference's rayon path spreads one convolution over the
 * host's cores (src/conv_fft/mod.rs:242-261); here one HOST-resident convolution is spread over `n_processors` handles (one
 * per GPU, any mix of devices): output rows of axis 0 are cut into one contiguous overlap-save slab per handle
 * (ndconv_slab_plan) and every handle runs H2D | kernels | D2H on its rows from its own host thread, over its own PCIe link,
 * with no data-path collective; all handles write into the one `out` array.  Problems too small to pipeline run on
 * processors[0] alone.  Results are those of ndconv_conv_fft up to rounding (slabs may pick other tile lengths). */
int ndconv_conv_fft_sharded(ndconv_processor *const *processors, int n_processors, const ndconv_problem *problem, void *out);
/* The same convolution when the array is ALREADY DEVICE-RESIDENT and partitioned along axis 0 (SURVEY 2.2 K10, 8e "Collective"; no
 * reference analogue): shard g -- processors[g], on its own device -- owns the contiguous rows [r_g, r_g + shards[g].rows) of the
 * data in standard layout (inner extents as in problem->data_shape; problem->data / data_strides / memory are ignored), with
 * ghost rows: shards[g].data points at the first OWNED row and the allocation extends halo_front rows before and halo_back rows
 * after the owned rows.  The call fills the ghost rows each shard needs -- the (Kd0 - 1)-row halos of its neighbours, or the far
 * end of the array for a Circular border on axis 0 -- with peer-to-peer copies over NVLink (cudaMemcpyPeerAsync on the consumer's
 * stream; nothing is re-materialised, the owned rows are never copied) and enqueues the single-GPU pipeline on every shard in
 * place; the true array edges keep their border rule.  Shard g produces output rows [out_begin, out_end) of ndconv_shard_plan
 * (the balanced split of ndconv_slab_plan) into shards[g].out, contiguous.  Like every device-resident call it returns once the
 * work is enqueued: the shard data must be complete on entry; synchronise each processor afterwards.  One host thread per shard. */
typedef struct ndconv_shard {
    void *data;          /* first owned row, device pointer on processors[g]'s device */
    int64_t rows;        /* owned rows of axis 0 (the shards tile [0, data_shape[0]) in order) */
    int64_t halo_front;  /* writable ghost rows available before / after the owned rows */
    int64_t halo_back;
    void *out;           /* output rows of this shard */
} ndconv_shard;
typedef struct ndconv_shard_info {
    int64_t out_begin, out_end;        /* output rows of axis 0 this shard produces */
    int64_t halo_front, halo_back;     /* ghost rows the call will fill (0 at a true array edge unless the border is Circular) */
    int64_t first_row;                 /* global index of the shard's first owned row */
} ndconv_shard_info;
/* host logic only (no device): what shard `shard` of `n_shards` (owned row counts in shard_rows[]) produces and needs */
int ndconv_shard_plan(const ndconv_problem *problem, int n_shards, const int64_t *shard_rows, int shard, ndconv_shard_info *out);
int ndconv_conv_fft_sharded_device(ndconv_processor *const *processors, int n_processors, const ndconv_problem *problem, const ndconv_shard *shards);
/* A batch of INDEPENDENT conv_fft problems distributed whole over processor handles (the reference has no batch call: a user loops
 * over conv_fft_with_processor, src/conv_fft/mod.rs:404-412): problem i runs on processors[i % n_processors], every
 * handle from its own host thread on its own stream.  Several handles on one device overlap launches that are each smaller than a
 * wave; handles on several devices spread the batch over the GPUs.  outs[i] is problem i's output buffer.  Device-resident
 * problems are only enqueued (synchronise each processor afterwards); host problems are complete on return.  Returns the first
 * failing status, the remaining problems still run. */
int ndconv_conv_fft_batch(ndconv_processor *const *processors, int n_processors, const ndconv_problem *problems, void *const *outs, int n_problems);

/* ---- Processor::{forward, backward}, src/conv_fft/processor/mod.rs:91-118 (real.rs:24-281, complex.rs:33-145) --------------
 * N-d FFT with the reference's spectrum layout: `shape` is always the shape of the REAL-SPACE array
 * [n0..n_{N-1}]; the spectrum has axis 0 moved to the end -- real input: [n1, .., n_{N-2}, n_{N-1}/2+1, n0], complex input:
 * [n1, .., n_{N-1}, n0] (N = 1: [n/2+1] / [n]).  forward is unnormalised (e^{-2 pi i ..}); backward divides by prod(shape)
 * (real.rs:278-279, complex.rs:141-142) and takes the original last-axis length from `shape` (the reference remembers it in
 * `rp_origin_len`, real.rs:41,108).  dtype F32/F64: real input, C32/C64: complex input.  Any length is taken, as rustfft
 * does: an axis that is {2,3,5,7}-smooth and fits one shared-memory transform (last axis <= 8192 real / 4096 complex, even
 * when real; other axes <= 1024, f64: 512) runs in the convolution pipeline's row / column kernels, every other axis (longer,
 * odd real, prime factors above 7) as global-memory Stockham passes -- a prime factor r costs O(n r) there. */
int ndconv_fft_forward(ndconv_processor *p, int dtype, int ndim, const int64_t *shape, const void *in, void *out, int memory);
int ndconv_fft_backward(ndconv_processor *p, int dtype, int ndim, const int64_t *shape, const void *spectrum, void *out, int memory);

/* ---- plan introspection (host logic only: works without a device) ---------------------------------------------------------------
 * What ndconv_conv_fft would do with `problem`: the overlap-save tiling per axis (the analogue of the reference's FFT sizes,
 * src/conv_fft/good_size.rs:6-42 -- any F >= P is valid and unobservable), the kernel family, the spectra workspace it
 * needs on the device and whether the call is cut along axis 0. */
typedef struct ndconv_plan_info {
    int path;                 /* 0 generic FFT kernels, 1 sm_100a fast path, 2 direct kernel (kernel longer than one FFT tile and not splittable),
                                 3 kernel longer than one FFT tile on an axis: cut into n_tiles[axis] segments of tile_valid[axis] taps, summed */
    int ndim;
    int tile_len[6];          /* F_a: transform length of one overlap-save tile */
    int tile_valid[6];        /* V_a = F_a - Kd_a + 1 alias-free positions per tile */
    int n_tiles[6];
    int64_t workspace_bytes;  /* spectra workspace of the whole (unsplit) plan */
    int64_t split_out_rows;   /* > 0: axis 0 is processed as two sub-convolutions, the first producing this many output rows */
    int pipelined;            /* host-resident and large: H2D | kernels | D2H overlapped over axis-0 slabs */
} ndconv_plan_info;
int ndconv_plan_query(const ndconv_problem *problem, ndconv_plan_info *out);

/* ---- multi-GPU slab planning (overlap-save along axis 0) ---------------------- */
typedef struct ndconv_slab {
    int64_t out_begin, out_end;   /* output rows [begin,end) of axis 0 owned by this slab */
    int64_t pad_begin, pad_end;   /* rows [begin,end) of the PADDED axis 0 the slab reads */
} ndconv_slab;
/* split the output rows of axis 0 into n_slabs contiguous ranges (balanced); slab index in [0,n_slabs) */
int ndconv_slab_plan(const ndconv_problem *problem, int path, int n_slabs, int slab, ndconv_slab *out);

/* ---- pinned host memory helpers (so callers can stage inputs for full-rate H2D/D2H) ------ */
void *ndconv_host_alloc(size_t bytes);
void ndconv_host_free(void *ptr);
/* Page-lock / release an allocation the caller already owns (the Vec behind an ndarray).  Large host-resident calls on ordinary
 * pageable memory are staged through pinned bounce buffers by host memcpy threads (30-40 GB/s of host<->device traffic; the
 * driver's own staging gives ~13) and run at 70-80 GB/s from page-locked memory (16384^2, k = 63^2: 57-75 vs 27 ms per call);
 * registering costs ~0.1 ms per MB, so it pays for buffers that serve more than about two calls. */
int ndconv_host_register(void *ptr, size_t bytes);
int ndconv_host_unregister(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* NDCONV_H */
