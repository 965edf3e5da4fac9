//! `extern "C"` declarations of libndconv_cuda.so -- one item per declaration of include/ndconv.h, in the header's order.
//! tests/test_rust_binding.py parses this file and checks field order / sizes / symbol names against the header and
//! against the ctypes mirror (there is no Rust toolchain in the repository's build image).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const MAX_DIM: usize = 6;

// ndconv_status
pub const OK: c_int = 0;
pub const ERR_DATA_SHAPE: c_int = 1;
pub const ERR_KERNEL_SHAPE: c_int = 2;
pub const ERR_MISMATCH_SHAPE: c_int = 3;
pub const ERR_PANIC: c_int = 4;
pub const ERR_BAD_ARG: c_int = 5;
pub const ERR_UNSUPPORTED: c_int = 6;
pub const ERR_CUDA: c_int = 100;
pub const ERR_INTERNAL: c_int = 101;

// ndconv_dtype
pub const I32: i32 = 0;
pub const I64: i32 = 1;
pub const F32: i32 = 2;
pub const F64: i32 = 3;
pub const C32: i32 = 4;
pub const C64: i32 = 5;
pub const I8: i32 = 6;
pub const I16: i32 = 7;
pub const U8: i32 = 8;
pub const U16: i32 = 9;
pub const U32: i32 = 10;
pub const U64: i32 = 11;
pub const I128: i32 = 12;
pub const U128: i32 = 13;

// ndconv_border_type
pub const BORDER_ZEROS: i32 = 0;
pub const BORDER_CONST: i32 = 1;
pub const BORDER_REFLECT: i32 = 2;
pub const BORDER_REPLICATE: i32 = 3;
pub const BORDER_CIRCULAR: i32 = 4;

// ndconv_memory / ndconv_path
pub const MEM_HOST: i32 = 0;
pub const MEM_DEVICE: i32 = 1;
pub const PATH_DIRECT: c_int = 0;
pub const PATH_FFT: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ndconv_border {
    pub r#type: i32,
    pub reserved: i32,
    pub value: [u8; 16],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ndconv_problem {
    pub dtype: i32,
    pub ndim: i32,
    pub memory: i32,
    pub reverse: i32,
    pub data: *const c_void,
    pub data_shape: [i64; MAX_DIM],
    pub data_strides: [i64; MAX_DIM],
    pub kernel: *const c_void,
    pub kernel_shape: [i64; MAX_DIM],
    pub kernel_strides: [i64; MAX_DIM],
    pub dilation: [i64; MAX_DIM],
    pub pad: [[i64; 2]; MAX_DIM],
    pub stride: [i64; MAX_DIM],
    pub border: [[ndconv_border; 2]; MAX_DIM],
}

#[repr(C)]
pub struct ndconv_processor {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct ndconv_plan_info {
    pub path: c_int,
    pub ndim: c_int,
    pub tile_len: [c_int; 6],
    pub tile_valid: [c_int; 6],
    pub n_tiles: [c_int; 6],
    pub workspace_bytes: i64,
    pub split_out_rows: i64,
    pub pipelined: c_int,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ndconv_shard {
    pub data: *mut c_void,
    pub rows: i64,
    pub halo_front: i64,
    pub halo_back: i64,
    pub out: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct ndconv_shard_info {
    pub out_begin: i64,
    pub out_end: i64,
    pub halo_front: i64,
    pub halo_back: i64,
    pub first_row: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct ndconv_slab {
    pub out_begin: i64,
    pub out_end: i64,
    pub pad_begin: i64,
    pub pad_end: i64,
}

extern "C" {
    // ---- library ----
    pub fn ndconv_version() -> *const c_char;
    pub fn ndconv_is_emulation() -> c_int;
    pub fn ndconv_last_error_string() -> *const c_char;
    pub fn ndconv_status_string(status: c_int) -> *const c_char;
    pub fn ndconv_dtype_size(dtype: c_int) -> usize;
    pub fn ndconv_device_count() -> c_int;
    // ---- host-side lowering helpers ----
    pub fn ndconv_unfold_conv_mode(mode: c_int, ndim: c_int, kernel_shape: *const i64, dilation: *const i64, padding: *const i64,
                                   strides: *const i64, out_pad: *mut [i64; 2], out_stride: *mut i64) -> c_int;
    pub fn ndconv_good_fft_size(n: i64) -> i64;
    pub fn ndconv_plan_fft_size(n: i64, real_axis: c_int) -> i64;
    pub fn ndconv_out_shape(problem: *const ndconv_problem, path: c_int, out_shape: *mut i64) -> c_int;
    pub fn ndconv_border_index_map(n: i64, pad_front: i64, pad_back: i64, border_front: c_int, border_back: c_int, out_map: *mut i32) -> c_int;
    // ---- processors ----
    pub fn ndconv_processor_create(device: c_int, out: *mut *mut ndconv_processor) -> c_int;
    pub fn ndconv_processor_destroy(p: *mut ndconv_processor) -> c_int;
    pub fn ndconv_processor_set_stream(p: *mut ndconv_processor, cuda_stream: *mut c_void) -> c_int;
    pub fn ndconv_processor_synchronize(p: *mut ndconv_processor) -> c_int;
    pub fn ndconv_processor_launch_count(p: *const ndconv_processor) -> i64;
    pub fn ndconv_processor_workspace_bytes(p: *const ndconv_processor) -> i64;
    pub fn ndconv_processor_set_profiling(p: *mut ndconv_processor, enable: c_int) -> c_int;
    pub fn ndconv_processor_get_profile(p: *mut ndconv_processor, max_entries: c_int, names: *mut [c_char; 64], total_ms: *mut f64,
                                        launches: *mut i64, alg_bytes: *mut f64) -> c_int;
    // ---- the hot path ----
    pub fn ndconv_conv_direct(p: *mut ndconv_processor, problem: *const ndconv_problem, out: *mut c_void) -> c_int;
    pub fn ndconv_conv_fft(p: *mut ndconv_processor, problem: *const ndconv_problem, out: *mut c_void) -> c_int;
    pub fn ndconv_conv_fft_par(p: *mut ndconv_processor, problem: *const ndconv_problem, out: *mut c_void) -> c_int;
    pub fn ndconv_conv_fft_sharded(processors: *const *mut ndconv_processor, n_processors: c_int, problem: *const ndconv_problem, out: *mut c_void) -> c_int;
    pub fn ndconv_shard_plan(problem: *const ndconv_problem, n_shards: c_int, shard_rows: *const i64, shard: c_int, out: *mut ndconv_shard_info) -> c_int;
    pub fn ndconv_conv_fft_sharded_device(processors: *const *mut ndconv_processor, n_processors: c_int, problem: *const ndconv_problem,
                                          shards: *const ndconv_shard) -> c_int;
    pub fn ndconv_conv_fft_batch(processors: *const *mut ndconv_processor, n_processors: c_int, problems: *const ndconv_problem,
                                 outs: *const *mut c_void, n_problems: c_int) -> c_int;
    // ---- Processor::{forward, backward} ----
    pub fn ndconv_fft_forward(p: *mut ndconv_processor, dtype: c_int, ndim: c_int, shape: *const i64, input: *const c_void, out: *mut c_void, memory: c_int) -> c_int;
    pub fn ndconv_fft_backward(p: *mut ndconv_processor, dtype: c_int, ndim: c_int, shape: *const i64, spectrum: *const c_void, out: *mut c_void, memory: c_int) -> c_int;
    // ---- plan introspection / slab planning ----
    pub fn ndconv_plan_query(problem: *const ndconv_problem, out: *mut ndconv_plan_info) -> c_int;
    pub fn ndconv_slab_plan(problem: *const ndconv_problem, path: c_int, n_slabs: c_int, slab: c_int, out: *mut ndconv_slab) -> c_int;
    // ---- pinned host memory ----
    pub fn ndconv_host_alloc(bytes: usize) -> *mut c_void;
    pub fn ndconv_host_free(ptr: *mut c_void);
    pub fn ndconv_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn ndconv_host_unregister(ptr: *mut c_void) -> c_int;
}
