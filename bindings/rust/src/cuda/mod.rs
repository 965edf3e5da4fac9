//! Feature `cuda`: ConvExt / ConvFFTExt / Processor forwarded to libndconv_cuda.so (B200, sm_100a) through the C ABI of
//! include/ndconv.h.  Replaces, under the feature, the bodies of src/conv/mod.rs:118-201 and src/conv_fft/mod.rs:185-423 and the
//! processors of src/conv_fft/processor/{real,complex}.rs; the public traits, enums and error type are unchanged.
//!
//! How it plugs into the crate (see bindings/rust/README.md):
//!   * src/lib.rs gains `#[cfg(feature = "cuda")] pub mod cuda;`
//!   * the CPU `impl ConvExt for ArrayBase` / `impl ConvFFTExt for ArrayBase` and the `GetProcessor` impls are gated
//!     `#[cfg(not(feature = "cuda"))]`; this module provides the `#[cfg(feature = "cuda")]` ones
//!   * `trait Processor` gains one hidden method under the feature, `fn cuda_raw(&mut self) -> *mut ffi::ndconv_processor`,
//!     which is how `conv_fft_with_processor(&mut impl Processor)` reaches the device handle
//!
//! There is no CPU fallback: without an sm_100 device every call panics with the CUDA error text (Error<N> is not
//! #[non_exhaustive], so a new error variant would be a semver break).
pub mod ffi;

use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};

use ndarray::{Array, ArrayBase, Data, DataMut, Dim, Dimension, IntoDimension, Ix, RemoveAxis};
use num::traits::NumAssign;
use num::Complex;
use rustfft::FftNum;

use crate::conv_fft::{GetProcessor, Processor};
use crate::dilation::{IntoKernelWithDilation, KernelWithDilation};
use crate::{BorderType, ConvExt, ConvFFTExt, ConvMode, Error, PaddingMode};

/// Host arrays at least this large take the library's pipelined path (H2D | kernels | D2H over axis-0 slabs).  An `Array` owns
/// pageable `Vec` memory: the library then stages the slabs through its own pinned bounce buffers (30-40 GB/s of host<->device
/// traffic; c5: 311 ms per call).  Page-locking the caller's allocation in place lifts that to 70-80 GB/s (c5: 101 ms) but costs
/// ~0.1 ms per MB (c5: ~0.43 s per array), so it only pays for an allocation that serves several calls: the binding registers
/// an allocation the SECOND time it sees it (same pointer and length) and keeps it registered until `unregister_all()` or thread
/// exit; one-shot arrays never pay for registration.
pub const AUTO_REGISTER_BYTES: usize = 96 << 20;

/// Element types the device path takes; `DTYPE` is the `ndconv_dtype` code.
///
/// # Safety
/// `Self` must have exactly the size and layout of the C element type named by `DTYPE`.
pub unsafe trait CudaElem: Copy + 'static {
    const DTYPE: i32;
}
macro_rules! cuda_elem {
    ($($t:ty => $c:expr),* $(,)?) => { $(unsafe impl CudaElem for $t { const DTYPE: i32 = $c; })* };
}
cuda_elem!(i32 => ffi::I32, i64 => ffi::I64, f32 => ffi::F32, f64 => ffi::F64, Complex<f32> => ffi::C32, Complex<f64> => ffi::C64,
           i8 => ffi::I8, i16 => ffi::I16, u8 => ffi::U8, u16 => ffi::U16, u32 => ffi::U32, u64 => ffi::U64,
           i128 => ffi::I128, u128 => ffi::U128);
#[cfg(target_pointer_width = "64")]
cuda_elem!(isize => ffi::I64, usize => ffi::U64);

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::ndconv_last_error_string()) }.to_string_lossy().into_owned()
}

fn border<T: CudaElem>(b: BorderType<T>) -> ffi::ndconv_border {
    let mut out = ffi::ndconv_border { r#type: ffi::BORDER_ZEROS, reserved: 0, value: [0u8; 16] };
    match b {
        BorderType::Zeros => {}
        BorderType::Const(c) => {
            out.r#type = ffi::BORDER_CONST;
            let n = std::mem::size_of::<T>();
            assert!(n <= 16);
            // the Const payload in the problem's own dtype, first size_of::<T>() bytes
            unsafe { std::ptr::copy_nonoverlapping(&c as *const T as *const u8, out.value.as_mut_ptr(), n) };
        }
        BorderType::Reflect => out.r#type = ffi::BORDER_REFLECT,
        BorderType::Replicate => out.r#type = ffi::BORDER_REPLICATE,
        BorderType::Circular => out.r#type = ffi::BORDER_CIRCULAR,
    }
    out
}

/// PaddingMode -> the BorderType of (axis, side), as the Custom / Explicit drivers lower it (src/padding/mod.rs:346-452)
fn side_border<T: NumAssign + Copy, const N: usize>(pm: &PaddingMode<N, T>, axis: usize, side: usize) -> BorderType<T> {
    match *pm {
        PaddingMode::Zeros => BorderType::Zeros,
        PaddingMode::Const(c) => BorderType::Const(c),
        PaddingMode::Reflect => BorderType::Reflect,
        PaddingMode::Replicate => BorderType::Replicate,
        PaddingMode::Circular => BorderType::Circular,
        PaddingMode::Custom(bs) => bs[axis],
        PaddingMode::Explicit(bs) => bs[axis][side],
    }
}

/// One convolution fully lowered: what conv / conv_fft see after `into_kernel_with_dilation()` and `ConvMode::unfold`
/// (src/conv/mod.rs:28-66, unchanged host code).  Strides are element strides and may be negative (ndarray views).
fn lower<'a, T, S, SK, const N: usize>(
    data: &ArrayBase<S, Dim<[Ix; N]>>,
    kwd: &KernelWithDilation<'a, SK, N>,
    conv_mode: &ConvMode<N>,
    padding_mode: &PaddingMode<N, T>,
) -> ffi::ndconv_problem
where
    T: NumAssign + CudaElem,
    S: Data<Elem = T>,
    SK: Data<Elem = T>,
    Dim<[Ix; N]>: Dimension,
{
    assert!(N >= 1 && N <= ffi::MAX_DIM, "ndconv cuda: rank 1..=6");
    let cm = conv_mode.unfold(kwd);
    let mut p: ffi::ndconv_problem = unsafe { std::mem::zeroed() };
    p.dtype = T::DTYPE;
    p.ndim = N as i32;
    p.memory = ffi::MEM_HOST;
    p.reverse = kwd.reverse as i32;
    p.data = data.as_ptr() as *const c_void;
    p.kernel = kwd.kernel.as_ptr() as *const c_void;
    for i in 0..N {
        p.data_shape[i] = data.shape()[i] as i64;
        p.data_strides[i] = data.strides()[i] as i64;
        p.kernel_shape[i] = kwd.kernel.shape()[i] as i64;
        p.kernel_strides[i] = kwd.kernel.strides()[i] as i64;
        p.dilation[i] = kwd.dilation[i] as i64;
        p.pad[i] = [cm.padding[i][0] as i64, cm.padding[i][1] as i64];
        p.stride[i] = cm.strides[i] as i64;
        p.border[i] = [border(side_border(padding_mode, i, 0)), border(side_border(padding_mode, i, 1))];
    }
    p
}

/// status -> the reference's error variants (src/lib.rs:148-159).  The conv_fft quirk -- DataShape carrying the KERNEL's dim for
/// an empty kernel, src/conv_fft/mod.rs:210-213 -- is decided on the C side (status 1 with an empty kernel, non-empty data).
fn check<'a, T, S, SK, const N: usize>(
    st: c_int,
    data: &ArrayBase<S, Dim<[Ix; N]>>,
    kwd: &KernelWithDilation<'a, SK, N>,
    conv_mode: &ConvMode<N>,
) -> Result<(), Error<N>>
where
    S: Data<Elem = T>,
    SK: Data<Elem = T>,
    Dim<[Ix; N]>: Dimension,
{
    match st {
        ffi::OK => Ok(()),
        ffi::ERR_DATA_SHAPE => {
            let data_empty = data.shape().iter().product::<usize>() == 0;
            Err(Error::DataShape(if data_empty { data.raw_dim() } else { kwd.kernel.raw_dim() }))
        }
        ffi::ERR_KERNEL_SHAPE => Err(Error::KernelShape(kwd.kernel.raw_dim())),
        ffi::ERR_MISMATCH_SHAPE => {
            let kd: [Ix; N] = std::array::from_fn(|i| kwd.kernel.shape()[i] * kwd.dilation[i] - kwd.dilation[i] + 1);
            Err(Error::MismatchShape(*conv_mode, kd))
        }
        _ => panic!("ndconv cuda (status {st}): {}", last_error()),
    }
}

/// RAII page-lock of a caller-owned host allocation (ndconv_host_register / unregister)
pub struct Pinned {
    ptr: *mut c_void,
}
impl Pinned {
    pub fn new<T>(s: &[T]) -> Option<Self> {
        let ptr = s.as_ptr() as *mut c_void;
        (unsafe { ffi::ndconv_host_register(ptr, std::mem::size_of_val(s)) } == ffi::OK).then_some(Self { ptr })
    }
}

/// Allocations seen by large host calls on this thread: (pointer, bytes) -> calls seen; registered from the second sighting on.
#[derive(Default)]
struct Registrations {
    seen: Vec<(usize, usize, u32)>,
    pinned: Vec<(usize, usize, Pinned)>,
}
thread_local! {
    static REGISTRATIONS: std::cell::RefCell<Registrations> = Default::default();
}

/// Note a large contiguous host array; page-lock it when it has been seen before.  A failed registration is not an error (the
/// library stages pageable arrays through its own pinned bounce buffers).  Freshly allocated outputs are never registered.
fn note_host_array<T, S: Data<Elem = T>, D: Dimension>(a: &ArrayBase<S, D>) {
    let Some(s) = a.as_slice_memory_order() else { return };
    let (ptr, bytes) = (s.as_ptr() as usize, std::mem::size_of_val(s));
    if bytes < AUTO_REGISTER_BYTES {
        return;
    }
    REGISTRATIONS.with(|r| {
        let mut r = r.borrow_mut();
        if r.pinned.iter().any(|(p, b, _)| *p == ptr && *b == bytes) {
            return;
        }
        // an allocation that overlaps a registered one at another size was freed and reused: drop the stale registration
        r.pinned.retain(|(p, b, _)| ptr + bytes <= *p || *p + *b <= ptr);
        if let Some(e) = r.seen.iter_mut().find(|(p, b, _)| *p == ptr && *b == bytes) {
            e.2 += 1;
            if let Some(pin) = Pinned::new(s) {
                r.pinned.push((ptr, bytes, pin));
            }
        } else {
            if r.seen.len() >= 64 {
                r.seen.remove(0);
            }
            r.seen.push((ptr, bytes, 1));
        }
    });
}

/// Release every page-lock this thread holds (call before freeing long-lived input arrays that were registered automatically).
pub fn unregister_all() {
    REGISTRATIONS.with(|r| {
        let mut r = r.borrow_mut();
        r.pinned.clear();
        r.seen.clear();
    });
}
impl Drop for Pinned {
    fn drop(&mut self) {
        unsafe { ffi::ndconv_host_unregister(self.ptr) };
    }
}

fn alloc_out<T: CudaElem + NumAssign, const N: usize>(pr: &ffi::ndconv_problem, path: c_int) -> Result<Array<T, Dim<[Ix; N]>>, c_int>
where
    [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
    Dim<[Ix; N]>: Dimension,
{
    let mut shape = [0i64; ffi::MAX_DIM];
    let st = unsafe { ffi::ndconv_out_shape(pr, path, shape.as_mut_ptr()) };
    if st != ffi::OK {
        return Err(st);
    }
    let dim: [Ix; N] = std::array::from_fn(|i| shape[i] as usize);
    Ok(Array::<T, _>::zeros(dim)) // contiguous, standard layout, owned
}

// ---------------------------------------------------------------------------------------------------------------------------
// ConvExt::conv  (src/conv/mod.rs:110-115, body :128-200)  ->  ndconv_conv_direct
// ---------------------------------------------------------------------------------------------------------------------------
impl<'a, T, S, SK, const N: usize> ConvExt<'a, T, S, SK, N> for ArrayBase<S, Dim<[Ix; N]>>
where
    T: NumAssign + CudaElem + 'a,
    S: Data<Elem = T> + 'a,
    SK: Data<Elem = T> + 'a,
    Dim<[Ix; N]>: RemoveAxis,
    [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
{
    fn conv(
        &self,
        kernel: impl IntoKernelWithDilation<'a, SK, N>,
        conv_mode: ConvMode<N>,
        padding_mode: PaddingMode<N, T>,
    ) -> Result<Array<T, Dim<[Ix; N]>>, Error<N>> {
        let kwd = kernel.into_kernel_with_dilation();
        let pr = lower(self, &kwd, &conv_mode, &padding_mode);
        let mut out = match alloc_out::<T, N>(&pr, ffi::PATH_DIRECT) {
            Ok(o) => o,
            Err(st) => return check(st, self, &kwd, &conv_mode).map(|_| unreachable!()),
        };
        note_host_array(self);
        let st = DEVICES.with(|d| unsafe { ffi::ndconv_conv_direct(d.first().raw, &pr, out.as_mut_ptr() as *mut c_void) });
        check(st, self, &kwd, &conv_mode)?;
        Ok(out)
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Processor / GetProcessor  (src/conv_fft/processor/mod.rs:71-143)  ->  ndconv_processor_*, ndconv_fft_forward / backward
// ---------------------------------------------------------------------------------------------------------------------------
/// The handle `get_fft_processor::<T, InElem>()` returns under `cuda`: device, stream, plans, twiddles, cached kernel spectrum
/// and workspaces live behind it and die with it.  `Send`, never `Sync` -- the reference's own bounds (processor/mod.rs:79).
pub struct CudaProcessor<T, InElem> {
    raw: *mut ffi::ndconv_processor,
    /// the real-space shape of the last `forward` (the reference keeps `rp_origin_len`, real.rs:41,108: `backward` needs the
    /// original last-axis length, which the half spectrum does not determine)
    origin_shape: [i64; ffi::MAX_DIM],
    _m: PhantomData<(T, InElem)>,
}
unsafe impl<T, InElem> Send for CudaProcessor<T, InElem> {}

impl<T, InElem> CudaProcessor<T, InElem> {
    pub fn on_device(device: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { ffi::ndconv_processor_create(device, &mut raw) };
        assert_eq!(st, ffi::OK, "ndconv cuda: {}", last_error());
        Self { raw, origin_shape: [0; ffi::MAX_DIM], _m: PhantomData }
    }
    pub fn raw(&mut self) -> *mut ffi::ndconv_processor {
        self.raw
    }
}
impl<T, InElem> Default for CudaProcessor<T, InElem> {
    fn default() -> Self {
        Self::on_device(0)
    }
}
impl<T, InElem> Drop for CudaProcessor<T, InElem> {
    fn drop(&mut self) {
        unsafe { ffi::ndconv_processor_destroy(self.raw) };
    }
}

/// spectrum shape of the reference's processors: axis 0 moved to the end, last axis halved for real input
fn spectrum_dim<const N: usize>(shape: &[usize], real: bool) -> [Ix; N] {
    let mut s: [Ix; N] = std::array::from_fn(|i| shape[(i + 1) % N]);
    let halved = if N == 1 { 0 } else { N - 2 };
    if real {
        s[halved] = shape[N - 1] / 2 + 1;
    }
    s
}

impl<T, InElem> Processor<T, InElem> for CudaProcessor<T, InElem>
where
    T: FftNum + CudaElem,
    InElem: GetProcessor<T, InElem> + CudaElem + NumAssign,
    Complex<T>: CudaElem,
{
    fn forward<S: DataMut<Elem = InElem>, const N: usize>(
        &mut self,
        input: &mut ArrayBase<S, Dim<[Ix; N]>>,
        _parallel: bool, // the parallelism is the device's
    ) -> Array<Complex<T>, Dim<[Ix; N]>>
    where
        Dim<[Ix; N]>: RemoveAxis,
        [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
    {
        let real = InElem::DTYPE == T::DTYPE;
        let owned;
        let src: &[InElem] = match input.as_slice() {
            Some(s) => s,
            None => {
                owned = input.as_standard_layout().into_owned();
                owned.as_slice().unwrap()
            }
        };
        for i in 0..N {
            self.origin_shape[i] = input.shape()[i] as i64;
        }
        let mut out = Array::<Complex<T>, _>::zeros(spectrum_dim::<N>(input.shape(), real));
        let st = unsafe {
            ffi::ndconv_fft_forward(self.raw, InElem::DTYPE, N as c_int, self.origin_shape.as_ptr(), src.as_ptr() as *const c_void,
                                    out.as_mut_ptr() as *mut c_void, ffi::MEM_HOST)
        };
        assert_eq!(st, ffi::OK, "ndconv cuda: {}", last_error());
        out
    }

    fn backward<const N: usize>(&mut self, input: &mut Array<Complex<T>, Dim<[Ix; N]>>, _parallel: bool) -> Array<InElem, Dim<[Ix; N]>>
    where
        Dim<[Ix; N]>: RemoveAxis,
        [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
    {
        let real = InElem::DTYPE == T::DTYPE;
        // real-space shape: spectrum axes rotated back; the halved axis takes the length remembered from `forward`
        let sd = input.shape();
        let mut shape = [0i64; ffi::MAX_DIM];
        for i in 0..N {
            shape[(i + 1) % N] = sd[i] as i64;
        }
        if real {
            let remembered = self.origin_shape[N - 1];
            let half = sd[if N == 1 { 0 } else { N - 2 }] as i64;
            shape[N - 1] = if remembered / 2 + 1 == half { remembered } else { 2 * (half - 1) };
        }
        let spec = input.as_standard_layout();
        let dim: [Ix; N] = std::array::from_fn(|i| shape[i] as usize);
        let mut out = Array::<InElem, _>::zeros(dim);
        let st = unsafe {
            ffi::ndconv_fft_backward(self.raw, InElem::DTYPE, N as c_int, shape.as_ptr(), spec.as_ptr() as *const c_void,
                                     out.as_mut_ptr() as *mut c_void, ffi::MEM_HOST)
        };
        assert_eq!(st, ffi::OK, "ndconv cuda: {}", last_error());
        out
    }

    #[doc(hidden)]
    fn cuda_raw(&mut self) -> *mut ffi::ndconv_processor {
        self.raw
    }
}

macro_rules! get_processor {
    ($($t:ty),*) => { $(
        impl GetProcessor<$t, $t> for $t {
            fn get_processor() -> impl Processor<$t, $t> { CudaProcessor::<$t, $t>::default() }
        }
        impl GetProcessor<$t, Complex<$t>> for Complex<$t> {
            fn get_processor() -> impl Processor<$t, Complex<$t>> { CudaProcessor::<$t, Complex<$t>>::default() }
        }
    )* };
}
get_processor!(f32, f64); // integer `conv_fft` is documented as broken in the reference (processor/mod.rs:44-52): not offered

/// Every visible sm_100 device, one handle each, created once per thread (handles are not thread-safe).
pub struct Devices {
    procs: Vec<Handle>,
}
pub struct Handle {
    pub raw: *mut ffi::ndconv_processor,
}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { ffi::ndconv_processor_destroy(self.raw) };
    }
}
impl Devices {
    fn open() -> Self {
        let n = unsafe { ffi::ndconv_device_count() };
        assert!(n > 0, "ndconv cuda: no sm_100 device ({})", last_error());
        let procs = (0..n)
            .map(|d| {
                let mut raw = std::ptr::null_mut();
                assert_eq!(unsafe { ffi::ndconv_processor_create(d, &mut raw) }, ffi::OK, "ndconv cuda: {}", last_error());
                Handle { raw }
            })
            .collect();
        Self { procs }
    }
    pub fn first(&self) -> &Handle {
        &self.procs[0]
    }
    pub fn raws(&self) -> Vec<*mut ffi::ndconv_processor> {
        self.procs.iter().map(|h| h.raw).collect()
    }
}
thread_local! {
    pub static DEVICES: Devices = Devices::open();
}

// ---------------------------------------------------------------------------------------------------------------------------
// ConvFFTExt::{conv_fft, conv_fft_with_processor, conv_fft_par}  (src/conv_fft/mod.rs:394-423)
// ---------------------------------------------------------------------------------------------------------------------------
enum FftCall {
    /// conv_fft: a fresh processor per call in the reference (:394-402); here the C side's transient processor (NULL handle)
    Fresh,
    /// conv_fft_with_processor (:404-412): plans, twiddles and the kernel spectrum are cached in the caller's handle
    With(*mut ffi::ndconv_processor),
    /// conv_fft_par (:414-423): the rayon pool becomes the GPUs of the box -- one overlap-save slab of output rows per device
    Par,
}

fn conv_fft_impl<'a, InElem, S, SK, const N: usize>(
    data: &ArrayBase<S, Dim<[Ix; N]>>,
    kernel: impl IntoKernelWithDilation<'a, SK, N>,
    conv_mode: ConvMode<N>,
    padding_mode: PaddingMode<N, InElem>,
    call: FftCall,
) -> Result<Array<InElem, Dim<[Ix; N]>>, Error<N>>
where
    InElem: NumAssign + CudaElem + 'a,
    S: Data<Elem = InElem> + 'a,
    SK: Data<Elem = InElem> + 'a,
    Dim<[Ix; N]>: RemoveAxis,
    [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
{
    let kwd = kernel.into_kernel_with_dilation();
    let pr = lower(data, &kwd, &conv_mode, &padding_mode);
    let mut out = match alloc_out::<InElem, N>(&pr, ffi::PATH_FFT) {
        Ok(o) => o,
        Err(st) => return check(st, data, &kwd, &conv_mode).map(|_| unreachable!()),
    };
    note_host_array(data);
    let outp = out.as_mut_ptr() as *mut c_void;
    let st = match call {
        FftCall::Fresh => unsafe { ffi::ndconv_conv_fft(std::ptr::null_mut(), &pr, outp) },
        FftCall::With(raw) => unsafe { ffi::ndconv_conv_fft(raw, &pr, outp) },
        FftCall::Par => DEVICES.with(|d| {
            let raws = d.raws();
            if raws.len() > 1 {
                unsafe { ffi::ndconv_conv_fft_sharded(raws.as_ptr(), raws.len() as c_int, &pr, outp) }
            } else {
                unsafe { ffi::ndconv_conv_fft_par(raws[0], &pr, outp) }
            }
        }),
    };
    check(st, data, &kwd, &conv_mode)?;
    Ok(out) // contiguous (the reference returns a strided slice_move of the FFT-sized buffer, :282-291: same values)
}

impl<'a, T, InElem, S, SK, const N: usize> ConvFFTExt<'a, T, InElem, S, SK, N> for ArrayBase<S, Dim<[Ix; N]>>
where
    T: NumAssign + FftNum,
    InElem: GetProcessor<T, InElem> + NumAssign + CudaElem + 'a,
    S: Data<Elem = InElem> + 'a,
    SK: Data<Elem = InElem> + 'a,
    Dim<[Ix; N]>: RemoveAxis,
    [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
{
    fn conv_fft(
        &self,
        kernel: impl IntoKernelWithDilation<'a, SK, N>,
        conv_mode: ConvMode<N>,
        padding_mode: PaddingMode<N, InElem>,
    ) -> Result<Array<InElem, Dim<[Ix; N]>>, Error<N>> {
        conv_fft_impl(self, kernel, conv_mode, padding_mode, FftCall::Fresh)
    }

    fn conv_fft_with_processor(
        &self,
        kernel: impl IntoKernelWithDilation<'a, SK, N>,
        conv_mode: ConvMode<N>,
        padding_mode: PaddingMode<N, InElem>,
        fft_processor: &mut impl Processor<T, InElem>,
    ) -> Result<Array<InElem, Dim<[Ix; N]>>, Error<N>> {
        conv_fft_impl(self, kernel, conv_mode, padding_mode, FftCall::With(fft_processor.cuda_raw()))
    }

    #[cfg(feature = "rayon")]
    fn conv_fft_par(
        &self,
        kernel: impl IntoKernelWithDilation<'a, SK, N>,
        conv_mode: ConvMode<N>,
        padding_mode: PaddingMode<N, InElem>,
    ) -> Result<Array<InElem, Dim<[Ix; N]>>, Error<N>> {
        conv_fft_impl(self, kernel, conv_mode, padding_mode, FftCall::Par)
    }
}

/// A loop over independent arrays with one kernel (`for x in xs { x.conv_fft_with_processor(&k, ..) }` in the reference)
/// as one call: problem i runs whole on device i % n (ndconv_conv_fft_batch); results in input order.
pub fn conv_fft_batch<'a, InElem, S, SK, const N: usize>(
    xs: &[ArrayBase<S, Dim<[Ix; N]>>],
    kernel: impl IntoKernelWithDilation<'a, SK, N>,
    conv_mode: ConvMode<N>,
    padding_mode: PaddingMode<N, InElem>,
) -> Result<Vec<Array<InElem, Dim<[Ix; N]>>>, Error<N>>
where
    InElem: NumAssign + CudaElem + 'a,
    S: Data<Elem = InElem> + 'a,
    SK: Data<Elem = InElem> + 'a,
    Dim<[Ix; N]>: RemoveAxis,
    [Ix; N]: IntoDimension<Dim = Dim<[Ix; N]>>,
{
    let kwd = kernel.into_kernel_with_dilation();
    let mut problems = Vec::with_capacity(xs.len());
    let mut outs = Vec::with_capacity(xs.len());
    for x in xs {
        let pr = lower(x, &kwd, &conv_mode, &padding_mode);
        match alloc_out::<InElem, N>(&pr, ffi::PATH_FFT) {
            Ok(o) => outs.push(o),
            Err(st) => return check(st, x, &kwd, &conv_mode).map(|_| unreachable!()),
        }
        problems.push(pr);
    }
    let ptrs: Vec<*mut c_void> = outs.iter_mut().map(|o| o.as_mut_ptr() as *mut c_void).collect();
    let st = DEVICES.with(|d| {
        let raws = d.raws();
        unsafe { ffi::ndconv_conv_fft_batch(raws.as_ptr(), raws.len() as c_int, problems.as_ptr(), ptrs.as_ptr(), problems.len() as c_int) }
    });
    if let Some(x) = xs.first() {
        check(st, x, &kwd, &conv_mode)?;
    }
    Ok(outs)
}
