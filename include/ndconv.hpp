// ndconv.hpp -- header-only C++17 mirror of TYPEmber/ndarray-conv's trait API over the C ABI of include/ndconv.h.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the host side above the C ABI is C++ (this
// header; tests/cpp/ reads like the reference's own unit tests) plus the ctypes mirror used by pytest.  Names, argument
// meaning and error behaviour follow the crate (file:line relative to the reference root):
//
//   ConvMode<N>::{Full,Same,Valid,Custom,Explicit}                src/lib.rs:80-105
//   PaddingMode<N,T>::{Zeros,Const,Reflect,Replicate,Circular,Custom,Explicit}   src/lib.rs:111-127
//   BorderType<T>::{Zeros,Const,Reflect,Replicate,Circular}       src/lib.rs:131-143
//   Error{DataShape,KernelShape,MismatchShape}                    src/lib.rs:148-159   (thrown as ndconv::Error)
//   with_dilation(kernel, d) / .reverse() / .no_reverse()         src/dilation/mod.rs:103-189
//   conv(x, kernel, mode, padding)                                src/conv/mod.rs:110-115       (ConvExt::conv)
//   conv_fft / conv_fft_with_processor / conv_fft_par             src/conv_fft/mod.rs:113-171   (ConvFFTExt)
//   get_fft_processor()                                           src/conv_fft/processor/mod.rs:71-73
//
// Anything the reference answers with a panic (and every CUDA failure) is thrown as ndconv::Panic.  There is no CPU path.
#pragma once
#include <array>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ndconv.h"

namespace ndconv {

template <class T> struct dtype_of;
#define NDCONV_DTYPE(T, code) template <> struct dtype_of<T> { static constexpr int value = code; };
NDCONV_DTYPE(int32_t, NDCONV_I32) NDCONV_DTYPE(int64_t, NDCONV_I64) NDCONV_DTYPE(float, NDCONV_F32) NDCONV_DTYPE(double, NDCONV_F64)
NDCONV_DTYPE(std::complex<float>, NDCONV_C32) NDCONV_DTYPE(std::complex<double>, NDCONV_C64)
NDCONV_DTYPE(int8_t, NDCONV_I8) NDCONV_DTYPE(int16_t, NDCONV_I16) NDCONV_DTYPE(uint8_t, NDCONV_U8) NDCONV_DTYPE(uint16_t, NDCONV_U16)
NDCONV_DTYPE(uint32_t, NDCONV_U32) NDCONV_DTYPE(uint64_t, NDCONV_U64)
#undef NDCONV_DTYPE

template <class T> struct scalar_of { using type = T; static constexpr bool is_complex = false; };
template <class R> struct scalar_of<std::complex<R>> { using type = R; static constexpr bool is_complex = true; };

enum class ErrorKind { DataShape = 1, KernelShape = 2, MismatchShape = 3 };

// Error<N>, src/lib.rs:148-159
struct Error : std::runtime_error {
    ErrorKind kind;
    Error(ErrorKind k, const std::string &m) : std::runtime_error(m), kind(k) {}
};
// what the reference would answer with panic!() -- plus CUDA failures (Error<N> is not #[non_exhaustive])
struct Panic : std::runtime_error {
    int status;
    Panic(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};

inline void check(int st)
{
    if (st == NDCONV_OK) return;
    const std::string msg = ndconv_last_error_string();
    if (st >= 1 && st <= 3) throw Error(static_cast<ErrorKind>(st), msg);
    throw Panic(st, msg);
}

// ---- arrays: the minimum of ndarray's Array / ArrayView the API needs ---------------------------------------
template <class T, size_t N> struct ArrayView {
    const T *ptr = nullptr;
    std::array<size_t, N> shape{};
    std::array<ptrdiff_t, N> strides{};     // in elements, may be negative (ndarray views)
};

template <class T, size_t N> struct Array {
    std::vector<T> data;
    std::array<size_t, N> shape{};
    Array() = default;
    explicit Array(const std::array<size_t, N> &s) : shape(s) { size_t n = 1; for (auto v : s) n *= v; data.assign(n, T{}); }
    Array(const std::array<size_t, N> &s, std::vector<T> d) : data(std::move(d)), shape(s) {}
    size_t len() const { return data.size(); }
    ArrayView<T, N> view() const
    {
        ArrayView<T, N> v; v.ptr = data.data(); v.shape = shape;
        ptrdiff_t st = 1;
        for (size_t i = N; i-- > 0;) { v.strides[i] = st; st *= (ptrdiff_t)shape[i]; }
        return v;
    }
    bool operator==(const Array &o) const { return shape == o.shape && data == o.data; }
};
template <class T> Array<T, 1> array1(std::vector<T> v) { const size_t n = v.size(); return Array<T, 1>({n}, std::move(v)); }
template <class T> Array<T, 2> array2(size_t r, size_t c, std::vector<T> v) { return Array<T, 2>({r, c}, std::move(v)); }
template <class T> Array<T, 3> array3(size_t a, size_t b, size_t c, std::vector<T> v) { return Array<T, 3>({a, b, c}, std::move(v)); }

// ---- enums ---------------------------------------------------------------------------------------------------
template <class T> struct BorderType {
    int kind = NDCONV_BORDER_ZEROS;
    T value{};
    static BorderType Zeros() { return {NDCONV_BORDER_ZEROS, T{}}; }
    static BorderType Const(T c) { return {NDCONV_BORDER_CONST, c}; }
    static BorderType Reflect() { return {NDCONV_BORDER_REFLECT, T{}}; }
    static BorderType Replicate() { return {NDCONV_BORDER_REPLICATE, T{}}; }
    static BorderType Circular() { return {NDCONV_BORDER_CIRCULAR, T{}}; }
};

template <size_t N, class T> struct PaddingMode {
    std::array<std::array<BorderType<T>, 2>, N> sides{};    // already lowered per side, as the Custom / Explicit drivers do (src/padding/mod.rs:346-452)
    static PaddingMode all(BorderType<T> b) { PaddingMode p; for (auto &s : p.sides) s = {b, b}; return p; }
    static PaddingMode Zeros() { return all(BorderType<T>::Zeros()); }
    static PaddingMode Const(T c) { return all(BorderType<T>::Const(c)); }
    static PaddingMode Reflect() { return all(BorderType<T>::Reflect()); }
    static PaddingMode Replicate() { return all(BorderType<T>::Replicate()); }
    static PaddingMode Circular() { return all(BorderType<T>::Circular()); }
    static PaddingMode Custom(const std::array<BorderType<T>, N> &b) { PaddingMode p; for (size_t i = 0; i < N; i++) p.sides[i] = {b[i], b[i]}; return p; }
    static PaddingMode Explicit(const std::array<std::array<BorderType<T>, 2>, N> &b) { PaddingMode p; p.sides = b; return p; }
};

template <size_t N> struct ConvMode {
    int kind = NDCONV_MODE_SAME;
    std::array<size_t, N> padding{};
    std::array<std::array<size_t, 2>, N> explicit_padding{};
    std::array<size_t, N> strides{};
    static ConvMode Full() { ConvMode m; m.kind = NDCONV_MODE_FULL; return m; }
    static ConvMode Same() { ConvMode m; m.kind = NDCONV_MODE_SAME; return m; }
    static ConvMode Valid() { ConvMode m; m.kind = NDCONV_MODE_VALID; return m; }
    static ConvMode Custom(const std::array<size_t, N> &p, const std::array<size_t, N> &s) { ConvMode m; m.kind = NDCONV_MODE_CUSTOM; m.padding = p; m.strides = s; return m; }
    static ConvMode Explicit(const std::array<std::array<size_t, 2>, N> &p, const std::array<size_t, N> &s) { ConvMode m; m.kind = NDCONV_MODE_EXPLICIT; m.explicit_padding = p; m.strides = s; return m; }
};

// KernelWithDilation, src/dilation/mod.rs:8-12 (reverse defaults to true = mathematical convolution)
template <class T, size_t N> struct KernelWithDilation {
    ArrayView<T, N> kernel;
    std::array<size_t, N> dilation{};
    bool reverse_flag = true;
    KernelWithDilation reverse() const { auto k = *this; k.reverse_flag = true; return k; }
    KernelWithDilation no_reverse() const { auto k = *this; k.reverse_flag = false; return k; }
};
template <class T, size_t N> KernelWithDilation<T, N> with_dilation(const Array<T, N> &k, size_t d)
{
    KernelWithDilation<T, N> r; r.kernel = k.view(); r.dilation.fill(d); return r;
}
template <class T, size_t N> KernelWithDilation<T, N> with_dilation(const Array<T, N> &k, const std::array<size_t, N> &d)
{
    KernelWithDilation<T, N> r; r.kernel = k.view(); r.dilation = d; return r;
}
// IntoKernelWithDilation for a bare array: dilation 1, reverse (src/dilation/mod.rs:196-203)
template <class T, size_t N> KernelWithDilation<T, N> into_kernel_with_dilation(const Array<T, N> &k) { return with_dilation(k, size_t(1)); }
template <class T, size_t N> KernelWithDilation<T, N> into_kernel_with_dilation(const KernelWithDilation<T, N> &k) { return k; }

// ---- processor -------------------------------------------------------------------------------------------------
class FftProcessor {
public:
    explicit FftProcessor(int device = 0) { check(ndconv_processor_create(device, &p_)); }
    ~FftProcessor() { if (p_) ndconv_processor_destroy(p_); }
    FftProcessor(const FftProcessor &) = delete;
    FftProcessor &operator=(const FftProcessor &) = delete;
    FftProcessor(FftProcessor &&o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    ndconv_processor *raw() { return p_; }
    int64_t launch_count() const { return ndconv_processor_launch_count(p_); }
    // Processor::forward / backward (src/conv_fft/processor/mod.rs:91-118): unnormalised N-d FFT, spectrum with axis 0 moved to
    // the end; backward divides by the element count and reuses the shape of the last forward (rp_origin_len).
    template <class T, size_t N> Array<std::complex<typename scalar_of<T>::type>, N> forward(const Array<T, N> &x)
    {
        using C = std::complex<typename scalar_of<T>::type>;
        constexpr bool cplx = scalar_of<T>::is_complex;
        std::array<size_t, N> os{};
        const size_t last = cplx ? x.shape[N - 1] : x.shape[N - 1] / 2 + 1;
        if (N == 1) os[0] = last;
        else { for (size_t i = 1; i + 1 < N; i++) os[i - 1] = x.shape[i]; os[N - 2] = last; os[N - 1] = x.shape[0]; }
        Array<C, N> out(os);
        int64_t shp[NDCONV_MAX_DIM];
        origin_.assign(x.shape.begin(), x.shape.end());
        for (size_t i = 0; i < N; i++) shp[i] = (int64_t)x.shape[i];
        check(ndconv_fft_forward(p_, dtype_of<T>::value, (int)N, shp, x.data.data(), out.data.data(), NDCONV_MEM_HOST));
        return out;
    }
    template <class T, size_t N> Array<T, N> backward(const Array<std::complex<typename scalar_of<T>::type>, N> &spectrum)
    {
        if (origin_.size() != N) throw Panic(NDCONV_ERR_BAD_ARG, "backward() before forward()");
        std::array<size_t, N> s{};
        int64_t shp[NDCONV_MAX_DIM];
        for (size_t i = 0; i < N; i++) { s[i] = origin_[i]; shp[i] = (int64_t)origin_[i]; }
        Array<T, N> out(s);
        check(ndconv_fft_backward(p_, dtype_of<T>::value, (int)N, shp, spectrum.data.data(), out.data.data(), NDCONV_MEM_HOST));
        return out;
    }
private:
    ndconv_processor *p_ = nullptr;
    std::vector<size_t> origin_;
};
inline FftProcessor get_fft_processor(int device = 0) { return FftProcessor(device); }

// ---- lowering ----------------------------------------------------------------------------------------------------
template <class T, size_t N>
ndconv_problem lower(const ArrayView<T, N> &x, const KernelWithDilation<T, N> &kwd, const ConvMode<N> &mode, const PaddingMode<N, T> &pm)
{
    static_assert(N >= 1 && N <= NDCONV_MAX_DIM, "rank 1..6");
    ndconv_problem pr;
    std::memset(&pr, 0, sizeof(pr));
    pr.dtype = dtype_of<T>::value; pr.ndim = (int)N; pr.memory = NDCONV_MEM_HOST; pr.reverse = kwd.reverse_flag ? 1 : 0;
    pr.data = x.ptr; pr.kernel = kwd.kernel.ptr;
    int64_t kshape[N], dil[N], pad_in[2 * N], str_in[N];
    for (size_t i = 0; i < N; i++) {
        pr.data_shape[i] = (int64_t)x.shape[i]; pr.data_strides[i] = (int64_t)x.strides[i];
        pr.kernel_shape[i] = kshape[i] = (int64_t)kwd.kernel.shape[i]; pr.kernel_strides[i] = (int64_t)kwd.kernel.strides[i];
        pr.dilation[i] = dil[i] = (int64_t)kwd.dilation[i];
        str_in[i] = (int64_t)mode.strides[i];
        if (mode.kind == NDCONV_MODE_CUSTOM) pad_in[i] = (int64_t)mode.padding[i];
    }
    if (mode.kind == NDCONV_MODE_EXPLICIT) for (size_t i = 0; i < N; i++) { pad_in[2 * i] = (int64_t)mode.explicit_padding[i][0]; pad_in[2 * i + 1] = (int64_t)mode.explicit_padding[i][1]; }
    check(ndconv_unfold_conv_mode(mode.kind, (int)N, kshape, dil, pad_in, str_in, pr.pad, pr.stride));     // ConvMode::unfold, src/conv/mod.rs:28-66
    for (size_t i = 0; i < N; i++)
        for (int s = 0; s < 2; s++) {
            pr.border[i][s].type = pm.sides[i][s].kind;
            std::memcpy(pr.border[i][s].value, &pm.sides[i][s].value, sizeof(T));
        }
    return pr;
}

template <class T, size_t N, class Fn> Array<T, N> run(const ndconv_problem &pr, int path, Fn &&call)
{
    int64_t shp[NDCONV_MAX_DIM];
    check(ndconv_out_shape(&pr, path, shp));
    std::array<size_t, N> s;
    for (size_t i = 0; i < N; i++) s[i] = (size_t)shp[i];
    Array<T, N> out(s);
    check(call(out.data.data()));
    return out;
}

// ---- the API -------------------------------------------------------------------------------------------------------
// ConvExt::conv
template <class T, size_t N, class K>
Array<T, N> conv(const Array<T, N> &x, const K &kernel, const ConvMode<N> &mode, const PaddingMode<N, T> &pm)
{
    const auto pr = lower(x.view(), into_kernel_with_dilation(kernel), mode, pm);
    return run<T, N>(pr, NDCONV_PATH_DIRECT, [&](void *o) { return ndconv_conv_direct(nullptr, &pr, o); });
}
// ConvFFTExt::conv_fft (fresh processor per call)
template <class T, size_t N, class K>
Array<T, N> conv_fft(const Array<T, N> &x, const K &kernel, const ConvMode<N> &mode, const PaddingMode<N, T> &pm)
{
    const auto pr = lower(x.view(), into_kernel_with_dilation(kernel), mode, pm);
    return run<T, N>(pr, NDCONV_PATH_FFT, [&](void *o) { return ndconv_conv_fft(nullptr, &pr, o); });
}
// ConvFFTExt::conv_fft_with_processor
template <class T, size_t N, class K>
Array<T, N> conv_fft_with_processor(const Array<T, N> &x, const K &kernel, const ConvMode<N> &mode, const PaddingMode<N, T> &pm, FftProcessor &proc)
{
    const auto pr = lower(x.view(), into_kernel_with_dilation(kernel), mode, pm);
    return run<T, N>(pr, NDCONV_PATH_FFT, [&](void *o) { return ndconv_conv_fft(proc.raw(), &pr, o); });
}
// ConvFFTExt::conv_fft_par
template <class T, size_t N, class K>
Array<T, N> conv_fft_par(const Array<T, N> &x, const K &kernel, const ConvMode<N> &mode, const PaddingMode<N, T> &pm)
{
    const auto pr = lower(x.view(), into_kernel_with_dilation(kernel), mode, pm);
    return run<T, N>(pr, NDCONV_PATH_FFT, [&](void *o) { return ndconv_conv_fft_par(nullptr, &pr, o); });
}

// conv_fft_par with several GPUs configured (ndconv_conv_fft_sharded): one axis-0 slab of output rows per processor
template <class T, size_t N, class K>
Array<T, N> conv_fft_par(const Array<T, N> &x, const K &kernel, const ConvMode<N> &mode, const PaddingMode<N, T> &pm, const std::vector<FftProcessor *> &procs)
{
    const auto pr = lower(x.view(), into_kernel_with_dilation(kernel), mode, pm);
    std::vector<ndconv_processor *> raw;
    for (FftProcessor *p : procs) raw.push_back(p->raw());
    return run<T, N>(pr, NDCONV_PATH_FFT, [&](void *o) { return ndconv_conv_fft_sharded(raw.data(), (int)raw.size(), &pr, o); });
}

// a batch of independent convolutions with one kernel, distributed whole over the processors (ndconv_conv_fft_batch)
template <class T, size_t N, class K>
std::vector<Array<T, N>> conv_fft_batch(const std::vector<Array<T, N>> &xs, const K &kernel, const ConvMode<N> &mode, const PaddingMode<N, T> &pm,
                                        const std::vector<FftProcessor *> &procs)
{
    const auto kwd = into_kernel_with_dilation(kernel);
    std::vector<ndconv_problem> prs;
    std::vector<Array<T, N>> outs;
    std::vector<void *> optr;
    for (const auto &x : xs) {
        prs.push_back(lower(x.view(), kwd, mode, pm));
        int64_t shp[NDCONV_MAX_DIM];
        check(ndconv_out_shape(&prs.back(), NDCONV_PATH_FFT, shp));
        std::array<size_t, N> sh;
        for (size_t i = 0; i < N; i++) sh[i] = (size_t)shp[i];
        outs.emplace_back(sh);
    }
    for (auto &o : outs) optr.push_back(o.data.data());
    std::vector<ndconv_processor *> raw;
    for (FftProcessor *p : procs) raw.push_back(p->raw());
    check(ndconv_conv_fft_batch(raw.data(), (int)raw.size(), prs.data(), optr.data(), (int)prs.size()));
    return outs;
}

}  // namespace ndconv
