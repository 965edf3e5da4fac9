// C++ host-mirror tests: the same cases, in the same shape, as the reference's own unit tests
// (src/conv/tests.rs, src/conv_fft/tests.rs), driven through include/ndconv.hpp -> C ABI -> sm_100a kernels.
// Expected values of the libtorch-derived tests are the ones pinned in tests/golden/torch_golden.json.
//   ./test_ndconv_hpp              all tests (needs a B200)
//   ./test_ndconv_hpp --host-only  only the checks that never reach the device (error order, lowering)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>

#include "../../include/ndconv.hpp"

using namespace ndconv;
static int failures = 0, ran = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("  FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)
static void run_test(const char *name, const std::function<void()> &f)
{
    ran++;
    int before = failures;
    try { f(); } catch (const std::exception &e) { std::printf("  EXCEPTION in %s: %s\n", name, e.what()); failures++; }
    std::printf("%s %s\n", failures == before ? "ok  " : "FAIL", name);
}
template <class T, size_t N> static Array<float, N> to_f32(const Array<T, N> &a) { Array<float, N> r(a.shape); for (size_t i = 0; i < a.len(); i++) r.data[i] = (float)a.data[i]; return r; }
template <class T, size_t N> static Array<double, N> to_f64(const Array<T, N> &a) { Array<double, N> r(a.shape); for (size_t i = 0; i < a.len(); i++) r.data[i] = (double)a.data[i]; return r; }
// assert_fft_matches_conv_f32 / _f64: |round(fft) - conv| < tol  (src/conv_fft/tests.rs:14-78)
template <class F, size_t N> static bool fft_matches_conv(const Array<F, N> &fft, const Array<int32_t, N> &conv_res, double tol)
{
    if (fft.shape != conv_res.shape) return false;
    for (size_t i = 0; i < fft.len(); i++) if (std::fabs(std::round((double)fft.data[i]) - (double)conv_res.data[i]) >= tol) return false;
    return true;
}

static void device_tests()
{
    using V = std::vector<int32_t>;
    // ---- src/conv/tests.rs vs_torch ----
    run_test("conv::full_mode::test_1d", [] {
        auto arr = array1<int32_t>({1, 2, 3, 4, 5}); auto kernel = array1<int32_t>({1, 2, 1});
        auto res = conv(arr, kernel, ConvMode<1>::Full(), PaddingMode<1, int32_t>::Zeros());
        CHECK(res.data == V({1, 4, 8, 12, 16, 14, 5}));
    });
    run_test("conv::full_mode::test_2d", [] {
        auto arr = array2<int32_t>(2, 2, {1, 2, 3, 4}); auto kernel = array2<int32_t>(2, 2, {1, 1, 1, 1});
        auto res = conv(arr, kernel, ConvMode<2>::Full(), PaddingMode<2, int32_t>::Zeros());
        CHECK((res.shape == std::array<size_t, 2>{3, 3})); CHECK(res.data == V({1, 3, 2, 4, 10, 6, 3, 7, 4}));
    });
    run_test("conv::full_mode::test_3d", [] {
        auto arr = array3<int32_t>(2, 1, 2, {1, 2, 3, 4}); auto kernel = array3<int32_t>(2, 1, 2, {1, 1, 1, 1});
        auto res = conv(arr, kernel, ConvMode<3>::Full(), PaddingMode<3, int32_t>::Zeros());
        CHECK((res.shape == std::array<size_t, 3>{3, 1, 3})); CHECK(res.data == V({1, 3, 2, 4, 10, 6, 3, 7, 4}));
    });
    run_test("conv::same_mode::test_2d (sobel, kernel reversed by default)", [] {
        auto arr = array2<int32_t>(3, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}); auto kernel = array2<int32_t>(3, 3, {1, 0, -1, 2, 0, -2, 1, 0, -1});
        auto res = conv(arr, kernel, ConvMode<2>::Same(), PaddingMode<2, int32_t>::Zeros());
        CHECK(res.data == V({9, 6, -9, 20, 8, -20, 21, 6, -21}));
    });
    run_test("conv::valid_mode::test_3d", [] {
        auto arr = array3<int32_t>(2, 2, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}); auto kernel = array3<int32_t>(2, 1, 2, {1, 1, 1, 1});
        auto res = conv(arr, kernel, ConvMode<3>::Valid(), PaddingMode<3, int32_t>::Zeros());
        CHECK((res.shape == std::array<size_t, 3>{1, 2, 2})); CHECK(res.data == V({18, 22, 30, 34}));
    });
    run_test("conv::with_strides::stride_2_2d", [] {
        auto arr = array2<int32_t>(3, 4, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}); auto kernel = array2<int32_t>(2, 2, {1, 1, 1, 1});
        auto res = conv(arr, kernel, ConvMode<2>::Custom({1, 1}, {2, 2}), PaddingMode<2, int32_t>::Zeros());
        CHECK((res.shape == std::array<size_t, 2>{2, 3})); CHECK(res.data == V({1, 5, 4, 14, 34, 20}));
    });
    run_test("conv::with_strides::stride_3_1d", [] {
        auto arr = array1<int32_t>({1, 2, 3, 4, 5, 6, 7, 8, 9}); auto kernel = array1<int32_t>({1, 2, 1});
        auto res = conv(arr, kernel, ConvMode<1>::Custom({2}, {3}), PaddingMode<1, int32_t>::Zeros());
        CHECK(res.data == V({1, 12, 24, 26}));
    });
    run_test("conv::with_dilation::dilation_2_1d (no_reverse)", [] {
        auto arr = array1<int32_t>({1, 2, 3, 4, 5, 6}); auto kernel = array1<int32_t>({1, 1, 2});
        auto res = conv(arr, with_dilation(kernel, 2).no_reverse(), ConvMode<1>::Custom({4}, {2}), PaddingMode<1, int32_t>::Zeros());
        CHECK(res.data == V({2, 7, 14, 8, 5}));
    });
    run_test("conv::with_dilation::dilation_2_3d (no_reverse)", [] {
        auto arr = array3<int32_t>(2, 2, 2, {1, 2, 3, 4, 5, 6, 7, 8}); auto kernel = array3<int32_t>(2, 3, 3, std::vector<int32_t>(18, 1));
        auto res = conv(arr, with_dilation(kernel, 2).no_reverse(), ConvMode<3>::Custom({2, 2, 2}, {1, 2, 1}), PaddingMode<3, int32_t>::Zeros());
        CHECK((res.shape == std::array<size_t, 3>{4, 1, 2})); CHECK(res.data == V({1, 2, 5, 6, 1, 2, 5, 6}));
    });
    run_test("conv::reverse_kernel::with_reverse", [] {
        auto arr = array1<int32_t>({1, 2, 3, 4, 5, 6}); auto kernel = array1<int32_t>({1, 1, 2});
        auto res = conv(arr, with_dilation(kernel, 2), ConvMode<1>::Custom({4}, {2}), PaddingMode<1, int32_t>::Zeros());
        CHECK(res.data == V({1, 4, 10, 11, 10}));
    });
    run_test("conv::edge_cases (single_element_array, single_element_kernel, identity_kernel)", [] {
        CHECK(conv(array1<int32_t>({42}), array1<int32_t>({2}), ConvMode<1>::Same(), PaddingMode<1, int32_t>::Zeros()).data == V({84}));
        CHECK(conv(array1<int32_t>({1, 2, 3, 4}), array1<int32_t>({3}), ConvMode<1>::Same(), PaddingMode<1, int32_t>::Zeros()).data == V({3, 6, 9, 12}));
        auto arr = array2<int32_t>(2, 2, {1, 2, 3, 4});
        CHECK(conv(arr, array2<int32_t>(1, 1, {1}), ConvMode<2>::Same(), PaddingMode<2, int32_t>::Zeros()) == arr);
    });
    // ---- src/conv_fft/tests.rs vs_conv ----
    run_test("conv_fft::one_d::same_mode_f32 (dilation 2)", [] {
        auto arr = array1<int32_t>({1, 2, 3, 4, 5, 6}); auto kernel = array1<int32_t>({1, 1, 1, 1});
        auto c = conv(arr, with_dilation(kernel, 2), ConvMode<1>::Same(), PaddingMode<1, int32_t>::Zeros());
        CHECK(c.data == V({6, 9, 12, 9, 12, 8}));
        auto kf = to_f32(kernel);
        auto f = conv_fft(to_f32(arr), with_dilation(kf, 2), ConvMode<1>::Same(), PaddingMode<1, float>::Zeros());
        CHECK(fft_matches_conv(f, c, 1e-5));
    });
    run_test("conv_fft::one_d::circular_padding (float data, |conv - fft| < 1e-6)", [] {
        auto arr = array1<float>({0.0f, 0.1f, 0.3f, 0.4f, 0.0f, 0.1f, 0.3f, 0.4f, 0.0f, 0.1f, 0.3f, 0.4f, 0.0f, 0.1f, 0.3f, 0.4f});
        auto kernel = array1<float>({0.1f, 0.3f, 0.6f, 0.3f, 0.1f});
        auto c = conv(arr, kernel, ConvMode<1>::Same(), PaddingMode<1, float>::Circular());
        auto f = conv_fft(arr, kernel, ConvMode<1>::Same(), PaddingMode<1, float>::Circular());
        for (size_t i = 0; i < c.len(); i++) CHECK(std::fabs(c.data[i] - f.data[i]) < 1e-6f);
    });
    run_test("conv_fft::two_d::same_mode_f32 (Replicate)", [] {
        auto arr = array2<int32_t>(6, 2, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}); auto kernel = array2<int32_t>(2, 2, {1, 0, 3, 1});
        auto c = conv(arr, kernel, ConvMode<2>::Same(), PaddingMode<2, int32_t>::Replicate());
        CHECK(c.data == V({5, 9, 7, 11, 17, 21, 27, 31, 37, 41, 47, 51}));
        CHECK(fft_matches_conv(conv_fft(to_f32(arr), to_f32(kernel), ConvMode<2>::Same(), PaddingMode<2, float>::Replicate()), c, 1e-5));
    });
    run_test("conv_fft::two_d::custom_mode_with_dilation (f64, no_reverse, stride 2, Replicate)", [] {
        auto arr = array2<int32_t>(2, 2, {1, 2, 3, 4}); auto kernel = array2<int32_t>(2, 2, {1, 0, 3, 1});
        auto c = conv(arr, with_dilation(kernel, 2).no_reverse(), ConvMode<2>::Custom({3, 3}, {2, 2}), PaddingMode<2, int32_t>::Replicate());
        CHECK(c.data == V({5, 6, 10, 13, 14, 18, 15, 16, 20}));
        auto kd = to_f64(kernel);
        CHECK(fft_matches_conv(conv_fft(to_f64(arr), with_dilation(kd, 2).no_reverse(), ConvMode<2>::Custom({3, 3}, {2, 2}), PaddingMode<2, double>::Replicate()), c, 1e-9));
    });
    run_test("conv_fft::three_d::same_mode_f32", [] {
        auto arr = array3<int32_t>(2, 2, 2, {1, 2, 3, 4, 5, 6, 7, 8}); auto kernel = array3<int32_t>(2, 3, 3, std::vector<int32_t>(18, 1));
        auto c = conv(arr, kernel, ConvMode<3>::Same(), PaddingMode<3, int32_t>::Zeros());
        CHECK(c.data == V({10, 10, 10, 10, 36, 36, 36, 36}));
        CHECK(fft_matches_conv(conv_fft(to_f32(arr), to_f32(kernel), ConvMode<3>::Same(), PaddingMode<3, float>::Zeros()), c, 1e-5));
    });
    run_test("conv_fft::padding_modes::const_padding_2d (Const(7), Full)", [] {
        auto arr = array2<int32_t>(2, 2, {1, 2, 3, 4}); auto kernel = array2<int32_t>(2, 2, {1, 1, 1, 1});
        auto c = conv(arr, kernel, ConvMode<2>::Full(), PaddingMode<2, int32_t>::Const(7));
        CHECK(c.data == V({22, 17, 23, 18, 10, 20, 24, 21, 25}));
        CHECK(fft_matches_conv(conv_fft(to_f32(arr), to_f32(kernel), ConvMode<2>::Full(), PaddingMode<2, float>::Const(7.0f)), c, 1e-5));
    });
    run_test("conv_fft::same_mode_complex (vs complex conv; the reference only self-compares)", [] {
        using C = std::complex<float>;
        auto arr = array1<C>({{1, 1}, {2, -1}, {3, 0.5f}, {4, 0}, {5, 2}, {6, -2}}); auto kernel = array1<C>({{1, 0}, {0, 1}, {1, 1}, {2, -1}});
        auto c = conv(arr, with_dilation(kernel, 2), ConvMode<1>::Same(), PaddingMode<1, C>::Zeros());
        auto f = conv_fft(arr, with_dilation(kernel, 2), ConvMode<1>::Same(), PaddingMode<1, C>::Zeros());
        auto f2 = conv_fft(arr, with_dilation(kernel, 2), ConvMode<1>::Same(), PaddingMode<1, C>::Zeros());
        CHECK(f.shape == c.shape);
        for (size_t i = 0; i < c.len(); i++) { CHECK(std::abs(c.data[i] - f.data[i]) < 1e-4f); CHECK(f.data[i] == f2.data[i]); }
    });
    run_test("conv_fft_par vs conv_fft (two_d_same_f32, modes Same/Valid/Full), with_processor reuse", [] {
        Array<float, 2> arr({32, 32});
        for (size_t i = 0; i < 32; i++) for (size_t j = 0; j < 32; j++) arr.data[i * 32 + j] = (float)((i + j) % 10);
        auto ker = array2<float>(3, 3, {1, 2, 1, 0, 0, 0, -1, -2, -1});
        auto proc = get_fft_processor();
        for (auto mode : {ConvMode<2>::Same(), ConvMode<2>::Valid(), ConvMode<2>::Full()}) {
            auto serial = conv_fft(arr, ker, mode, PaddingMode<2, float>::Zeros());
            auto par = conv_fft_par(arr, ker, mode, PaddingMode<2, float>::Zeros());
            auto wp = conv_fft_with_processor(arr, ker, mode, PaddingMode<2, float>::Zeros(), proc);
            CHECK(serial.shape == par.shape);
            for (size_t i = 0; i < serial.len(); i++) { CHECK(std::fabs(serial.data[i] - par.data[i]) < 1e-4f); CHECK(std::fabs(serial.data[i] - wp.data[i]) < 1e-4f); }
        }
        CHECK(proc.launch_count() > 0);
        // several processors configured: ndconv_conv_fft_sharded (a problem this small runs on the first one)
        auto proc2 = get_fft_processor();
        auto multi = conv_fft_par(arr, ker, ConvMode<2>::Same(), PaddingMode<2, float>::Zeros(), std::vector<FftProcessor *>{&proc, &proc2});
        auto serial = conv_fft(arr, ker, ConvMode<2>::Same(), PaddingMode<2, float>::Zeros());
        for (size_t i = 0; i < serial.len(); i++) CHECK(std::fabs(serial.data[i] - multi.data[i]) < 1e-4f);
        // independent problems distributed whole over the processors: ndconv_conv_fft_batch
        std::vector<Array<float, 2>> xs;
        for (int b = 0; b < 5; b++) { Array<float, 2> a({20, 24}); for (size_t i = 0; i < a.len(); i++) a.data[i] = (float)((i * (b + 3)) % 11); xs.push_back(a); }
        auto ys = conv_fft_batch(xs, ker, ConvMode<2>::Full(), PaddingMode<2, float>::Zeros(), std::vector<FftProcessor *>{&proc, &proc2});
        CHECK(ys.size() == xs.size());
        for (size_t b = 0; b < xs.size(); b++) {
            auto one = conv_fft(xs[b], ker, ConvMode<2>::Full(), PaddingMode<2, float>::Zeros());
            CHECK(one.shape == ys[b].shape);
            for (size_t i = 0; i < one.len(); i++) CHECK(std::fabs(one.data[i] - ys[b].data[i]) < 1e-4f);
        }
    });
    run_test("processor::real::round_trip_2d (forward o backward = id, rotated layout [n1/2+1, n0])", [] {
        Array<double, 2> x({6, 10});
        for (size_t i = 0; i < x.len(); i++) x.data[i] = std::sin(0.37 * (double)i) + 0.01 * (double)i;
        auto proc = get_fft_processor();
        auto spec = proc.forward(x);
        CHECK((spec.shape == std::array<size_t, 2>{6, 6}));          // [10/2+1, 6]
        double dc = 0; for (double v : x.data) dc += v;
        CHECK(std::abs(spec.data[0] - std::complex<double>(dc, 0)) < 1e-10);
        auto back = proc.backward<double, 2>(spec);
        for (size_t i = 0; i < x.len(); i++) CHECK(std::fabs(back.data[i] - x.data[i]) < 1e-10);
    });
}

static void host_tests()
{
    // error order and variants: never reaches the device
    run_test("errors: DataShape / KernelShape / MismatchShape (conv) and the conv_fft DataShape quirk", [] {
        auto expect = [](std::function<void()> f, ErrorKind k) { try { f(); } catch (const Error &e) { return e.kind == k; } catch (...) { return false; } return false; };
        Array<int32_t, 2> empty({0, 3}); auto k11 = array2<int32_t>(1, 1, {1}); Array<int32_t, 2> kempty({0, 1}); auto x22 = array2<int32_t>(2, 2, {1, 2, 3, 4});
        CHECK(expect([&] { conv(empty, k11, ConvMode<2>::Same(), PaddingMode<2, int32_t>::Zeros()); }, ErrorKind::DataShape));
        CHECK(expect([&] { conv(x22, kempty, ConvMode<2>::Same(), PaddingMode<2, int32_t>::Zeros()); }, ErrorKind::KernelShape));
        CHECK(expect([&] { conv(array1<int32_t>({1, 2, 3}), array1<int32_t>({1, 1, 1, 1, 1}), ConvMode<1>::Valid(), PaddingMode<1, int32_t>::Zeros()); }, ErrorKind::MismatchShape));
        Array<float, 2> kemptyf({0, 1}); auto x22f = array2<float>(2, 2, {1, 2, 3, 4});
        CHECK(expect([&] { conv_fft(x22f, kemptyf, ConvMode<2>::Same(), PaddingMode<2, float>::Zeros()); }, ErrorKind::DataShape));   // src/conv_fft/mod.rs:210-213
    });
    run_test("lowering: ConvMode::Same with an even dilated kernel pads [(Kd-1)/2+1, (Kd-1)/2]", [] {
        auto k = array1<int32_t>({1, 1, 1, 1});
        auto pr = lower(array1<int32_t>({1, 2, 3, 4, 5, 6}).view(), with_dilation(k, 1), ConvMode<1>::Same(), PaddingMode<1, int32_t>::Zeros());
        CHECK(pr.pad[0][0] == 2 && pr.pad[0][1] == 1 && pr.stride[0] == 1);
        CHECK(ndconv_good_fft_size(5030) == 5120);
    });
}

int main(int argc, char **argv)
{
    const bool host_only = argc > 1 && !std::strcmp(argv[1], "--host-only");
    host_tests();
    if (!host_only) device_tests();
    std::printf("%d tests, %d failed checks\n", ran, failures);
    if (!failures) std::printf(host_only ? "HOST_CPP_TESTS_PASSED\n" : "ALL_CPP_TESTS_PASSED\n");
    return failures ? 1 : 0;
}
