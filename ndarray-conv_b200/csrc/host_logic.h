// host_logic.h -- pure host-side logic of the hot path (no CUDA): shape checks in the
// reference's order, ConvMode unfolding, border index maps, tap lists, FFT length / tile planning,
// slab planning.  Everything here is reachable through the C ABI without a GPU.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/ndconv.h"
#include "common.h"

namespace ndc {

// Lowered, validated geometry of one ndconv_problem.
struct Geom {
    int ndim = 0, dtype = 0, es = 0;
    int64_t n[NDC_MAX_DIM], xstr[NDC_MAX_DIM];
    int64_t k[NDC_MAX_DIM], kstr[NDC_MAX_DIM], d[NDC_MAX_DIM], Kd[NDC_MAX_DIM];
    int64_t pf[NDC_MAX_DIM], pb[NDC_MAX_DIM], P[NDC_MAX_DIM];
    int64_t s[NDC_MAX_DIM], O[NDC_MAX_DIM];
    int bf[NDC_MAX_DIM], bb[NDC_MAX_DIM];
    int64_t out_total = 0, data_total = 0, kernel_total = 0;
    bool data_contiguous = false;   // standard layout
    bool reverse = true;
};

void set_error(const std::string &msg);
const char *get_error();

size_t dtype_size(int dtype);
bool dtype_is_float(int dtype);     // f32/f64/c32/c64
bool dtype_is_complex(int dtype);

// conv: src/conv/mod.rs:136-168 ; conv_fft: src/conv_fft/mod.rs:205-227.  Fills g (and the border
// maps, which is where a reference panic inside half_dim.rs is detected).
int check_problem(const ndconv_problem *pr, int path, Geom *g, std::vector<int32_t> maps[NDC_MAX_DIM]);

int unfold_mode(int mode, int ndim, const int64_t *kshape, const int64_t *dil, const int64_t *padding,
                const int64_t *strides, int64_t out_pad[][2], int64_t *out_stride);
int64_t good_size_cc(int64_t n);

// symbolic replay of src/padding/half_dim.rs on a 1-D axis
int build_border_map(int64_t n, int64_t pf, int64_t pb, int bf, int bb, std::vector<int32_t> &map);

// gen_offset_list, src/dilation/mod.rs:34-60: row-major over the (flipped) kernel, zero weights dropped.
struct Taps {
    int ntap = 0;
    std::vector<int32_t> off;     // [ntap][NDC_MAX_DIM] dilated offsets idx*d
    std::vector<int64_t> lin;     // [ntap] sum off*xstr (valid in the un-padded interior)
    std::vector<unsigned char> w; // [ntap] weights, es bytes each
};
void build_taps(const Geom &g, const void *kernel, Taps &t);

// ---- FFT planning ----
struct FftLen {
    int L = 0, npass = 0;
    int radix[NDC_MAX_PASS];
};
bool factor_radices(int L, FftLen *out, int max_radix = 32);          // false unless L is {2,3,5,7}-smooth
int64_t smooth_ge(int64_t n, bool even);
double fft_len_cost(int F, bool real_axis);

struct AxisTiling {
    int F = 0;       // transform length of one tile along this axis
    int V = 0;       // valid (alias-free) positions per tile = F - Kd + 1
    int ntiles = 0;
};
// choose (F, ntiles) for one axis: P padded extent, Kd dilated kernel extent, cap = largest
// transform length the shared-memory kernels take on this axis.
int plan_axis(int64_t P, int64_t Kd, int cap, bool real_axis, AxisTiling *out);

int slab_plan(const Geom &g, int n_slabs, int slab, ndconv_slab *out);

}  // namespace ndc
