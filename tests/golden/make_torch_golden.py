"""Generate tests/golden/torch_golden.json.

The reference pins its direct `conv` against libtorch (src/conv/tests.rs:36-547,
`assert_eq_tch`: rounded libtorch output must equal the i32 result exactly) and pins
`conv_fft` against `conv` on the cases of src/conv_fft/tests.rs:148-619.  The expectations
are computed by the reference's tests at run time, so they are not literal in the source.
This script recomputes them with torch CPU (the same external oracle the reference uses;
kernels are flipped where the reference test flips them, i.e. whenever `reverse` is on) and
writes inputs + expected outputs as a small committed fixture.

Run in the build container:  python tests/golden/make_torch_golden.py
Neither /root/reference nor torch is needed to *consume* the fixture.
"""
import json
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

CASES = [
    # (src, x, kernel, mode, padding, dilation, reverse)
    ("src/conv/tests.rs:45", [1, 2, 3, 4, 5], [1, 2, 1], "full", "zeros", 1, True),
    ("src/conv/tests.rs:69", [[1, 2], [3, 4]], [[1, 1], [1, 1]], "full", "zeros", 1, True),
    ("src/conv/tests.rs:93", [[[1, 2]], [[3, 4]]], [[[1, 1]], [[1, 1]]], "full", "zeros", 1, True),
    ("src/conv/tests.rs:123", [1, 2, 3, 4, 5], [1, 2, 1], "same", "zeros", 1, True),
    ("src/conv/tests.rs:146", [[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1, 0, -1], [2, 0, -2], [1, 0, -1]], "same", "zeros", 1, True),
    ("src/conv/tests.rs:179", [[[1, 2, 3]], [[4, 5, 6]]], [[[1, 1, 1]]], "same", "zeros", 1, True),
    ("src/conv/tests.rs:211", [1, 2, 3, 4, 5], [1, 2, 1], "valid", "zeros", 1, True),
    ("src/conv/tests.rs:234", [[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1, 1], [1, 1]], "valid", "zeros", 1, True),
    ("src/conv/tests.rs:257", [[[1, 2, 3], [4, 5, 6]], [[7, 8, 9], [10, 11, 12]]], [[[1, 1]], [[1, 1]]], "valid", "zeros", 1, True),
    ("src/conv/tests.rs:286", [1, 2, 3, 4, 5, 6], [1, 1, 1], ["custom", [1], [2]], "zeros", 1, True),
    ("src/conv/tests.rs:316", [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], [[1, 1], [1, 1]], ["custom", [1, 1], [2, 2]], "zeros", 1, True),
    ("src/conv/tests.rs:346", [1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 2, 1], ["custom", [2], [3]], "zeros", 1, True),
    ("src/conv/tests.rs:382", [1, 2, 3, 4, 5, 6], [1, 1, 2], ["custom", [4], [2]], "zeros", 2, False),
    ("src/conv/tests.rs:412", [[1, 1, 1], [1, 1, 1], [1, 1, 2]], [[2, 1, 1], [1, 1, 1]], "same", "zeros", 2, False),
    ("src/conv/tests.rs:439", [[[1, 2], [3, 4]], [[5, 6], [7, 8]]], [[[1, 1, 1]] * 3] * 2, ["custom", [2, 2, 2], [1, 2, 1]], "zeros", 2, False),
    ("src/conv/tests.rs:478", [1, 2, 3, 4, 5, 6], [1, 1, 2], ["custom", [4], [2]], "zeros", 2, True),
    ("src/conv/tests.rs:518", [1, 2, 3, 4, 5, 6], [1, 1, 2], ["custom", [4], [2]], "zeros", 2, False),
    # conv_fft/tests.rs cases (conv_fft must round to the same integers, tol 1e-5 f32 / 1e-9 f64, :15-16)
    ("src/conv_fft/tests.rs:152", [1, 2, 3, 4, 5, 6], [1, 1, 1, 1], "same", "zeros", 2, True),
    ("src/conv_fft/tests.rs:240", [1, 2, 3, 4, 5], [1, 2, 1], "full", "zeros", 1, True),
    ("src/conv_fft/tests.rs:261", [1, 2, 3, 4, 5, 6], [1, 1, 1], "valid", "zeros", 1, True),
    ("src/conv_fft/tests.rs:288", [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12]], [[1, 0], [3, 1]], "same", "replicate", 1, True),
    ("src/conv_fft/tests.rs:309", [[1, 2], [3, 4]], [[1, 0], [3, 1]], ["custom", [3, 3], [2, 2]], "replicate", 2, False),
    ("src/conv_fft/tests.rs:377", [[1, 2], [3, 4]], [[1, 1], [1, 1]], "full", "zeros", 1, True),
    ("src/conv_fft/tests.rs:398", [[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1, 1], [1, 1]], "valid", "zeros", 1, True),
    ("src/conv_fft/tests.rs:425", [[[1, 2], [3, 4]], [[5, 6], [7, 8]]], [[[1, 1, 1]] * 3] * 2, "same", "zeros", 1, True),
    ("src/conv_fft/tests.rs:510", [[[1, 2]], [[3, 4]]], [[[1, 1]], [[1, 1]]], "full", "zeros", 1, True),
    ("src/conv_fft/tests.rs:531", [[[1, 2, 3], [4, 5, 6]], [[7, 8, 9], [10, 11, 12]]], [[[1, 1]], [[1, 1]]], "valid", "zeros", 1, True),
    ("src/conv_fft/tests.rs:558", [[1, 2, 3], [4, 5, 6]], [[1, 1], [1, 1]], "same", "replicate", 1, True),
    ("src/conv_fft/tests.rs:579", [[1, 2, 3], [4, 5, 6]], [[1, 1], [1, 1]], "same", "zeros", 1, True),
    ("src/conv_fft/tests.rs:600", [[1, 2], [3, 4]], [[1, 1], [1, 1]], "full", ["const", 7], 1, True),
]

# SURVEY Appendix B values (derived at survey time); the generator asserts it reproduces them.
SURVEY_B = {
    "src/conv/tests.rs:45": [1, 4, 8, 12, 16, 14, 5],
    "src/conv/tests.rs:69": [1, 3, 2, 4, 10, 6, 3, 7, 4],
    "src/conv/tests.rs:146": [9, 6, -9, 20, 8, -20, 21, 6, -21],
    "src/conv/tests.rs:346": [1, 12, 24, 26],
    "src/conv/tests.rs:382": [2, 7, 14, 8, 5],
    "src/conv/tests.rs:412": [2, 1, 2, 5, 2, 6, 2, 1, 3],
    "src/conv/tests.rs:439": [1, 2, 5, 6, 1, 2, 5, 6],
    "src/conv/tests.rs:478": [1, 4, 10, 11, 10],
    "src/conv_fft/tests.rs:152": [6, 9, 12, 9, 12, 8],
    "src/conv_fft/tests.rs:288": [5, 9, 7, 11, 17, 21, 27, 31, 37, 41, 47, 51],
    "src/conv_fft/tests.rs:309": [5, 6, 10, 13, 14, 18, 15, 16, 20],
    "src/conv_fft/tests.rs:425": [10, 10, 10, 10, 36, 36, 36, 36],
    "src/conv_fft/tests.rs:600": [22, 17, 23, 18, 10, 20, 24, 21, 25],
}

TORCH_PAD_MODE = {"zeros": "constant", "replicate": "replicate", "reflect": "reflect", "circular": "circular"}


def unfold(mode, kshape, dil):
    """Independent Python statement of ConvMode::unfold (src/conv/mod.rs:28-66)."""
    kd = [k * d - d + 1 for k, d in zip(kshape, dil)]
    if mode == "full":
        return [[v - 1, v - 1] for v in kd], [1] * len(kd)
    if mode == "same":
        return [[(v - 1) // 2 + 1, (v - 1) // 2] if v % 2 == 0 else [(v - 1) // 2] * 2 for v in kd], [1] * len(kd)
    if mode == "valid":
        return [[0, 0] for _ in kd], [1] * len(kd)
    if mode[0] == "custom":
        return [[p, p] for p in mode[1]], list(mode[2])
    return [list(p) for p in mode[1]], list(mode[2])


def torch_expect(x, k, mode, padding, dilation, reverse):
    x = np.asarray(x, np.float64)
    k = np.asarray(k, np.float64)
    nd = x.ndim
    dil = [dilation] * nd
    pads, strides = unfold(mode, k.shape, dil)
    if reverse:  # the reference tests flip the kernel before handing it to libtorch
        k = k[tuple(slice(None, None, -1) for _ in range(nd))].copy()
    t = torch.from_numpy(x)[None, None]
    flat = []
    for p in reversed(pads):
        flat += p
    if isinstance(padding, str):
        pmode, val = TORCH_PAD_MODE[padding], 0.0
    else:
        pmode, val = "constant", float(padding[1])
    if any(flat):
        t = F.pad(t, flat, mode=pmode, value=val) if pmode == "constant" else F.pad(t, flat, mode=pmode)
    w = torch.from_numpy(k)[None, None]
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    y = fn(t, w, stride=strides, dilation=dil)[0, 0]
    return np.rint(y.numpy()).astype(np.int64)


def main():
    out = []
    for src, x, k, mode, padding, dilation, reverse in CASES:
        y = torch_expect(x, k, mode, padding, dilation, reverse)
        if src in SURVEY_B:
            assert y.ravel().tolist() == SURVEY_B[src], (src, y.ravel().tolist(), SURVEY_B[src])
        out.append({"src": src, "x": x, "kernel": k, "mode": mode, "padding": padding, "dilation": dilation,
                    "reverse": reverse, "expect_shape": list(y.shape), "expect": y.ravel().tolist()})
    # float KAT: src/conv_fft/tests.rs:209-237 (conv vs conv_fft, |diff| < 1e-6, circular) -- expectation from
    # torch circular pad + conv1d in float64
    x = [0.0, 0.1, 0.3, 0.4] * 4
    k = [0.1, 0.3, 0.6, 0.3, 0.1]
    t = F.pad(torch.tensor(x, dtype=torch.float64)[None, None], [2, 2], mode="circular")
    y = F.conv1d(t, torch.tensor(k[::-1], dtype=torch.float64)[None, None])[0, 0].numpy()
    float_case = {"src": "src/conv_fft/tests.rs:209", "x": x, "kernel": k, "mode": "same", "padding": "circular",
                  "dilation": 1, "reverse": True, "tol": 1e-6, "expect": y.tolist()}
    path = Path(__file__).with_name("torch_golden.json")
    path.write_text(json.dumps({"generator": "tests/golden/make_torch_golden.py", "torch": torch.__version__,
                                "int_cases": out, "float_cases": [float_case]}, indent=1))
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
