"""ndconv_conv_fft_sharded_device (SURVEY 2.2 K10, 8e "Collective"): one conv_fft over an array that is already device-resident and
partitioned along axis 0, ghost rows filled peer-to-peer, the single-GPU pipeline in place on every shard.
  * not gpu: ndconv_shard_plan (host logic) -- the shards' output rows tile the output, ghost-row needs are what overlap-save says
  * gpu: equals the one-call convolution of the whole array (ghost rows poisoned with NaN beforehand); shards live on DISTINCT
    devices whenever more than one GPU is visible (the count is printed), otherwise on several handles of the one device."""
import numpy as np
import pytest

from test_parity_small import fft_tol, mode_from_spec, padding_from_spec

CASES = [
    # shape, kernel, dilation, mode, padding, shard rows
    ((3000, 1500), (31, 15), 1, "full", "reflect", [750, 750, 750, 750]),
    ((2500, 1300), (9, 5), 2, "same", ("custom", ["circular", "replicate"]), [1000, 700, 800]),          # Circular on the sharded axis: far-end halos
    ((2047, 1100), (7, 3), 1, ("custom", [3, 1], [2, 3]), ("const", 1.5), [600, 447, 1000]),               # strided outputs
    ((600, 40, 300), (5, 3, 3), 1, "valid", "zeros", [300, 300]),                                           # rank 3
]


def plans(pkg, lib, case):
    shape, ks, dil, mode, padding, rows = case
    k = np.zeros(ks, np.float32)
    kw = pkg.with_dilation(k, dil)
    return [pkg.shard_plan(shape, np.float32, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), rows, g, lib) for g in range(len(rows))]


@pytest.mark.parametrize("case", CASES, ids=[str(c[0]) for c in CASES])
def test_shard_plan_tiles_output_and_bounds_halos(ndc, case):
    pkg, lib = ndc
    shape, ks, dil, mode, padding, rows = case
    pl = plans(pkg, lib, case)
    kw = pkg.with_dilation(np.zeros(ks, np.float32), dil)
    pr, keep = pkg._global_problem(shape, np.float32, kw, mode_from_spec(pkg, mode), padding_from_spec(pkg, padding), lib)
    O = pkg.out_shape(pr, pkg.PATH_FFT, lib)
    assert pl[0]["out_begin"] == 0 and pl[-1]["out_end"] == O[0]
    first = 0
    for g, p in enumerate(pl):
        assert p["first_row"] == first
        first += rows[g]
        if g:
            assert p["out_begin"] == pl[g - 1]["out_end"]
        # a shard never needs more than the rows its outputs read: (out rows - 1) * stride + Kd0
        kd0 = (ks[0] - 1) * dil + 1
        s0 = int(pr.stride[0])
        assert p["halo_front"] + p["halo_back"] <= max(0, (p["out_end"] - p["out_begin"] - 1) * s0 + kd0)
    # the BASELINE workload on 8 GPUs: 62-row halos in total per interior boundary side
    c5 = [pkg.shard_plan((32768, 32768), np.float32, np.zeros((63, 63), np.float32), pkg.ConvMode.Full, pkg.PaddingMode.Reflect, [4096] * 8, g, lib) for g in range(8)]
    assert c5[0]["halo_front"] == 0 and c5[7]["halo_back"] == 0
    assert all(0 <= p["halo_front"] <= 62 and 0 <= p["halo_back"] <= 62 and p["halo_front"] + p["halo_back"] <= 70 for p in c5)


def test_shard_plan_argument_errors(ndc):
    pkg, lib = ndc
    k = np.zeros((3, 3), np.float32)
    with pytest.raises(pkg.NdConvError):
        pkg.shard_plan((100, 100), np.float32, k, pkg.ConvMode.Same, pkg.PaddingMode.Zeros, [60, 30], 0, lib)        # rows do not add up
    with pytest.raises(pkg.NdConvError):
        pkg.shard_plan((100,), np.float32, np.zeros(3, np.float32), pkg.ConvMode.Same, pkg.PaddingMode.Zeros, [50, 50], 0, lib)   # rank 1


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c[0]) for c in CASES])
def test_sharded_device_equals_single_call(pkg, cuda_lib, case):
    import torch
    shape, ks, dil, mode, padding, rows = case
    ndev = torch.cuda.device_count()
    print(f"[sharded_device] visible GPUs: {ndev} -> shards on {'distinct devices' if ndev > 1 else 'several handles of device 0'}")
    rng = np.random.default_rng(17)
    x = rng.random(shape, dtype=np.float32) - 0.3
    k = rng.random(ks, dtype=np.float32) - 0.5
    kw = pkg.with_dilation(k, dil)
    cm, pm = mode_from_spec(pkg, mode), padding_from_spec(pkg, padding)
    one = pkg.get_fft_processor(0, cuda_lib)
    ref = pkg.conv_fft_with_processor(x, kw, cm, pm, one)
    one.close()
    n = len(rows)
    devs = [g % ndev for g in range(n)]
    procs = [pkg.get_fft_processor(d, cuda_lib) for d in devs]
    pl = [pkg.shard_plan(shape, np.float32, kw, cm, pm, rows, g, cuda_lib) for g in range(n)]
    inner = int(np.prod(shape[1:]))
    bufs, outs, shards = [], [], []
    first = 0
    for g in range(n):
        hf, hb = pl[g]["halo_front"], pl[g]["halo_back"]
        buf = torch.full((hf + rows[g] + hb, inner), float("nan"), dtype=torch.float32, device=f"cuda:{devs[g]}")      # ghost rows poisoned
        buf[hf:hf + rows[g]] = torch.from_numpy(x[first:first + rows[g]].reshape(rows[g], inner)).to(buf.device)
        o = torch.full((pl[g]["out_end"] - pl[g]["out_begin"],) + tuple(ref.shape[1:]), float("nan"), dtype=torch.float32, device=buf.device)
        bufs.append(buf); outs.append(o)
        shards.append(dict(data=buf.data_ptr() + hf * inner * 4, rows=rows[g], halo_front=hf, halo_back=hb, out=o.data_ptr()))
        first += rows[g]
    for d in range(ndev):
        torch.cuda.synchronize(d)
    pkg.conv_fft_sharded_device(procs, shape, np.float32, kw, cm, pm, shards)
    for p in procs:
        p.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert got.shape == ref.shape
    tol = fft_tol(np.float32, 1024 * 2048, ref, float(np.max(np.abs(x)) * np.sum(np.abs(k))))
    assert np.isfinite(got).all()
    assert np.max(np.abs(got - ref)) <= tol, (np.max(np.abs(got - ref)), tol)
    # too few ghost rows is an argument error, not a wrong answer
    if any(p["halo_front"] + p["halo_back"] for p in pl):
        bad = [dict(s) for s in shards]
        for s in bad:
            s["halo_front"] = s["halo_back"] = 0
        with pytest.raises(pkg.NdConvError):
            pkg.conv_fft_sharded_device(procs, shape, np.float32, kw, cm, pm, bad)
    for p in procs:
        p.close()
