"""Multi-GPU sharding of one convolution: overlap-save slabs along axis 0.

Output row o of axis 0 reads padded rows o*s .. o*s+Kd0-1 only (src/conv/mod.rs:188-196; same crop identity in
src/conv_fft/mod.rs:282-289), so rank r of `world` owns a contiguous range of output rows and needs the padded rows
[pad_begin, pad_end) (its input rows plus a (Kd0-1)-row halo).  Axis-0 padding of the slab is materialised on the host
while cutting (each padded row is the source row the border map names, or a constant / zero row): that is exactly
"pad axis 0 first" of the reference's sequential padding (src/padding/mod.rs:119-153), so the slab then runs as an
ordinary problem with ConvMode::Explicit pads (0,0) on axis 0 and the unchanged borders on the other axes.
No collective is involved; ranks only agree on the plan (pure arithmetic).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import (BorderType, ConvMode, PATH_DIRECT, PATH_FFT, PaddingMode, Processor, _into_kwd, conv, conv_fft_with_processor,
               get_library, slab_plan)


def plan_rank(x_shape, dtype, kernel, conv_mode, padding_mode, path, rank, world, lib=None):
    """-> dict(out_begin, out_end, pad_begin, pad_end, rows: source row per padded row (>=0) or -1/-2/-3 codes, pads, strides)"""
    lib = lib or get_library()
    kwd = _into_kwd(kernel)
    sl = slab_plan(x_shape, dtype, kwd, conv_mode, padding_mode, path, world, rank, lib)
    pads, strides = conv_mode.unfold(kwd.kernel.shape, kwd.dilation, lib)
    n0 = x_shape[0]
    pf, pb = int(pads[0][0]), int(pads[0][1])
    b = padding_mode.lower(len(x_shape))[0]
    m = np.zeros(n0 + pf + pb, np.int32)
    lib.check(lib.c.ndconv_border_index_map(n0, pf, pb, b[0].kind, b[1].kind, m.ctypes.data))
    sl["rows"] = m[sl["pad_begin"]:sl["pad_end"]].copy()
    sl["pads"], sl["strides"], sl["border0"] = pads, strides, b
    return sl


def cut_slab(x, plan, dtype=None):
    """materialise the padded rows [pad_begin, pad_end) of axis 0 from the host array x"""
    rows = plan["rows"]
    x = np.asarray(x)
    slab = np.empty((len(rows),) + x.shape[1:], dtype or x.dtype)
    data = rows >= 0
    slab[data] = x[rows[data]]
    b = plan["border0"]
    for code, border in ((-1, b[0]), (-2, b[1])):
        slab[rows == code] = border.value if border.kind == BorderType.CONST else 0
    slab[rows == -3] = 0
    return slab


def slab_conv_mode(plan):
    """the slab's ConvMode: no axis-0 padding left, the caller's pads / strides on the other axes"""
    pads = [[0, 0]] + [[int(p[0]), int(p[1])] for p in plan["pads"][1:]]
    return ConvMode.Explicit(pads, [int(s) for s in plan["strides"]])


def slab_padding_mode(padding_mode, ndim):
    b = padding_mode.lower(ndim)
    return PaddingMode.Explicit([[BorderType.Zeros, BorderType.Zeros]] + [[bb[0], bb[1]] for bb in b[1:]])


def conv_rank(x, kernel, conv_mode, padding_mode, rank, world, path=PATH_FFT, processor: Processor | None = None, lib=None):
    """this rank's rows of conv / conv_fft(x, kernel, ...): returns (out_slab, out_begin, out_end)"""
    lib = lib or (processor.lib if processor is not None else get_library())
    x = np.asarray(x)
    plan = plan_rank(x.shape, x.dtype, kernel, conv_mode, padding_mode, path, rank, world, lib)
    if plan["out_end"] <= plan["out_begin"]:
        return None, plan["out_begin"], plan["out_end"]
    slab = cut_slab(x, plan)
    mode, pm = slab_conv_mode(plan), slab_padding_mode(padding_mode, x.ndim)
    if path == PATH_DIRECT:
        y = conv(slab, kernel, mode, pm, processor=processor, lib=lib)
    else:
        own = processor is None
        proc = processor or Processor(0, lib)
        y = conv_fft_with_processor(slab, kernel, mode, pm, proc)
        if own:
            proc.close()
    assert y.shape[0] == plan["out_end"] - plan["out_begin"]
    return y, plan["out_begin"], plan["out_end"]


# ------------------------------------------------------------------------------------------------------------
# device-resident input: the array is already row-partitioned across ranks; only halo rows move (NCCL send/recv
# over NVLink; gloo on CPU for the tests).  torch.distributed is plumbing here: rows in, rows out.
# ------------------------------------------------------------------------------------------------------------
def row_partition(n0, world):
    """balanced contiguous ownership of the input rows: rank r owns [bounds[r], bounds[r+1])"""
    base, rem = divmod(n0, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def exchange_halo_rows(x_local, n0, plans, rank, world, group=None):
    """x_local: torch tensor holding this rank's rows [bounds[rank], bounds[rank+1]) (any device).
    plans[r]["rows"]: the source row of every padded row rank r reads (from plan_rank; identical on every rank, pure
    arithmetic).  Returns the materialised slab of this rank (torch tensor, rows = pad_end - pad_begin) after ONE
    batched neighbour exchange: every rank sends exactly the rows a peer needs and it owns."""
    import torch
    import torch.distributed as dist
    bounds = row_partition(n0, world)
    lo, hi = bounds[rank], bounds[rank + 1]

    def needed_from(owner, reader):
        rows = plans[reader]["rows"]
        rows = np.unique(rows[rows >= 0])
        return rows[(rows >= bounds[owner]) & (rows < bounds[owner + 1])]

    ops, recv_bufs = [], {}
    for peer in range(world):
        if peer == rank:
            continue
        send_rows = needed_from(rank, peer)
        if len(send_rows):
            idx = torch.as_tensor(send_rows - lo, device=x_local.device)
            ops.append(dist.P2POp(dist.isend, x_local.index_select(0, idx).contiguous(), peer, group))
        recv_rows = needed_from(peer, rank)
        if len(recv_rows):
            buf = torch.empty((len(recv_rows),) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
            recv_bufs[peer] = (recv_rows, buf)
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    # assemble the slab: local rows, received rows, constant / zero rows
    rows = plans[rank]["rows"]
    slab = torch.zeros((len(rows),) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    mine = (rows >= lo) & (rows < hi)
    if mine.any():
        slab[torch.as_tensor(np.nonzero(mine)[0], device=x_local.device)] = x_local.index_select(0, torch.as_tensor(rows[mine] - lo, device=x_local.device))
    for peer, (recv_rows, buf) in recv_bufs.items():
        sel = (rows >= bounds[peer]) & (rows < bounds[peer + 1])
        pos = np.searchsorted(recv_rows, rows[sel])
        slab[torch.as_tensor(np.nonzero(sel)[0], device=x_local.device)] = buf.index_select(0, torch.as_tensor(pos, device=x_local.device))
    b = plans[rank]["border0"]
    for code, border in ((-1, b[0]), (-2, b[1])):
        sel = rows == code
        if sel.any() and border.kind == BorderType.CONST:
            slab[torch.as_tensor(np.nonzero(sel)[0], device=x_local.device)] = border.value
    return slab


def conv_fft_device_resident(x_local, n0, kernel, conv_mode, padding_mode, rank, world, processor: Processor, group=None):
    """conv_fft of an array that is already row-partitioned over the ranks' GPUs: halo exchange (NCCL), then the unchanged
    single-GPU pipeline on the slab.  Returns (y_local torch tensor, out_begin, out_end); enqueued on the current stream."""
    import torch
    from . import conv_device
    np_dtype = {torch.float32: np.float32, torch.float64: np.float64}[x_local.dtype]
    shape = (n0,) + tuple(x_local.shape[1:])
    plans = [plan_rank(shape, np_dtype, kernel, conv_mode, padding_mode, PATH_FFT, r, world, processor.lib) for r in range(world)]
    slab = exchange_halo_rows(x_local, n0, plans, rank, world, group)
    plan = plans[rank]
    mode, pm = slab_conv_mode(plan), slab_padding_mode(padding_mode, len(shape))
    strides = [int(np.prod(slab.shape[i + 1:])) for i in range(slab.dim())]
    processor.set_stream(torch.cuda.current_stream(slab.device).cuda_stream)
    oshape = conv_device("ndconv_conv_fft", processor, slab.data_ptr(), tuple(slab.shape), strides, np_dtype, kernel, mode, pm, None)
    y = torch.empty(oshape, dtype=x_local.dtype, device=slab.device)
    conv_device("ndconv_conv_fft", processor, slab.data_ptr(), tuple(slab.shape), strides, np_dtype, kernel, mode, pm, y.data_ptr())
    return y, plan["out_begin"], plan["out_end"]
