"""Host logic of the multi-GPU band sharding: the planner, and the N>1 path end to end over
gloo with world sizes 2 and 3.  No GPU here, and the product has no CPU fallback: the per-module work is a
numpy stand-in injected through BandedChain(process=...) whose result depends on the band origin, the halo
and the 94-row block grid exactly the way RCD's does, so a wrong cut shows up as a wrong frame."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import util  # noqa: F401  (path set-up)


def _plan(*a, **k):
    from ansel_b200 import bands
    return bands.plan(*a, **k)


@pytest.mark.parametrize("height", [5504, 4000, 8736, 200, 95, 19])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 8])
def test_grid_plan_covers_frame_on_the_block_grid(built, height, n):
    b = _plan(height, n, 94, 9, 2)
    assert b[0].out_y0 == 0 and b[-1].out_y1 == height
    for i, q in enumerate(b):
        assert q.out_y0 <= q.out_y1
        if i:
            assert q.out_y0 == b[i - 1].out_y1
        if q.out_y1 == q.out_y0:
            continue
        assert q.in_y0 == max(0, q.out_y0 - 9) and q.in_y1 == min(height, q.out_y1 + 9)
        assert q.in_y0 % 94 == 0                              # band tile grid == frame tile grid
        assert q.in_y1 == height or (q.in_y1 - 18) % 94 == 0  # ends with a whole tile
    if height >= 94 * n * 4:
        hs = [q.out_y1 - q.out_y0 for q in b]
        assert max(hs) - min(hs) <= 2 * 94 + 9                # each cut rounds to its nearest block row


@pytest.mark.parametrize("n", [2, 4, 8])
def test_overlap_plan_is_tiling_c_shaped(built, n):
    b = _plan(5504, n, 1, 8, 2)
    for i, q in enumerate(b):
        assert q.out_y0 % 2 == 0 and q.in_y0 % 2 == 0
        assert q.in_y0 <= max(0, q.out_y0 - 8) and q.in_y1 >= min(5504, q.out_y1 + 8)
        assert q.in_y0 >= max(0, q.out_y0 - 8 - 2) and q.in_y1 <= min(5504, q.out_y1 + 8 + 2)


def test_plan_rejects_bad_arguments(built):
    import ansel_b200 as ab
    from ansel_b200 import bands
    for args in ((0, 2, 1, 0, 1), (100, 0, 1, 0, 1), (100, 2, 94, 9, 4), (100, 65, 1, 0, 1)):
        with pytest.raises(ab.B200Error):
            bands.plan(*args)


def test_chain_cuts_follow_the_modules(built):
    import ansel_b200 as ab
    from ansel_b200 import bands
    dem = bands.Node("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), channels_in=1)
    cin = bands.Node("colorin", ab.colorin_data(ab.make_conversion(util.MATRIX_CAM_TO_REC2020)))
    assert bands.chain_cuts([dem, cin], 8256, 5504) == (94, 9, 2)
    nlm = bands.Node("denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=7))
    g, halo, align = bands.chain_cuts([dem, nlm, cin], 8256, 5504)
    assert g == 1 and halo == 10 + 1 + 7 and align == 2          # demosaic.c:1972-1982 + denoiseprofile.c:803-811
    with pytest.raises(NotImplementedError):
        bands.chain_cuts([bands.Node("bilat", ab.bilat_data())], 800, 600)
    # whole-frame green sums: refused; colour smoothing: one more halo row per pass, tiling.c-style cuts (no block grid)
    for mode in (ab.GREEN_EQ_FULL, ab.GREEN_EQ_BOTH):
        dd = ab.demosaic_data(ab.DEMOSAIC_RCD)
        dd.green_eq = mode
        with pytest.raises(NotImplementedError):
            bands.chain_cuts([bands.Node("demosaic", dd, channels_in=1), cin], 8256, 5504)
    dd = ab.demosaic_data(ab.DEMOSAIC_RCD)
    dd.green_eq, dd.color_smoothing = ab.GREEN_EQ_LOCAL, 3
    g, halo, align = bands.chain_cuts([bands.Node("demosaic", dd, channels_in=1), cin], 8256, 5504)
    assert g == 1 and halo == 10 + 3
    # the dual demosaic (RCD | 2048) blends pointwise: still on RCD's block grid; AMaZE is not
    g2, h2, a2 = C.c_int(), C.c_int(), C.c_int()
    for method, want in ((ab.DEMOSAIC_RCD | 2048, 94), (ab.DEMOSAIC_AMAZE, 1)):
        piece = ab.make_piece(800, 600, filters=0x94949494, channels=1, data=ab.demosaic_data(method))
        ab.lib().b200_demosaic_band_grid(C.byref(piece), C.byref(g2), C.byref(h2), C.byref(a2))
        assert g2.value == want
    with pytest.raises(NotImplementedError):
        bands.chain_cuts([bands.Node("denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_WAVELETS))], 800, 600)


# ---- the N>1 path over gloo -------------------------------------------------------------------------
def standin_demosaic(mosaic, roi_y):
    """depends on: the band origin (ch1), the +-9 row neighbourhood clipped at the band edge (ch0), the block
    index a band-local 94-row grid gives (ch2) -- like rcd.c."""
    h, w = mosaic.shape
    out = np.zeros((h, w, 4), np.float32)
    acc = np.zeros((h, w), np.float64)
    for d in range(-9, 10):
        lo, hi = max(0, -d), min(h, h - d)
        acc[lo:hi] += mosaic[lo + d:hi + d]
    out[..., 0] = acc
    out[..., 1] = (np.arange(h) + roi_y)[:, None]
    out[..., 2] = (np.arange(h) // 94 + roi_y // 94)[:, None]
    return out


def standin_pointwise(rgba):
    return rgba * np.float32(0.5) + np.float32(1.0)


def _process(op, piece, src, dst, stream):
    a = src.numpy()
    if op == "demosaic":
        r = standin_demosaic(a, piece.roi_in.y)
    else:
        r = standin_pointwise(a[: piece.roi_in.height])
    dst.numpy()[: r.shape[0]] = r


def _worker(rank, world, port, mode, w, h, q):
    import torch
    import torch.distributed as dist
    import ansel_b200 as ab
    from ansel_b200 import bands
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mosaic = np.random.default_rng(7).random((h, w), dtype=np.float32)
        nodes = [bands.Node("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), channels_in=1),
                 bands.Node("colorin", ab.colorin_data(ab.make_conversion(util.MATRIX_CAM_TO_REC2020)))]
        tiles = mode.endswith("-tiles")   # a module with an overlap in the chain: tiling.c-style cuts, equal bands, gathered in place
        if tiles:
            nodes.insert(1, bands.Node("denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=7)))
            mode = mode[:-6]
        ch = bands.BandedChain(nodes, w, h, rank, world, process=_process)
        band_in = torch.from_numpy(np.ascontiguousarray(ch.band_rows(mosaic)))
        frame = ch(band_in, mode=mode)
        want = standin_pointwise(standin_demosaic(mosaic, 0))
        if tiles:
            want = standin_pointwise(want)
        if frame is None:
            ok = mode == "gather" and rank != 0
        else:
            got = frame.numpy()
            # ch0 of the untiled stand-in clips its window at the FRAME edge only; ch1, ch2 are global row / block ids
            # (ch2 follows the band-local block grid, which tiling.c-style cuts do not align with the frame's)
            ok = bool((got[..., :2] == want[..., :2]).all()) if tiles else bool((got == want).all())
            ok = ok and ch.equal_bands == tiles and "one ncclAllGather" in ch.collective
        q.put((rank, ok, [(b.out_y0, b.out_y1) for b in ch.bands]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,mode", [(2, "allgather"), (3, "allgather"), (2, "gather"), (2, "allgather-tiles"), (4, "allgather-tiles")])
def test_banded_chain_over_gloo(built, world, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, 96, 704 if mode.endswith("-tiles") else 700, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


# ---- whole-frame modules inside a banded chain (C4): SegmentedChain over gloo ----------------------------------
def standin_whole(rgba):
    """needs the whole frame: every pixel minus the frame mean of its lane (like a pyramid's coarsest level reaching everywhere)"""
    return (rgba - rgba.reshape(-1, 4).mean(axis=0, dtype=np.float64).astype(np.float32)).astype(np.float32)


def _process_seg(op, piece, src, dst, stream):
    a = src.numpy()
    if op == "demosaic":
        r = standin_demosaic(a, piece.roi_in.y)[..., :4]
        r[..., 2] = 0.0                                   # the band-local block index is not a frame property
    elif op == "bilat":
        r = standin_whole(a[: piece.roi_in.height])
    else:
        r = standin_pointwise(a[: piece.roi_in.height])
    dst.numpy()[: r.shape[0]] = r


def _seg_worker(rank, world, port, w, h, q):
    import torch
    import torch.distributed as dist
    import ansel_b200 as ab
    from ansel_b200 import bands
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mosaic = np.random.default_rng(9).random((h, w), dtype=np.float32)
        cin = bands.Node("colorin", ab.colorin_data(ab.make_conversion(util.MATRIX_CAM_TO_REC2020)))
        nodes = [bands.Node("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), channels_in=1),
                 bands.Node("denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=7)), cin,
                 bands.Node("bilat", ab.bilat_data()), cin, cin]
        ch = bands.SegmentedChain(nodes, w, h, rank, world, process=_process_seg)
        frame = ch(torch.from_numpy(np.ascontiguousarray(ch.band_rows(mosaic)))).numpy()
        d = standin_demosaic(mosaic, 0)
        d[..., 2] = 0.0
        want = standin_pointwise(standin_pointwise(standin_whole(standin_pointwise(standin_pointwise(d)))))
        ok = bool((frame == want).all()) and ch.collectives == 2 and [p[0] for p in ch.plan] == ["bands", "whole", "bands"]
        q.put((rank, ok, ch.plan))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_segmented_chain_with_a_whole_frame_module_over_gloo(built, world):
    """banded segment -> all-gather -> whole-frame module on every rank -> banded segment -> all-gather == the untiled chain"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seg_worker, args=(r, world, port, 96, 704, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


def test_band_rows_widened_until_non_local_means_chunks_fit_the_pipelined_kernel():
    """bands.nlm_slice_height restates compute_slice_height() (the oracle's copy is pinned to the reference's); widen_for_nlm only ever adds
    input rows inside the frame, keeps the kept rows and the row alignment, and ends with chunks of 64 rows or less"""
    import ctypes as C
    import util
    from ansel_b200 import bands
    o = util.oracle()
    for hgt in list(range(16, 400)) + [2752, 2770, 2788, 4418, 5504, 8736]:
        assert bands.nlm_slice_height(hgt) == o.orc_nlm_slice_height(hgt), hgt
    for height, n, halo in ((5504, 2, 18), (5504, 4, 18), (5504, 8, 18), (8736, 2, 50), (8736, 4, 50), (3000, 3, 18), (999 * 2, 5, 20)):
        before = bands.plan(height, n, 1, halo, 2)
        after = bands.widen_for_nlm(before, height, 2)
        for b, a in zip(before, after):
            assert (a.out_y0, a.out_y1) == (b.out_y0, b.out_y1)
            assert 0 <= a.in_y0 <= b.in_y0 and b.in_y1 <= a.in_y1 <= height and a.in_y0 % 2 == 0
            assert bands.nlm_slice_height(a.in_y1 - a.in_y0) <= 64
