"""GPU: one frame cut into row bands.  On the RCD block grid the banded C2 chain must give the
UNTILED frame bit for bit; with tiling.c-style overlap cuts (NLM in the chain) every band must equal the oracle
run on that band's rows, which is what the reference's own tiling produces."""
import ctypes as C
import socket

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def _c2_nodes():
    import ansel_b200 as ab
    from ansel_b200 import bands
    enc = util.srgb_encode_lut()
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=util.fit_unbounded_coeffs(enc))
    nodes = [bands.Node("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), channels_in=1), bands.Node("colorin", ab.colorin_data(conv_in)),
             bands.Node("colorout", ab.colorout_data(conv_out))]
    return nodes, (conv_in, conv_out)


def _untiled(nodes, mosaic):
    import torch
    from ansel_b200 import bands
    h, w = mosaic.shape
    ch = bands.BandedChain(nodes, w, h, 0, 1, device=torch.device("cuda", 0))
    out = ch(torch.from_numpy(mosaic).cuda(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("size", [(1500, 1000), (2002, 1601)])
def test_banded_c2_equals_untiled_bit_for_bit(built, world, size):
    import torch
    import ansel_b200 as ab
    from ansel_b200 import bands
    ab.init()
    w, h = size
    mosaic = util.frame_natural(w, h, 11)
    nodes, _keep = _c2_nodes()
    want = _untiled(nodes, mosaic)
    got = np.zeros_like(want)
    dev = torch.device("cuda", 0)
    for r in range(world):
        ch = bands.BandedChain(nodes, w, h, r, world, device=dev)
        assert (ch.grid, ch.halo) == (94, 9)
        mine = ch.run_band(torch.from_numpy(np.ascontiguousarray(ch.band_rows(mosaic))).to(dev), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got[ch.band.out_y0:ch.band.out_y1] = mine.cpu().numpy()
    assert same_bits(got, want).all()


def test_banded_nlm_chain_equals_per_band_oracle(built):
    """overlap cuts: demosaic -> profiled NLM -> colorin; each band == oracle chain on the band rows (the
    reference's tiled result for the same cuts)."""
    import torch
    import ansel_b200 as ab
    from ansel_b200 import bands
    ab.init()
    w, h = 320, 400
    mosaic = util.frame_natural(w, h, 13)
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    dn = ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=3)
    nodes = [bands.Node("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), channels_in=1), bands.Node("denoiseprofile", dn),
             bands.Node("colorin", ab.colorin_data(conv_in))]
    dev = torch.device("cuda", 0)
    for r in range(2):
        ch = bands.BandedChain(nodes, w, h, r, 2, device=dev)
        assert ch.grid == 1 and ch.halo == 10 + 1 + 3
        rows = np.ascontiguousarray(ch.band_rows(mosaic))
        mine = ch.run_band(torch.from_numpy(rows).to(dev), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = mine.cpu().numpy()
        # oracle chain on the band's rows = what the reference's tiling engine computes for this tile
        # (roi_in.y is even, so the CFA phase is the frame's)
        assert ch.band.in_y0 % 2 == 0
        rgb = util.oracle_rcd(rows, util.BAYER["RGGB"])
        undefined = (util.oracle_rcd_mask(rows, util.BAYER["RGGB"]) & 1) != 0
        f = util.oracle().orc_denoiseprofile_nlmeans
        f.restype = C.c_int
        den = np.zeros_like(rgb)
        assert f(util.fptr(rgb), util.fptr(den), w, rows.shape[0], C.byref(dn), C.c_float(1.0), 1, (C.c_float * 4)(2.0, 1.0, 1.5, 0.0),
                 (C.c_float * 4)(1.0, 1.0, 1.0, 1.0)) == 0
        want = util.oracle_convert(den, util.MATRIX_CAM_TO_REC2020, fp=util.FP_CONTRACT)
        lo, hi = ch.band.out_y0 - ch.band.in_y0, ch.band.out_y1 - ch.band.in_y0
        # pixels whose NLM window touches a reference-undefined demosaic pixel (frame ring of the band) are not compared
        k = 1 + 3
        taint = np.zeros(undefined.shape, bool)
        for dy in range(-k, k + 1):
            for dx in range(-k, k + 1):
                taint |= np.roll(np.roll(undefined, dy, 0), dx, 1)
        ok = same_bits(got[..., :3], want[lo:hi, :, :3]).all(axis=2) | taint[lo:hi]
        assert ok.all(), f"band {r}: {int((~ok).sum())} pixels differ"
        assert (~taint[lo:hi]).mean() > 0.8


def _nccl_worker(rank, world, port, w, h, q, p2p=False, p2p_dst=None):
    import torch
    import torch.distributed as dist
    import ansel_b200 as ab
    from ansel_b200 import bands
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ab.init()
        mosaic = util.frame_natural(w, h, 11)
        nodes, _keep = _c2_nodes()
        dev = torch.device("cuda", rank)
        ch = bands.BandedChain(nodes, w, h, rank, world, device=dev, p2p=p2p, p2p_dst=p2p_dst)
        band = torch.from_numpy(np.ascontiguousarray(ch.band_rows(mosaic))).to(dev)
        one = bands.BandedChain(nodes, w, h, 0, 1, device=dev)
        want = one(torch.from_numpy(mosaic).to(dev), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ok = True
        for _ in range(2):                      # twice: the second pass overwrites frames the peers have already read
            frame = ch(band, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            if frame is None:
                ok = ok and p2p_dst is not None and rank != p2p_dst
            else:
                ok = ok and bool(torch.equal(frame.view(torch.int32), want.view(torch.int32)))
            dist.barrier()
        ch.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True, "gather"])
def test_banded_c2_all_gather_over_nccl(built, p2p):
    """needs >= 2 GPUs (gpurun --gpus 2): every rank ends with the untiled frame, bit for bit -- gathered by NCCL
    (p2p=False) or by colorout's own stores into the peers' frames over NVLink (p2p=True)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, 3000, 2000, q, bool(p2p), 0 if p2p == "gather" else None)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
