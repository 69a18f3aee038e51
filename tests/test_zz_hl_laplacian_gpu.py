"""highlights, mode "guided laplacians" on the B200: ansel_b200/csrc/highlights_laplacian.cu through the C ABI against the oracle
(oracle/restate/highlights_laplacian_oracle.c, pinned on the reference's lines by tests/test_cpu_hl_laplacian.py)."""
import ctypes as C
import os
import numpy as np
import pytest
import util
import hl_laplacian_util as hu
from test_cpu_hl_laplacian import CASES, same_bits

pytestmark = pytest.mark.gpu
RGGB = util.BAYER["RGGB"]


@pytest.fixture(scope="module")
def built():
    import ansel_b200 as ab
    ab.init()
    return ab


@pytest.mark.parametrize("name,w,h,filters,kw", CASES, ids=[c[0] for c in CASES])
def test_bayer_bit_exact(built, name, w, h, filters, kw):
    m = hu.clipped_mosaic(w, h, len(name))
    want, norm = hu.oracle(m, filters, hu.clips_of(), **kw)
    got = hu.cuda(built, m, filters, norm=norm, **kw)
    assert same_bits(got, want).all()


def test_rgba_bit_exact(built):
    img = hu.clipped_rgba(500, 340, 5)
    kw = dict(iterations=3, noise_level=0.1, solid_color=0.1)
    want, norm = hu.oracle(img, 0, hu.clips_of(), **kw)
    assert same_bits(hu.cuda(built, img, 0, norm=norm, **kw), want).all()


def test_xtrans_bit_exact(built):
    from test_cpu_hl_laplacian import XCASES
    for name, w, h, kw in XCASES + [("wide", 1500, 700, dict(xtrans=hu.XTRANS, x=1, y=4, iterations=2))]:
        m = hu.clipped_mosaic(w, h, 3)
        want, norm = hu.oracle(m, 9, hu.clips_of(), **kw)
        assert same_bits(hu.cuda(built, m, 9, norm=norm, **kw), want).all(), name
    own = hu.cuda(built, m, 9, **kw)                                 # the library's own normalization on an X-Trans frame
    assert np.abs(own - want).max() < 1e-4 and same_bits(own, hu.cuda(built, m, 9, through_module=True, **kw)).all()


@pytest.mark.parametrize("name", list(hu.GOLDEN))
def test_committed_reference_output(built, name):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hl_laplacian.npz"))
    w, h, f, kw = hu.GOLDEN[name]
    img = hu.clipped_mosaic(w, h, len(name)) if f else hu.clipped_rgba(w, h, len(name))
    assert same_bits(hu.cuda(built, img, f, norm=g[name + "_norm"], **kw), g[name]).all()


def test_processed_maximum_and_clip_reach_the_clips(built):
    m = hu.clipped_mosaic(300, 220, 8, level=1.7) * np.float32(1.0)
    pmax, clip = (2.0, 1.0, 1.5), 0.9
    clips = hu.clips_of(clip, pmax)
    want, norm = hu.oracle(m, RGGB, clips, iterations=2)
    assert same_bits(hu.cuda(built, m, RGGB, norm=norm, clip=clip, pmax=pmax, iterations=2), want).all()


def test_own_normalization_is_the_mean_in_double(built):
    """without a vector the library sums in double in a fixed order: the frame is the oracle's under the exactly rounded means (or under a
    neighbouring float of one of them), twice the same; through the module's entry point the same frame comes out past the bypass"""
    m = hu.clipped_mosaic(1200, 900, 21)
    exact = np.array([m[0::2, 0::2].sum(dtype=np.float64), m[0::2, 1::2].sum(dtype=np.float64) + m[1::2, 0::2].sum(dtype=np.float64),
                      m[1::2, 1::2].sum(dtype=np.float64), m.size]) / np.float64(np.float32(m.size))
    norm = exact.astype(np.float32)
    got = hu.cuda(built, m, RGGB)
    want, _ = hu.oracle(m, RGGB, hu.clips_of(), norm=norm)
    assert same_bits(got, hu.cuda(built, m, RGGB)).all()
    assert same_bits(got, want).mean() > 0.999 and np.abs(got - want).max() < 1e-5
    assert same_bits(hu.cuda(built, m, RGGB, through_module=True), got).all()


def test_bypass_and_refusals(built):
    import torch
    ab = built
    calm = (util.frame_natural(320, 240, 3) * np.float32(0.5)).astype(np.float32)
    assert (hu.cuda(ab, calm, RGGB, through_module=True) == calm).all()          # fewer than 25 clipped samples: copied through
    d = ab.HighlightsData()
    d.mode, d.clip = ab.HIGHLIGHTS_HARMONIC, 1.0
    m = hu.clipped_mosaic(240, 180, 2)
    piece = ab.make_piece(240, 180, filters=RGGB, data=d, devid=0)
    t = torch.from_numpy(m).cuda()
    o = torch.empty_like(t)
    assert ab.lib().b200_highlights_process_dev(piece, t.data_ptr(), o.data_ptr(), torch.cuda.current_stream().cuda_stream) == ab.B200_ERR_UNSUPPORTED
    with pytest.raises(ab.B200Error) as e:
        hu.cuda(ab, hu.clipped_mosaic(6, 40, 2), RGGB)
    assert e.value.code == ab.B200_ERR_UNSUPPORTED
    torch.cuda.synchronize()


def test_45mp_properties(built):
    """the full frame: deterministic, finite, and the input wherever the feathered mask is zero (more than three pixels from any clipped sample:
    one for the interpolated flags, two for the box mean); a full-width strip against the oracle"""
    w, h = util.SIZE_45MP
    m = hu.clipped_mosaic(w, h, 45, blobs=9)
    got = hu.cuda(built, m, RGGB, iterations=2)
    assert np.isfinite(got).all() and same_bits(got, hu.cuda(built, m, RGGB, iterations=2)).all()
    clipped = m > np.float32(0.995)
    assert clipped.mean() > 0.01
    import scipy.ndimage as ndi
    near = ndi.binary_dilation(clipped, structure=np.ones((3, 3), bool), iterations=4)   # interpolation 1 px + box mean 2 px, and one spare
    # the column pass of the box mean is a running float sum: below a clipped area it does not return to zero exactly, so the reference
    # blends a residue of about 1e-8 of the reconstruction into the rest of those columns; elsewhere the input comes through untouched
    far = ~near
    assert np.abs(got[far] - m[far]).max() < 1e-5 and same_bits(got[far], m[far]).mean() > 0.5
    assert (got[clipped] != m[clipped]).mean() > 0.5
    top, norm = hu.oracle(np.ascontiguousarray(m[:400]), RGGB, hu.clips_of(), iterations=2)
    strip = hu.cuda(built, np.ascontiguousarray(m[:400]), RGGB, norm=norm, iterations=2)      # a full-width strip against the oracle
    assert same_bits(strip, top).all()
