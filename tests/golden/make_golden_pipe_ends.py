"""Generate tests/golden/pipe_ends.npz from the reference's own lines (oracle/_ref, strict build): rawprepare,
temperature, highlights (clip + bypass), exposure, gamma and the export conversions on the cases of
tests/test_cpu_pipe_ends.py.  Run in the authoring container only:

    python tests/golden/make_golden_pipe_ends.py
"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import ansel_b200 as ab  # noqa: E402
import pipe_ends_util as pe  # noqa: E402
import test_cpu_pipe_ends as t  # noqa: E402
import util  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
save = {}
for name in t.RAWPREPARE_CASES:
    piece, src, g = t.rawprepare_case(name)
    save["rawprepare_" + name] = pe.ref_rawprepare(piece, src, **g)
for name in t.TEMPERATURE_CASES:
    save["temperature_" + name] = pe.ref_temperature(*t.temperature_case(name))
for name in t.HIGHLIGHTS_CASES:
    save["highlights_" + name] = pe.ref_highlights(*t.highlights_case(name))
for name in t.EXPOSURE_CASES:
    save["exposure_" + name] = pe.ref_exposure(*t.exposure_case(name))
for name in t.CHANNELMIXER_CASES:
    save["channelmixerrgb_" + name] = pe.ref_channelmixerrgb(*t.channelmixer_case(name))
for name in t.INITIALSCALE_CASES:
    save["initialscale_" + name] = pe.ref_clip_and_zoom(*t.initialscale_case(name))
for orientation in range(8):
    save[f"flip_{orientation}"] = pe.ref_flip(util.rgba_test_image(37, 23, 3), orientation)
for name in t.FINALSCALE_CASES:
    save["finalscale_" + name] = pe.ref_finalscale(*t.finalscale_case(name))
img = pe.awkward_rgba(141, 67, 12)
save["gamma"] = pe.ref_gamma(img)
for fmt in (ab.EXPORT_UINT8, ab.EXPORT_UINT8_SWAP, ab.EXPORT_UINT16):
    save[f"export_{fmt}"] = pe.ref_export(img, fmt)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipe_ends.npz")
np.savez_compressed(out, **save)
print("written", out, os.path.getsize(out) // 1024, "KiB")
