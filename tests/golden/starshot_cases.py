"""Seeded synthetic Starshot cases shared by the golden generator, the oracle tests and the GPU parity tests
(docs/source/starshot_docs.rst:232-330 recipe through oracle/synth.py)."""
from __future__ import annotations

import numpy as np

from oracle import synth

CASES = ["perfect6", "offset6", "noisy6", "spokes4", "spokes8", "inverted", "as1200", "user_invert", "start_point", "wide_wobble",
         "no_fwhm", "faint"]


def _offsets(seed, n, amp):
    rng = np.random.default_rng(seed)
    return [tuple(rng.uniform(-amp, amp, 2)) for _ in range(n)]


def case_frame(name):
    """-> (frame uint16, pixel_spacing_mm, sid, analyze_kwargs)"""
    ps, sid = 0.390625, 1000.0
    if name == "perfect6":
        return synth.starshot_frame(synth.epid1024()), ps, sid, {}
    if name == "offset6":
        return synth.starshot_frame(synth.epid1024(), offsets_mm=_offsets(11, 6, 0.5), noise_sigma=0.002, seed=11), ps, sid, {}
    if name == "noisy6":
        return synth.starshot_frame(synth.epid1024(), offsets_mm=_offsets(12, 6, 0.3), noise_sigma=0.01, seed=12), ps, sid, {}
    if name == "spokes4":
        return synth.starshot_frame(synth.epid1024(), spokes=4, offsets_mm=_offsets(13, 4, 0.4), noise_sigma=0.002, seed=13), ps, sid, {}
    if name == "spokes8":
        return synth.starshot_frame(synth.epid1024(), spokes=8, offsets_mm=_offsets(14, 8, 0.4), noise_sigma=0.002, seed=14), ps, sid, {}
    if name == "inverted":
        fr = synth.epid1024()
        synth.starshot_frame(fr, offsets_mm=_offsets(15, 6, 0.5), noise_sigma=0.002, seed=15)
        return fr.inverted(), ps, sid, {}
    if name == "as1200":
        fr = synth.as1200(1000.0)
        return synth.starshot_frame(fr, offsets_mm=_offsets(16, 6, 0.5), noise_sigma=0.002, seed=16), fr.pixel_size, 1000.0, {}
    if name == "user_invert":
        # the histogram check leaves this frame alone; invert=True flips it once more, the analysis then fails or wanders
        fr = synth.epid1024()
        synth.starshot_frame(fr, offsets_mm=_offsets(17, 6, 0.3), noise_sigma=0.002, seed=17)
        return fr.inverted(), ps, sid, {"invert": False, "min_peak_height": 0.3}
    if name == "start_point":
        return synth.starshot_frame(synth.epid1024(), offsets_mm=_offsets(18, 6, 0.5), noise_sigma=0.002, seed=18), ps, sid, {
            "start_point": (500, 520), "radius": 0.7}
    if name == "wide_wobble":
        return synth.starshot_frame(synth.epid1024(), offsets_mm=_offsets(19, 6, 1.5), noise_sigma=0.002, seed=19), ps, sid, {
            "tolerance": 0.5}
    if name == "no_fwhm":
        return synth.starshot_frame(synth.epid1024(), offsets_mm=_offsets(20, 6, 0.5), noise_sigma=0.002, seed=20), ps, sid, {
            "fwhm": False}
    if name == "faint":
        return synth.starshot_frame(synth.epid1024(), offsets_mm=_offsets(21, 6, 0.5), alpha=0.1, noise_sigma=0.004, seed=21), ps, sid, {
            "min_peak_height": 0.5}
    raise KeyError(name)
