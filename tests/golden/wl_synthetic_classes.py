"""The reference's synthetic Winston-Lutz test classes (tests_basic/test_winstonlutz.py:1244-1520): BB offsets, image axes and the
expected statistics, verbatim; frames from the restated generator (oracle/synth.py)."""
import numpy as np

AXES8 = ((0, 0, 0), (90, 0, 0), (180, 0, 0), (270, 0, 0), (0, 0, 45), (0, 0, 90), (0, 0, 270), (0, 0, 315))
AXES4 = ((0, 0, 0), (90, 0, 0), (180, 0, 0), (270, 0, 0))

# name: (left, up, in, axes, {expected attribute: value})   -- verbatim from the reference's test classes
CLASSES = {
    "Synthetic1mmLeftNoCouch": (1, 0, 0, AXES4, dict(bb_max=1, bb_mean=0.5, bb_median=0.5, epid_max=1)),
    "Synthetic1mmLeft": (1, 0, 0, AXES8, dict(bb_max=1, bb_mean=0.67, bb_median=1, epid_max=1.02, couch_iso=2)),
    "Synthetic1mmRight": (-1, 0, 0, AXES8, dict(bb_max=1.0, bb_mean=0.75, bb_median=1, epid_max=1.02, couch_iso=2)),
    "Synthetic1mmUp": (0, 1, 0, AXES8, dict(bb_max=1.0, bb_mean=0.25, bb_median=0, epid_max=1)),
    "Synthetic1mmDown": (0, -1, 0, AXES8, dict(bb_max=1.0, bb_mean=0.25, bb_median=0, epid_max=1, couch_iso=0)),
    "Synthetic1mmIn": (0, 0, 1, AXES8, dict(bb_max=1.0, bb_mean=1.0, bb_median=1, epid_max=1.02, couch_iso=2.0)),
    "Synthetic1mmOut": (0, 0, -1, AXES8, dict(bb_max=1.0, bb_mean=1.0, bb_median=1, epid_max=1, couch_iso=2.0)),
    "Synthetic1mmIn1mmLeft": (1, 0, 1, AXES8, dict(bb_max=1.41, bb_mean=1.3, bb_median=1.4, epid_max=1.42, couch_iso=2.8)),
    "Synthetic1mmOut1mmRight": (-1, 0, -1, AXES8, dict(bb_max=1.41, bb_mean=1.3, bb_median=1.4, epid_max=1.42, couch_iso=2.8)),
    "Synthetic2mmUp1mmLeft": (1, 2, 0, AXES8, dict(bb_max=2.0, bb_mean=1.25, bb_median=1.0, epid_max=2.02, couch_iso=2.0)),
    "Synthetic2mmRight1mmDown": (-2, -1, 0, AXES8, dict(bb_max=2.0, bb_mean=1.75, bb_median=2.0, epid_max=2.02, couch_iso=4.0)),
    "Synthetic1mmOut1SidedCouch": (0, 0, -1, AXES8[:6], dict(bb_max=1.0, bb_mean=1.0, bb_median=1, epid_max=1.02, couch_iso=1.42)),
}


def _set(left, up, inn, axes):
    from oracle import synth

    frames = []
    for g, c, p in axes:
        fr = synth.as1200(1000.0)
        frames.append(synth.winstonlutz_frame(fr, field_size_mm=(20, 20), bb_size_mm=5.0, offset_mm_left=left, offset_mm_up=up,
                                              offset_mm_in=inn, gantry=g, coll=c, couch=p, bb_alpha=-0.8, blur_mm=1.5, field="perfect"))
    return np.stack(frames), 1 / 0.336
