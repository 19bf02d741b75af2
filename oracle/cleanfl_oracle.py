"""TEST INFRASTRUCTURE — CPU oracle for the filament-cleanup candidates.  NOT product code.

Restates `CleanFl.create_coronal_hole_candidates` (`/root/reference/chips/cleanup.py:69-200`)
with the same numpy / OpenCV calls in the same order.  One third-party call is restated rather
than called: `skimage.transform.resize(data, (res, res), mode="reflect", anti_aliasing=True)`
(cleanup.py:113-130; scikit-image is not pinned by the reference's setup.py and is absent from
this image).  For an input that already has the requested shape the anti-aliasing sigma is
`max(0, (1 - 1) / 2) = 0` and the order-1 warp samples every pixel at its own centre, i.e. the
identity; for floating-point input no range rescaling happens.  The oracle therefore uses
`np.array(data, dtype=np.float64)` and rejects any other shape.

Pinned (tests/golden/make_golden.py) against the UNMODIFIED reference class run in the build
container with that same identity stub installed as `skimage.transform.resize`.
"""
from __future__ import annotations

from argparse import Namespace

import cv2
import numpy as np

DEFAULT_PARAMS = dict(clip_negative=0, clip_211_log=[0.8, 2.7], clip_193_log=[1.4, 3.0],
                      clip_171_log=[1.2, 3.9], bmmix_value=0.6357, bmhot_value=0.7,
                      bmcool_value=1.5102)


def identity_resize(data, shape, **_kw):
    data = np.asarray(data)
    if tuple(data.shape) != tuple(shape):
        raise NotImplementedError("resize restated only for unchanged shapes")
    return np.array(data, dtype=np.float64)


def create_coronal_hole_candidates(d171, d193, d211, pixel_radius, resolution, params=DEFAULT_PARAMS,
                                   disk_mask=None):
    """Returns Namespace(candidate, bmmix, bmhot, bmcool  [unmasked float32],
    candidate_bitmap, bmmix_m, bmhot_m, bmcool_m [masked, as stored on the object], disk_mask,
    thresholds=(mix, hot, cool), means=(171, 193, 211))."""
    p = Namespace(**params)
    if disk_mask is None:                                                   # cleanup.py:101-111
        disk_mask = np.zeros((resolution, resolution))
        cv2.circle(disk_mask, (int(resolution / 2), int(resolution / 2)), pixel_radius, 255, -1)
        disk_mask[disk_mask > 1] = 1.0

    def clip_limits(data, lims):
        return np.clip(data, lims[0], lims[1])

    def scale_limits(data, lims):
        return np.array(255 * (data - lims[0]) / (lims[1] - lims[0]), dtype=np.float32)

    shape = (resolution, resolution)
    a171, a193, a211 = (identity_resize(d, shape) for d in (d171, d193, d211))   # :113-130
    a171[a171 <= p.clip_negative] = 0                                       # :131-133
    a193[a193 <= p.clip_negative] = 0
    a211[a211 <= p.clip_negative] = 0
    bmcool = np.zeros(shape, dtype=np.float32)
    bmmix, bmhot = np.copy(bmcool), np.copy(bmcool)
    with np.errstate(divide="ignore", invalid="ignore"):
        l211, l193, l171 = np.log10(a211), np.log10(a193), np.log10(a171)  # :141-144
    s211 = scale_limits(clip_limits(l211, p.clip_211_log), p.clip_211_log)  # :146-152
    s193 = scale_limits(clip_limits(l193, p.clip_193_log), p.clip_193_log)
    s171 = scale_limits(clip_limits(l171, p.clip_171_log), p.clip_171_log)
    m171, m193, m211 = np.mean(a171), np.mean(a193), np.mean(a211)
    t_mix = m171 * p.bmmix_value / m211
    t_hot = p.bmhot_value * (m193 + m211)
    t_cool = m171 * p.bmcool_value / m193
    with np.errstate(divide="ignore", invalid="ignore"):
        bmmix[s171 / s211 >= t_mix] = 1                                     # :155-164
        bmhot[s211 + s193 < t_hot] = 1                                      # :165-172
        bmcool[s171 / s193 >= t_cool] = 1                                   # :173-181
    cand = bmcool * bmmix * bmhot                                           # :184
    return Namespace(candidate=cand, bmmix=bmmix, bmhot=bmhot, bmcool=bmcool,
                     candidate_bitmap=cand * disk_mask, bmmix_m=bmmix * disk_mask,
                     bmhot_m=bmhot * disk_mask, bmcool_m=bmcool * disk_mask, disk_mask=disk_mask,
                     thresholds=(t_mix, t_hot, t_cool), means=(m171, m193, m211),
                     scaled=(s171, s193, s211))
