"""info['radar_debug'] of the reference (core.py:649-682, environment.py:842): a logging aid that inference.py:546 and
scripts/evaluate_hrl.py:323 pass to the episode logger.  The reference rebuilds 30 Python objects per env and step for it
(SURVEY quirk Q13); here it is computed on the host, on demand, from what the step already returns -- the env's true state,
its info flags and its observation row -- and only for small batches (HlynrVecEnv fills it when n <= 64).

Exact: every geometric field, the beam / range / elevation gates, the detection flags, the onboard quality, datalink quality
and fusion confidence.  Two fields need history the kernel does not keep and are approximated, with values the reference
itself uses:
  * onboard 'detection_reason' of a NOT detected sample under a sensor delay: the reason travels with the delayed sample in the
    reference; here it is derived from the CURRENT geometry (a few ticks newer: out_of_range / outside_beam, else poor_signal),
    'sensor_delay_initialization' while the delay buffer fills (without a delay the reason is exact);
  * ground 'quality' is the delayed sample's quality: reported as obs[23], i.e. 0.0 when the ground track is not in the
    observation (no current detection, or datalink quality <= 0.1).
"""
import math

import numpy as np

from . import abi


def forward_vector(q):
    """core.py:1143-1152 (float32)."""
    w, x, y, z = (np.float32(v) for v in q)
    f = np.array([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], dtype=np.float32)
    return f / (np.linalg.norm(f) + np.float32(1e-6))


def forward_from_euler(roll, pitch, yaw):
    """Third column of Rz(yaw) Ry(pitch) Rx(roll) = the forward vector of the quaternion these angles came from
    (core.py:1103-1121): used for the terminal step of a finished episode, whose quaternion is already reset."""
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    return np.array([cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr], dtype=np.float32)


def radar_debug(P, beam_width_deg, ipos, quat, mpos, steps, flags, obs_row, onboard_delay=None, forward=None):
    """One env.  P: HlynrParams; beam_width_deg: current curriculum value; ipos / quat / mpos: true float32 state after the
    step; steps: env steps of the running episode; flags: info flag word; obs_row: the 26-D observation of the step."""
    ip = np.asarray(ipos, np.float32); mp = np.asarray(mpos, np.float32)
    rel = mp - ip
    rng = float(np.linalg.norm(rel))
    fwd = forward_vector(quat) if forward is None else np.asarray(forward, np.float32)
    beam = float(np.arccos(np.clip(np.dot(fwd, rel / (np.float32(rng) + np.float32(1e-6))), -1, 1)))
    half = math.radians(float(beam_width_deg) / 2.0)
    o_det = bool(flags & abi.INFO_RADAR_DETECTED)
    g_det = bool(flags & abi.INFO_GROUND_DETECTED)
    delay = int(P.onboard_delay if onboard_delay is None else onboard_delay)
    if o_det:
        o_reason = "detected"
    elif delay > 0 and steps < delay:
        o_reason = "sensor_delay_initialization"
    elif rng > P.radar_range:      # (under a sensor delay: today's geometry stands in for the delayed sample's)
        o_reason = "out_of_range"
    elif beam > half:
        o_reason = "outside_beam"
    else:
        o_reason = "poor_signal"
    gpos = np.array(list(P.ground_pos), dtype=np.float64)
    g_range, g_elev = 0.0, 0.0
    if not P.ground_enabled:
        g_reason = "ground_radar_disabled"
    else:
        g = mp - gpos.astype(np.float32)
        gr = float(np.linalg.norm(g))
        if gr > P.g_max_range:
            g_reason = "out_of_range"      # no 'range' / 'elevation_deg' keys in the reference's dict: reported as 0.0
        else:
            el = float(np.arcsin(np.clip(g[2] / np.float32(gr), -1.0, 1.0))) if gr > 1e-6 else 0.0
            g_range, g_elev = gr, math.degrees(el)
            if gr > 1e-6 and el < P.g_min_el:
                g_reason = "below_horizon"
            elif gr > 1e-6 and el > P.g_max_el:
                g_reason = "above_coverage"
            elif mp[2] < 50.0:
                g_reason = "terrain_masking"
            elif g_det:
                g_reason = "detected"
            else:   # gates passed, no detection flag: a weak return -- or, while the ground delay buffer still fills, a current
                g_reason = "weak_return" if steps >= P.ground_delay else "unknown"   # detection whose delayed flag is False (core.py:673)
    g_quality = float(obs_row[23])
    return {
        "onboard": {
            "position": [float(v) for v in ip], "forward_vector": [float(v) for v in fwd],
            "beam_width_deg": float(beam_width_deg), "beam_angle_to_target_deg": math.degrees(beam),
            "half_beam_width_deg": math.degrees(half), "in_beam": bool(beam <= half),
            "range_to_target": rng, "max_range": float(P.radar_range), "detected": o_det,
            "detection_reason": o_reason, "quality": float(P.radar_quality) if o_det else 0.0},
        "ground": {
            "position": [float(v) for v in gpos] if P.ground_enabled else [0, 0, 0], "enabled": bool(P.ground_enabled),
            "max_range": float(P.g_max_range) if P.ground_enabled else 0.0,
            "min_elevation_deg": math.degrees(P.g_min_el) if P.ground_enabled else 0.0,
            "max_elevation_deg": math.degrees(P.g_max_el) if P.ground_enabled else 0.0,
            "range_to_target": g_range, "elevation_deg": g_elev, "detected": g_det, "detection_reason": g_reason,
            "quality": g_quality},
        "fusion": {"datalink_quality": float(obs_row[24]), "fusion_confidence": float(obs_row[25]),
                   "both_detected": bool(o_det and g_det), "any_detected": bool(o_det or g_det)},
    }
