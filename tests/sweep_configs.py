"""Deterministic pseudo-random env dicts that mix the feature switches in combinations no shipped YAML uses
(every physics v2.0 sub-switch on its own, odd delays, no ground radar, spherical spawns, precision mode + fuze +
volley + observation modes on top of domain randomization ...).  Shared by the oracle-vs-reference sweep (CPU, build
container) and the CUDA-vs-oracle sweep (GPU)."""
import copy

import numpy as np

from hlynr_intercept_b200 import config

N_SWEEP = 30


def sweep_config(k):
    rng = np.random.default_rng(1000 + k)
    pick = lambda *xs: xs[int(rng.integers(len(xs)))]  # noqa: E731
    flip = lambda p=0.5: bool(rng.random() < p)  # noqa: E731
    e = copy.deepcopy(config.baseline_config(pick("cfg1", "cfg2", "cfg4", "cfg4", "cfg3")))
    e["max_steps"] = int(pick(400, 900, 2000))
    e["missile_evasion"] = flip()
    e["wind"] = dict(velocity=[float(rng.uniform(-12, 12)), float(rng.uniform(-8, 8)), float(rng.uniform(-2, 2))],
                     variability=float(pick(0.0, 0.05, 0.3)))
    pe = dict(enabled=flip(0.8))
    if pe["enabled"]:
        pe["atmospheric_model"] = dict(enabled=flip(0.7))
        pe["mach_effects"] = dict(enabled=flip(0.7), subsonic_mach=float(pick(0.8, 0.3)), supersonic_mach=float(pick(1.2, 0.6)),
                                  transonic_peak_multiplier=float(pick(3.0, 2.0)), supersonic_multiplier=2.5)
        pe["sensor_delays"] = dict(enabled=flip(0.7), radar_delay_ms=float(pick(10.0, 30.0, 70.0)))
        pe["thrust_dynamics"] = dict(enabled=flip(0.7), response_time_constant=float(pick(0.05, 0.1, 0.3)))
        pe["enhanced_wind"] = dict(enabled=flip(0.7), boundary_layer_height=float(pick(300.0, 1000.0)),
                                   turbulence_intensity=float(pick(0.0, 0.1, 0.25)), max_gust_speed=float(pick(5.0, 12.0)))
        pe["domain_randomization"] = dict(enabled=flip(0.4), drag_coefficient_variation=0.2, air_density_variation=0.1,
                                          sensor_delay_variation=0.5, thrust_response_variation=0.3, wind_variation=0.3)
    e["physics_enhancements"] = pe
    g = dict(e.get("ground_radar", {}))
    g["enabled"] = flip(0.8)
    g["ground_sensor_delay_ms"] = float(pick(0.0, 20.0, 50.0, 90.0))
    g["datalink_packet_loss"] = float(pick(0.0, 0.05, 0.3))
    g["max_range"] = float(pick(20000.0, 2500.0))
    e["ground_radar"] = g
    e["radar"] = dict(radar_beam_width=float(pick(120.0, 60.0, 30.0)), radar_quality=float(pick(1.0, 0.8)),
                      radar_range=float(pick(5000.0, 2000.0)))
    mode = pick("world_frame", "world_frame", "body_frame", "los_frame")
    if mode != "world_frame":
        e["observation_mode"] = mode
    if flip(0.35):
        e["precision_mode"] = True
    if flip(0.35):
        e["proximity_fuze_enabled"] = True
        e["proximity_kill_radius"] = float(pick(10.0, 25.0))
    if flip(0.3):
        e["volley_mode"] = True
        e["volley_size"] = int(pick(2, 3, 6))
    if flip(0.4):   # spherical missile spawn around the target + launch toward the missile (configs/hrl/terminal_360_*.yaml)
        e["missile_spawn"] = dict(position_mode="spherical", radius_min=float(pick(300.0, 900.0)), radius_max=float(pick(1200.0, 2500.0)),
                                  azimuth_range=[0.0, 360.0], elevation_range=[float(pick(5.0, 20.0)), float(pick(45.0, 70.0))],
                                  velocity_mode="toward_target", speed_min=80.0, speed_max=float(pick(120.0, 220.0)),
                                  position=[[800, 800, 800], [1500, 1500, 1500]], velocity=[[-60, -60, -30], [-100, -100, -50]])
        isp = dict(e["interceptor_spawn"])
        isp.update(velocity_mode="toward_missile", speed_min=40.0, speed_max=90.0)
        e["interceptor_spawn"] = isp
    return e


def sweep_policy(cfg, seed):
    """Open-loop smooth random actions (no reference needed): sum of two sinusoids per channel, clipped to [-1, 1]."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.002, 0.03, (2, 6)); ph = rng.uniform(0, 6.28, (2, 6)); am = rng.uniform(0.2, 0.8, (2, 6))
    bias = rng.uniform(-0.3, 0.6, 6)

    def act(t, n):
        a = bias + am[0] * np.sin(f[0] * t + ph[0]) + am[1] * np.sin(f[1] * t + ph[1])
        return np.clip(np.tile(a, (n, 1)) + 0.15 * np.sin(0.37 * np.arange(n))[:, None], -1, 1).astype(np.float32)
    return act
