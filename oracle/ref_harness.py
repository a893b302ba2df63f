"""TEST INFRASTRUCTURE (oracle) -- drives the UNMODIFIED reference environment with injected draws.

Only usable where /root/reference exists (the build container); it cannot travel to the GPU box.
It is used to (a) validate the C restatement in oracle/hlynr_oracle.c and (b) generate the golden
fixtures under tests/golden/ (see oracle/gen_golden.py).

Recipe = SURVEY Appendix B:
  * a 2-class `gymnasium` stub (the reference needs only gym.Env.reset and spaces.Box,
    rl_system/environment.py:6-7,15,192-197,355);
  * the four NumPy generators of the reference are replaced by tape objects that return the
    site-keyed Philox draws of include/hlynr_rng.h (oracle/draws.py):
      - global np.random (spawn uniforms environment.py:409-467, evasion :1105, AR(1) wind :1128)
      - observation_generator.rng (core.py:417,425,426,470,563)
      - enhanced_wind_model.rng (physics_models.py:372,381-384)
      - physics_randomizer.rng (physics_randomizer.py:166-214)
  * float64 mode: after every reset the state arrays are up-cast to float64 so that the
    integrator, ISA, drag, distance and reward run in float64 inside the same reference code.
"""
import os
import sys
import types

import numpy as np

from . import draws

REFERENCE_ROOT = os.environ.get("HLYNR_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rl_system", "environment.py"))


def install_shim():
    if "gymnasium" not in sys.modules:
        g = types.ModuleType("gymnasium")
        s = types.ModuleType("gymnasium.spaces")

        class Env:
            def reset(self, seed=None, options=None):
                return None

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        g.Env = Env
        s.Box = Box
        g.spaces = s
        sys.modules["gymnasium"] = g
        sys.modules["gymnasium.spaces"] = s
    p = os.path.join(REFERENCE_ROOT, "rl_system")
    if p not in sys.path:
        sys.path.insert(0, p)


class _Ctx:
    """Draw address of the env currently being stepped."""

    def __init__(self, seed, env_id):
        self.seed, self.env_id = seed, env_id
        self.episode, self.step = 0, 0
        self.uniform_idx = 0  # spawn uniforms consumed in this reset
        self.dr_idx = 0
        self._cache = {}

    def at(self, episode, step):
        self.episode, self.step = episode, step
        self.uniform_idx = 0
        self.dr_idx = 0
        self._cache = {}

    def raw(self, blk, step=None):
        st = self.step if step is None else step
        k = (blk, st)
        if k not in self._cache:
            self._cache[k] = draws.block(self.seed, self.env_id, self.episode, st, blk)
        return self._cache[k]

    def uni(self, blk, i, step=None):
        return float(draws.u01(self.raw(blk, step))[i])

    def nrm(self, blk, step=None):
        return draws.normals(self.raw(blk, step)).astype(np.float64)


class _GlobalRandomTape:
    """Replaces numpy.random inside the reference's environment module."""

    def __init__(self):
        self.ctx = None

    def seed(self, *a, **k):
        pass

    def uniform(self, low=0.0, high=1.0, size=None):
        lo = np.asarray(low, dtype=np.float64)
        hi = np.asarray(high, dtype=np.float64)
        shape = np.broadcast(lo, hi).shape
        n = int(np.prod(shape)) if shape else 1
        u = np.empty(n, dtype=np.float64)
        k = getattr(self.ctx, "volley_k", 1)   # reset draws 4 uniforms per missile first (environment.py:389-415), then the interceptor's
        for j in range(n):
            idx = self.ctx.uniform_idx
            if idx < 4 * k:
                m = idx // 4
                blk = draws.BLK_SPAWN0 if m == 0 else draws.blk_vspawn(m)
                u[j] = self.ctx.uni(blk, idx % 4, step=0)
            else:
                u[j] = self.ctx.uni(draws.BLK_SPAWN1 + (idx - 4 * k) // 4, (idx - 4 * k) % 4, step=0)
            self.ctx.uniform_idx += 1
        u = u.reshape(shape) if shape else u[0]
        out = lo + (hi - lo) * u
        return out if shape else float(out)

    def randn(self, *shape):
        caller = sys._getframe(1).f_code.co_name
        if caller == "_update_missile_state":
            # which missile of the volley: the caller's `missile_state` argument is an entry of env.missile_states
            ms = sys._getframe(1).f_locals["missile_state"]
            env = self.ctx.env
            m = next((k for k, x in enumerate(getattr(env, "missile_states", [])) if x is ms), 0)
            blk = draws.BLK_EVADE if m == 0 else draws.blk_vevade(m)
        elif caller == "_update_wind":
            blk = draws.BLK_WIND
        else:
            raise RuntimeError("unexpected randn site: " + caller)
        assert shape == (3,)
        return self.ctx.nrm(blk)[:3].copy()


class _NpProxy:
    """`np` as seen by the reference's environment module: numpy with .random replaced."""

    def __init__(self, tape):
        self.random = tape

    def __getattr__(self, name):
        return getattr(np, name)


class _ObsTape:
    def __init__(self, ctx):
        self.ctx = ctx

    def random(self):
        caller = sys._getframe(1).f_code.co_name
        i = {"compute_radar_detection": 1, "_compute_ground_radar_detection": 2,
             "_compute_datalink_quality": 3}[caller]
        return self.ctx.uni(draws.BLK_UNI, i)

    def normal(self, loc, scale, size=None):
        fr = sys._getframe(1)
        assert fr.f_code.co_name == "_compute_ground_radar_detection" and size == 3
        # position noise is drawn at core.py:425, velocity noise at core.py:426
        assert fr.f_lineno in (425, 426), fr.f_lineno
        blk = draws.BLK_GPOS if fr.f_lineno == 425 else draws.BLK_GVEL
        return loc + scale * self.ctx.nrm(blk)[:3]


class _WindTape:
    def __init__(self, ctx):
        self.ctx = ctx

    def random(self):
        return self.ctx.uni(draws.BLK_UNI, 0)

    def normal(self, loc, scale, size=None):
        assert size == 3
        # turbulence is drawn at physics_models.py:372, the gust direction at :382
        line = sys._getframe(1).f_lineno
        assert line in (372, 382), line
        blk = draws.BLK_WIND if line == 372 else draws.BLK_GUST_DIR
        return loc + scale * self.ctx.nrm(blk)[:3]

    def exponential(self, scale=1.0):
        return scale * float(draws.exponential(self.ctx.raw(draws.BLK_GUST_MAG)))


class _DRTape:
    def __init__(self, ctx):
        self.ctx = ctx

    def normal(self, loc=0.0, scale=1.0, size=None):
        assert size is None
        i = self.ctx.dr_idx
        self.ctx.dr_idx += 1
        z = self.ctx.nrm(draws.BLK_DR0 + i // 4, step=0)[i % 4]
        return float(loc + scale * z)

    def random(self):
        return 0.0


_GLOBAL_TAPE = _GlobalRandomTape()
_ENV_MODULE = None


def _environment_module():
    global _ENV_MODULE
    if _ENV_MODULE is None:
        install_shim()
        import environment as envmod  # the reference's rl_system/environment.py

        envmod.np = _NpProxy(_GLOBAL_TAPE)
        _ENV_MODULE = envmod
    return _ENV_MODULE


INFO_FLAG_BITS = dict(intercepted=1, hit_target=2, clamped=4, radar_detected=8, ground_detected=16,
                      crossed=32, fuze=64, kf_init=128)


class RefBatch:
    """N unmodified reference envs stepped with SB3-DummyVecEnv auto-reset semantics and injected draws."""

    def __init__(self, env_cfg, n_envs, seed=1234, env_id_offset=0, float64=False, training_step_count=None):
        envmod = _environment_module()
        self.n = n_envs
        self.float64 = float64
        self.seed = seed
        self.envs, self.ctxs = [], []
        for i in range(n_envs):
            env = envmod.InterceptEnvironment(dict(env_cfg))
            ctx = _Ctx(seed, env_id_offset + i)
            ctx.env = env
            ctx.volley_k = int(env.volley_size) if env.volley_mode else 1
            env.observation_generator.rng = _ObsTape(ctx)
            if env.enhanced_wind_model is not None:
                env.enhanced_wind_model.rng = _WindTape(ctx)
                env.enhanced_wind_model.seed = lambda s: None  # reset() would swap the tape out (environment.py:545-546)
            if env.physics_randomizer is not None:
                env.physics_randomizer.rng = _DRTape(ctx)
            if training_step_count is not None:
                env.set_training_step_count(training_step_count)
            self.envs.append(env)
            self.ctxs.append(ctx)
        self.episode = np.full(n_envs, -1, dtype=np.int64)
        self.ep_return = np.zeros(n_envs, dtype=np.float64)
        self.ep_length = np.zeros(n_envs, dtype=np.int64)

    # -- helpers -----------------------------------------------------------------------------
    def _upcast(self, env):
        if not self.float64:
            return
        for st in [env.interceptor_state] + list(env.missile_states):   # missile_state aliases missile_states[0]
            st["position"] = st["position"].astype(np.float64)
            st["velocity"] = st["velocity"].astype(np.float64)
        if getattr(env, "thrust_dynamics_enabled", False):
            env.interceptor_thrust_actual = env.interceptor_thrust_actual.astype(np.float64)
        env.interceptor_state["fuel"] = np.float64(env.interceptor_state["fuel"])

    def _reset_one(self, i):
        env, ctx = self.envs[i], self.ctxs[i]
        self.episode[i] += 1
        ctx.at(int(self.episode[i]), 0)
        _GLOBAL_TAPE.ctx = ctx
        obs, info = env.reset()
        self._upcast(env)
        self.ep_return[i] = 0.0
        self.ep_length[i] = 0
        return obs

    def curriculum(self):
        e = self.envs[0]
        og = e.observation_generator
        return dict(intercept_radius=float(e.get_current_intercept_radius()),
                    beam_width_deg=float(og.radar_beam_width),
                    onboard_reliability=float(og.onboard_detection_reliability),
                    ground_reliability=float(og.ground_detection_reliability))

    def reset(self):
        return np.stack([self._reset_one(i) for i in range(self.n)]).astype(np.float32)

    def step(self, actions, auto_reset=True):
        n = self.n
        obs = np.zeros((n, 26), np.float32)
        term_obs = np.full((n, 26), np.nan, np.float32)
        reward = np.zeros(n, np.float64)
        terminated = np.zeros(n, np.uint8)
        truncated = np.zeros(n, np.uint8)
        info = dict(distance=np.zeros(n), min_distance=np.zeros(n), fuel_remaining=np.zeros(n),
                    fuel_used=np.zeros(n), steps=np.zeros(n, np.int32), flags=np.zeros(n, np.uint8),
                    interceptor_pos=np.zeros((n, 3)), missile_pos=np.zeros((n, 3)),
                    episode_return=np.zeros(n), episode_length=np.zeros(n, np.int32),
                    missiles_intercepted=np.zeros(n, np.int32), missiles_remaining=np.zeros(n, np.int32),
                    missile_min_distances=np.zeros((n, 8)), radar_quality=np.zeros(n))
        for i in range(n):
            env, ctx = self.envs[i], self.ctxs[i]
            ctx.at(int(self.episode[i]), env.steps + 1)
            _GLOBAL_TAPE.ctx = ctx
            a = actions[i]
            if self.float64:
                a = np.asarray(a, dtype=np.float64)
            o, r, te, tr, inf = env.step(a)
            self.ep_return[i] += float(r)
            self.ep_length[i] += 1
            reward[i], terminated[i], truncated[i] = r, te, tr
            og = env.observation_generator
            fl = 0
            fl |= 1 if inf["intercepted"] else 0
            fl |= 2 if inf["missile_hit_target"] else 0
            fl |= 4 if inf["clamped"] else 0
            fl |= 8 if inf["radar_detected"] else 0
            fl |= 16 if og._last_ground_detection_info["detected"] else 0
            fl |= 32 if inf["crossed_threshold"] else 0
            fl |= 64 if inf["proximity_fuze_triggered"] else 0
            fl |= 128 if og.kalman_filter.initialized else 0
            info["flags"][i] = fl
            info["distance"][i] = inf["distance"]
            info["min_distance"][i] = inf["min_distance"]
            info["fuel_remaining"][i] = inf["fuel_remaining"]
            info["fuel_used"][i] = inf["fuel_used"]
            info["steps"][i] = inf["steps"]
            info["interceptor_pos"][i] = inf["interceptor_pos"]
            info["missile_pos"][i] = inf["missile_pos"]
            info["episode_return"][i] = self.ep_return[i]
            info["episode_length"][i] = self.ep_length[i]
            info["missiles_intercepted"][i] = inf["missiles_intercepted"]
            info["radar_quality"][i] = inf["radar_quality"]
            info["missiles_remaining"][i] = inf["missiles_remaining"]
            md = list(inf["missile_min_distances"])
            info["missile_min_distances"][i, :len(md)] = md
            if (te or tr) and auto_reset:
                term_obs[i] = o
                o = self._reset_one(i)
            obs[i] = o
        return obs, reward, terminated, truncated, term_obs, info

    def export_state(self):
        """Per-env mutable state in HlynrEnvState field order (dict of arrays)."""
        n = self.n
        out = dict(ipos=np.zeros((n, 3)), ivel=np.zeros((n, 3)), quat=np.zeros((n, 4)), fuel=np.zeros(n),
                   fuel_used=np.zeros(n), mpos=np.zeros((n, 3)), mvel=np.zeros((n, 3)), wind=np.zeros((n, 3)),
                   thrust=np.zeros((n, 3)), prev_d=np.zeros(n), last_d=np.zeros(n), min_d=np.zeros(n),
                   episode_return=np.zeros(n), kf_x=np.zeros((n, 6)), kf_P=np.zeros((n, 4)), T0=np.zeros(n),
                   base_cd=np.zeros(n), peak=np.zeros(n), steps=np.zeros(n, np.int32),
                   worsen_count=np.zeros(n, np.int32), crossed=np.zeros(n, np.int32),
                   kf_init=np.zeros(n, np.int32), onboard_delay=np.zeros(n, np.int32),
                   episode=np.zeros(n, np.int32), kf_decoupling_err=np.zeros(n),
                   vpos=np.zeros((n, 24)), vvel=np.zeros((n, 24)), vmin=np.zeros((n, 8)), vactive=np.zeros((n, 8), np.int32),
                   vcur=np.zeros(n, np.int32), vcount=np.zeros(n, np.int32))
        for i, env in enumerate(self.envs):
            if env.volley_mode:
                for k, ms in enumerate(env.missile_states):
                    out["vpos"][i, 3 * k:3 * k + 3], out["vvel"][i, 3 * k:3 * k + 3] = ms["position"], ms["velocity"]
                    out["vmin"][i, k], out["vactive"][i, k] = env.missile_min_distances[k], int(ms["active"])
                    if ms is env.missile_state:
                        out["vcur"][i] = k
                out["vcount"][i] = len(env.intercepted_missile_indices)
            s, m = env.interceptor_state, env.missile_state
            out["ipos"][i], out["ivel"][i], out["quat"][i] = s["position"], s["velocity"], s["orientation"]
            out["fuel"][i], out["fuel_used"][i] = s["fuel"], env.total_fuel_used
            out["mpos"][i], out["mvel"][i] = m["position"], m["velocity"]
            out["wind"][i] = env.current_wind
            if getattr(env, "thrust_dynamics_enabled", False):
                out["thrust"][i] = env.interceptor_thrust_actual
            out["prev_d"][i], out["last_d"][i] = env._prev_distance, env._last_distance
            out["min_d"][i] = env._episode_min_distance
            out["episode_return"][i] = self.ep_return[i]
            kf = env.observation_generator.kalman_filter
            out["kf_x"][i] = kf.state
            P = np.asarray(kf.P, dtype=np.float64)
            out["kf_P"][i] = (P[0, 0], P[0, 3], P[3, 0], P[3, 3])
            blockP = np.zeros((6, 6))
            for a in range(3):
                blockP[a, a], blockP[a, a + 3], blockP[a + 3, a], blockP[a + 3, a + 3] = out["kf_P"][i]
            out["kf_decoupling_err"][i] = np.abs(P - blockP).max()
            out["T0"][i] = env.atmospheric_model.constants.SEA_LEVEL_TEMPERATURE if env.atmospheric_model else 288.15
            out["base_cd"][i] = env.mach_drag_model.base_cd if env.mach_drag_model else 0.3
            out["peak"][i] = env.mach_drag_model.transonic_peak_multiplier if env.mach_drag_model else 3.0
            out["steps"][i] = env.steps
            out["worsen_count"][i] = env._distance_worsening_count
            out["crossed"][i] = int(env._crossed_threshold)
            out["kf_init"][i] = int(kf.initialized)
            buf = env.observation_generator.sensor_delay_buffer
            out["onboard_delay"][i] = buf.delay_samples if buf else 0
            out["episode"][i] = self.episode[i]
        return out


# ---- scripted policies (vectorised; identical code drives the reference and the CUDA path) --------
def policy_random(seed):
    """Returns f(t, n) -> float32 actions U(-1,1)^(n,6) from numpy's PCG64 (cfg1/cfg2 of BASELINE.json)."""
    rng = np.random.default_rng(seed)

    def f(t, obs):
        return rng.uniform(-1, 1, (obs.shape[0], 6)).astype(np.float32)

    return f


def policy_pursuit(gain=1.0):
    """Pure pursuit on the observation only (cf. scripts/paper_benchmark_suite.py PurePursuit): thrust along
    the filtered relative position obs[0:3] when a track exists, else along the ground-radar position
    obs[17:20], else straight up.  Angular command zero."""

    def f(t, obs):
        n = obs.shape[0]
        a = np.zeros((n, 6), np.float32)
        rel = obs[:, 0:3].copy()
        no_track = rel[:, 0] <= -1.5
        g = obs[:, 17:20]
        use_g = no_track & (g[:, 0] > -1.5)
        rel[use_g] = g[use_g]
        none = no_track & ~use_g
        rel[none] = np.array([0.0, 0.0, 1.0], np.float32)
        nrm = np.linalg.norm(rel, axis=1, keepdims=True) + np.float32(1e-9)
        a[:, 0:3] = np.clip(gain * rel / nrm, -1, 1)
        return a.astype(np.float32)

    return f


def policy_mixed(seed, bias=(0.3, 0.3, 0.7)):
    """Biased random thrust + random angular rates (exercises every action dimension; used where the observation
    frame gives no world direction to pursue, i.e. body_frame)."""
    rng = np.random.default_rng(seed)

    def f(t, obs):
        a = rng.uniform(-1, 1, (obs.shape[0], 6)).astype(np.float32)
        a[:, 0:3] = np.clip(0.5 * a[:, 0:3] + np.asarray(bias, np.float32), -1, 1)
        a[:, 3:6] *= np.float32(0.3)
        return a.astype(np.float32)

    return f


def policy_los_pn(seed, gain=3.0, jitter=0.15):
    """LOS-frame guidance on the observation only (observation_mode 'los_frame'): full thrust along the LOS
    (action[0]) plus proportional-navigation style corrections from the LOS rates obs[2], obs[3]
    (action[1], action[2]; same basis, core.py:812-845 / environment.py:1006-1020), small random jitter."""
    rng = np.random.default_rng(seed)

    def f(t, obs):
        n = obs.shape[0]
        a = np.zeros((n, 6), np.float32)
        track = obs[:, 0] > -1.5
        a[:, 0] = 1.0
        a[track, 1] = np.clip(gain * obs[track, 2], -1, 1)
        a[track, 2] = np.clip(gain * obs[track, 3], -1, 1)
        a += (jitter * rng.uniform(-1, 1, (n, 6))).astype(np.float32)
        return np.clip(a, -1, 1).astype(np.float32)

    return f


def policy_true_guidance(ref, los_frame=False, seed=17, jitter=0.05):
    """Scripted interceptor for fixtures that must reach the intercept / proximity-fuze / crossed-threshold branches:
    zero-effort-miss guidance computed from the TRUE simulator state of the reference envs (a fixture only needs an
    action sequence, it does not have to be a legal radar-only policy).  los_frame=True expresses the thrust in the
    LOS basis the reference's action transform uses (environment.py:965-1061)."""
    rng = np.random.default_rng(seed)

    def f(t, obs):
        n = obs.shape[0]
        a = np.zeros((n, 6), np.float32)
        for i, env in enumerate(ref.envs):
            ip = np.asarray(env.interceptor_state["position"], np.float64)
            iv = np.asarray(env.interceptor_state["velocity"], np.float64)
            mp = np.asarray(env.missile_state["position"], np.float64)
            mv = np.asarray(env.missile_state["velocity"], np.float64)
            rel, vrel = mp - ip, mv - iv
            rng_ = np.linalg.norm(rel) + 1e-9
            closing = max(-np.dot(rel, vrel) / rng_, 20.0)
            tgo = min(rng_ / closing, 12.0)
            zem = rel + vrel * tgo + np.array([0.0, 0.0, 0.5 * 9.81 * tgo * tgo])
            d = zem / (np.linalg.norm(zem) + 1e-9)
            if los_frame:
                lu = rel / rng_
                lr = np.cross(lu, [0.0, 0.0, 1.0])
                lh = lr / np.linalg.norm(lr) if np.linalg.norm(lr) > 1e-6 else np.array([1.0, 0.0, 0.0])
                lv = np.cross(lu, lh)
                d = np.array([np.dot(d, lu), np.dot(d, lh), np.dot(d, lv)])
            a[i, 0:3] = d
        a += (jitter * rng.uniform(-1, 1, (n, 6))).astype(np.float32)
        return np.clip(a, -1, 1).astype(np.float32)

    return f
