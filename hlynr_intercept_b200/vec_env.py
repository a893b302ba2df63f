"""HlynrVecEnv: the GPU simulator behind the Stable-Baselines3 `VecEnv` contract the reference uses.

Boundary: the reference builds `DummyVecEnv([lambda: Monitor(InterceptEnvironment(cfg))] * n)`
(rl_system/scripts/train_flat_ppo.py:344-371, inference.py:413) or `SubprocVecEnv` (scripts/train_hrl_pretrain.py:358)
and drives it through reset / step_async / step_wait / env_method / get_attr.  This class provides exactly that
surface over `hlynr_step_host` (numpy in, numpy out, pinned staging inside the C library), with the SB3
auto-reset semantics: on done the returned observation is the reset observation and
infos[i] = {..., 'terminal_observation', 'TimeLimit.truncated', 'episode': {'r','l','t'}}.

If stable_baselines3 / gymnasium are importable the class subclasses the real `VecEnv` and uses real
`spaces.Box`; otherwise it is duck-typed (neither package exists in the build image).
"""
import collections.abc
import ctypes as C
import itertools
import time

import numpy as np

from . import _lib, abi
from .sim import HlynrSim

try:  # pragma: no cover - not installed in the build image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
    from gymnasium import spaces as _spaces

    _HAVE_SB3 = True
except Exception:  # duck-typed stand-ins
    _HAVE_SB3 = False

    class _VecEnvBase:  # minimal mirror of stable_baselines3.common.vec_env.VecEnv
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs = num_envs
            self.observation_space = observation_space
            self.action_space = action_space
            self.reset_infos = [{} for _ in range(num_envs)]
            self.render_mode = None
            self._seeds = [None] * num_envs
            self._options = [{} for _ in range(num_envs)]

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

        @property
        def unwrapped(self):
            return self

    class _Box:
        def __init__(self, low, high, shape, dtype):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            return np.asarray(x).shape == self.shape

    class _spaces:  # noqa: N801
        Box = _Box


class _ObservationGeneratorView:
    """What `get_attr('observation_generator')[0]` exposes to the reference's callbacks
    (scripts/train_flat_ppo.py:221-232)."""

    def __init__(self, venv):
        self._v = venv

    @property
    def radar_beam_width(self):
        return self._v.sim.curriculum.beam_width

    @property
    def onboard_detection_reliability(self):
        return self._v.sim.curriculum.onboard_reliability

    @property
    def ground_detection_reliability(self):
        return self._v.sim.curriculum.ground_reliability

    @property
    def measurement_noise_level(self):
        return self._v.sim.curriculum.noise_level


class LazyInfos(collections.abc.Sequence):
    """`infos` of one step for a large batch: a read-only sequence of length num_envs whose entry i is the
    reference's info dict (+ the SB3 done keys) when env i finished an episode in this step, materialised on first
    access and cached (so wrappers that patch infos[i]['terminal_observation'] in place keep working), and one shared
    empty dict otherwise.  Creating it is O(1) Python work; vectorised consumers read `.records` (structured array,
    abi.done_record_numpy_dtype) and `.done_indices` instead of touching dicts at all.

    With HlynrVecEnv(info_arrays=True) the per-step info of EVERY env (environment.py:829-857: distance, fuel_remaining,
    radar flags ...) is available as well: `.arrays` is a dict of [N]-sized numpy arrays (fetched from the device on first
    access; valid until the next step), and entry i of an unfinished env is then the reference's full info dict built from
    row i instead of the shared empty dict."""

    def __init__(self, venv, records, now):
        self._venv, self.records, self._now = venv, records, now
        self._n = venv.num_envs
        self._index = None
        self._cache = {}
        self._arrays = None

    @property
    def done_indices(self):
        return self.records["env"]

    @property
    def arrays(self):
        """{name: array[N, ...]} of abi.INFO_FIELDS for the tick just executed (terminal tick for envs that finished)."""
        if self._arrays is None:
            self._arrays = self._venv._fetch_info_arrays()
        return self._arrays

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        if self._index is None:
            self._index = {e: k for k, e in enumerate(self.records["env"].tolist())}
        k = self._index.get(i)
        if k is None:
            if not self._venv.info_arrays:
                return self._venv._shared_info
            d = self._cache.get(-1 - i)
            if d is None:
                f = self.arrays
                d = self._cache[-1 - i] = self._venv._info_dicts(*[f[nm][i:i + 1] for nm in (
                    "flags", "distance", "min_distance", "fuel_remaining", "fuel_used", "steps", "missile_pos", "interceptor_pos",
                    "missiles_intercepted", "missiles_remaining", "missile_min_distances", "radar_quality")])[0]
            return d
        d = self._cache.get(k)
        if d is None:
            d = self._cache[k] = self._venv._done_info_dicts(self.records[k:k + 1], self._now)[0]
        return d


class HlynrVecEnv(_VecEnvBase):
    def __init__(self, env_cfg=None, n_envs=1, device=0, seed=1234, env_id_offset=0, precision="fp32", warn_dead=True,
                 lazy_infos=None, copy_outputs=None, radar_debug=None, obs_dim=26, info_arrays=False):
        self.sim = HlynrSim(env_cfg, n_envs=n_envs, device=device, seed=seed, env_id_offset=env_id_offset,
                            precision=precision, warn_dead=warn_dead, obs_dim=obs_dim)
        od = self.obs_dim = self.sim.obs_dim   # 26 = the reference's vector; 17 = its radar channels obs[0:17] (17-D layout)
        self.config = dict(env_cfg or {})
        n = self.sim.n
        obs_space = _spaces.Box(low=-2.0, high=1.0, shape=(od,), dtype=np.float32)   # environment.py:192-194
        act_space = _spaces.Box(low=-1.0, high=1.0, shape=(6,), dtype=np.float32)    # environment.py:195-197
        super().__init__(n, obs_space, act_space)
        # many envs: no per-env Python dict churn unless asked for
        self.lazy_infos = (n > 4096) if lazy_infos is None else bool(lazy_infos)
        self.copy_outputs = (n <= 65536) if copy_outputs is None else bool(copy_outputs)
        # info['radar_debug'] is a per-env Python dict the reference rebuilds every step for its episode logger: small batches only
        self.radar_debug = (n <= 64 and not self.lazy_infos) if radar_debug is None else (bool(radar_debug) and not self.lazy_infos)
        L = self.sim.L
        ptrs = [C.c_void_p() for _ in range(5)]
        _lib.check(L.hlynr_pinned_buffers(self.sim.h, *[C.byref(p) for p in ptrs]))

        def view(p, shape, ctype, dtype):
            cnt = int(np.prod(shape))
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(cnt,)).view(dtype).reshape(shape)

        self._act = view(ptrs[0], (n, 6), C.c_float, np.float32)
        pd = C.c_void_p()
        _lib.check(L.hlynr_pinned_done(self.sim.h, C.byref(pd)))
        # Two page-locked output sets that alternate from step to step: the arrays a step returns stay untouched while the
        # NEXT step runs (SB3's collect_rollouts reads `_last_obs` / `_last_episode_starts` of step k after step k+1 has
        # returned) and are overwritten by the one after it.  Set 0 is the C library's own pinned buffers, set 1 is torch
        # pinned memory; hlynr_step_host DMAs into either without a staging copy, and the kernel writes reward /
        # terminated / truncated / dones straight into them.
        import torch

        self._pinned = [torch.empty(shape, dtype=dt, pin_memory=True) for shape, dt in
                        (((n, od), torch.float32), ((n,), torch.float32), ((n,), torch.uint8), ((n,), torch.uint8), ((n,), torch.uint8))]
        self._sets = [
            (view(ptrs[1], (n, od), C.c_float, np.float32), view(ptrs[2], (n,), C.c_float, np.float32),
             view(ptrs[3], (n,), C.c_uint8, np.uint8), view(ptrs[4], (n,), C.c_uint8, np.uint8),
             view(pd, (n,), C.c_uint8, np.uint8)),
            tuple(t.numpy() for t in self._pinned)]
        self._set_ptrs = [([x.ctypes.data_as(C.c_void_p) for x in st[:4]], st[4].ctypes.data_as(C.c_void_p) if k else None,
                           st[4].view(np.bool_)) for k, st in enumerate(self._sets)]
        self._bind_set(0)
        self._info = None           # [N]-sized info arrays, only in the eager (small-n) mode
        self._info_struct = None
        # lazy infos + info_arrays: the kernel also fills the [N]-sized info arrays, read on request through infos.arrays
        self.info_arrays = bool(info_arrays) or not self.lazy_infos
        self.sim.set_option("host_info", 1 if self.info_arrays else 0)
        self._rec_dtype = abi.done_record_numpy_dtype()
        self._t_start = time.time()
        self._shared_info = {}
        self._pending = None
        self._pending_seed = None
        self.training_step_count = 0
        self.observation_generator = _ObservationGeneratorView(self)

    def _bind_set(self, k):
        self._cur = k
        self._obs, self._rew, self._term, self._trunc, _ = self._sets[k]
        self._out_ptrs, done_ptr, self._done = self._set_ptrs[k]   # _done: terminated | truncated, written by the kernel
        _lib.check(self.sim.L.hlynr_host_done_buffer(self.sim.h, done_ptr))

    # ---- VecEnv API -----------------------------------------------------------------------------------
    def reset(self):
        if self._pending_seed is not None:   # SB3 semantics: seed() takes effect at the next reset(), never mid-episode
            self.sim.seed(self._pending_seed)
            self._pending_seed = None
        _lib.check(self.sim.L.hlynr_reset_host(self.sim.h, None, self._obs.ctypes.data_as(C.c_void_p)))
        self.reset_infos = ([{} for _ in range(self.num_envs)] if not self.lazy_infos
                            else LazyInfos(self, np.zeros(0, dtype=self._rec_dtype), 0.0))
        return self._obs.copy() if self.copy_outputs else self._obs

    def step_async(self, actions):
        a = np.asarray(actions, dtype=np.float32)  # other callers pass float64 zeros (helpers/check_missile_trajectory.py:58)
        if a.shape != (self.num_envs, 6):
            a = a.reshape(self.num_envs, 6)
        if not a.flags.c_contiguous:
            np.copyto(self._act, a)
            a = self._act
        self._pending = a   # C-contiguous float32: staged chunk by chunk inside hlynr_step_host, overlapped with the copies

    def step_wait(self):
        assert self._pending is not None, "step_wait() without step_async()"
        a, self._pending = self._pending, None
        self._bind_set(self._cur ^ 1)
        _lib.check(self.sim.L.hlynr_step_host(self.sim.h, a.ctypes.data_as(C.c_void_p), *self._out_ptrs, None, 1))
        infos = self._build_infos()
        if self.copy_outputs:
            return self._obs.copy(), self._rew.copy(), self._done.copy(), infos
        return self._obs, self._rew, self._done, infos

    def done_records(self):
        """Finished episodes of the last step as a structured array (abi.done_record_numpy_dtype), a view of pinned
        memory owned by the C library: two buffers alternate, so it stays valid while the next step runs and is
        overwritten by the step after that."""
        ptr, cnt = C.c_void_p(), C.c_int32()
        _lib.check(self.sim.L.hlynr_done_records_host(self.sim.h, C.byref(ptr), C.byref(cnt)))
        if cnt.value == 0:
            return np.zeros(0, dtype=self._rec_dtype)
        buf = (C.c_char * (cnt.value * self._rec_dtype.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=self._rec_dtype, count=cnt.value)

    _INFO_KEYS = ("distance", "intercepted", "missile_hit_target", "fuel_remaining", "fuel_used", "clamped", "missile_pos",
                  "interceptor_pos", "steps", "radar_detected", "radar_quality", "ground_radar_detected", "volley_mode",
                  "volley_size", "missiles_intercepted", "missiles_remaining", "missile_min_distances", "min_distance",
                  "crossed_threshold", "precision_mode", "proximity_fuze_enabled", "proximity_fuze_triggered",
                  "proximity_kill_radius")

    def _info_dicts(self, fl, distance, min_distance, fuel_remaining, fuel_used, steps, missile_pos, interceptor_pos,
                    missiles_intercepted, missiles_remaining, missile_min_distances, radar_quality, extra=None):
        """The reference's info dict, environment.py:829-857 (single-missile mode), for a batch of envs given as
        arrays.  All conversions are vectorised; the per-env Python work is one dict(zip(keys, row)).
        extra: (terminal_obs[k,26], truncated_only list, episode dicts) appended as the SB3 done keys."""
        P = self.sim.params
        fl = np.asarray(fl).astype(np.int64)
        m = len(fl)
        bit = lambda mask: ((fl & mask) != 0).tolist()  # noqa: E731
        hit = (fl & abi.INFO_INTERCEPTED) != 0
        dist = np.asarray(distance).tolist()
        rep = itertools.repeat
        vk = int(P.volley_size)   # environment.py:844-848
        mmd = np.asarray(missile_min_distances, dtype=np.float32).reshape(m, -1)[:, :max(vk, 1)].tolist()
        cols = [dist, hit.tolist(), bit(abi.INFO_HIT_TARGET), np.asarray(fuel_remaining).tolist(), np.asarray(fuel_used).tolist(),
                bit(abi.INFO_CLAMPED), list(np.array(missile_pos, dtype=np.float32)), list(np.array(interceptor_pos, dtype=np.float32)),
                np.asarray(steps).tolist(), bit(abi.INFO_RADAR_DETECTED), np.asarray(radar_quality, dtype=np.float64).tolist(), bit(abi.INFO_GROUND_DETECTED),
                rep(vk > 0, m), rep(max(vk, 1), m), np.asarray(missiles_intercepted).tolist(), np.asarray(missiles_remaining).tolist(), mmd,
                np.asarray(min_distance).tolist(), bit(abi.INFO_CROSSED), rep(bool(P.precision_mode), m), rep(bool(P.fuze_enabled), m),
                bit(abi.INFO_FUZE), rep(float(P.kill_radius), m)]
        keys = self._INFO_KEYS
        if extra is not None:
            keys = keys + ("terminal_observation", "TimeLimit.truncated", "episode")
            cols += [list(extra[0]), extra[1], extra[2]]
        return [dict(zip(keys, row)) for row in zip(*cols)]

    def _done_info_dicts(self, rec, now):
        fl = rec["flags"]
        tl = (((fl & abi.DONE_TRUNCATED) != 0) & ((fl & abi.DONE_TERMINATED) == 0)).tolist()
        ret, ln = rec["episode_return"].astype(np.float64).round(6).tolist(), rec["steps"].tolist()
        episodes = [{"r": r, "l": l, "t": now} for r, l in zip(ret, ln)]
        return self._info_dicts(fl, rec["distance"], rec["min_distance"], rec["fuel_remaining"], rec["fuel_used"],
                                rec["steps"], rec["missile_pos"], rec["interceptor_pos"], rec["missiles_intercepted"],
                                rec["missiles_remaining"], rec["missile_min_distances"],
                                np.where((fl & abi.DONE_ONBOARD_FILL) != 0, 0.0, float(self.sim.params.radar_quality)),
                                extra=(rec["terminal_obs"][:, :self.obs_dim].copy(), tl, episodes))

    def _build_infos(self):
        n = self.num_envs
        rec = self.done_records()
        now = round(time.time() - self._t_start, 6)
        if self.lazy_infos:
            return LazyInfos(self, rec, now)   # a view: the C library alternates two record buffers, valid through the next step
        f = self._fetch_info_arrays()
        infos = self._info_dicts(f["flags"], f["distance"], f["min_distance"], f["fuel_remaining"], f["fuel_used"],
                                 f["steps"], f["missile_pos"], f["interceptor_pos"], f["missiles_intercepted"],
                                 f["missiles_remaining"], f["missile_min_distances"], f["radar_quality"])
        if len(rec):
            for i, d in zip(rec["env"].tolist(), self._done_info_dicts(rec, now)):
                infos[i] = d
        if self.radar_debug:
            self._add_radar_debug(infos, f, rec)
        return infos

    def _fetch_info_arrays(self):
        if not self.info_arrays:
            raise _lib.HlynrError("per-env info arrays were not requested: HlynrVecEnv(..., info_arrays=True)")
        if self._info is None:
            n = self.num_envs
            self._info = {nm: np.zeros((n,) + shp, dtype=dt) for nm, dt, shp in abi.INFO_FIELDS}
            self._info_struct = abi.HlynrInfoSoA(**{nm: self._info[nm].ctypes.data for nm, _, _ in abi.INFO_FIELDS})
        _lib.check(self.sim.L.hlynr_info_host(self.sim.h, C.byref(self._info_struct)))
        return self._info

    def _add_radar_debug(self, infos, f, rec):
        """info['radar_debug'] (core.py:649-682, consumed by inference.py:546): small batches only, see radar_debug.py."""
        from . import radar_debug as rd

        P, beam = self.sim.params, self.sim.curriculum.beam_width
        st = self.sim.export_state()
        done = set(rec["env"].tolist()) if len(rec) else set()
        for i in range(self.num_envs):
            if i in done:
                continue
            infos[i]["radar_debug"] = rd.radar_debug(P, beam, st["ipos"][i], st["quat"][i], st["mpos"][i], int(f["steps"][i]),
                                                     int(f["flags"][i]), self._obs[i], onboard_delay=int(st["onboard_delay"][i]))
        for k, i in enumerate(rec["env"].tolist() if len(rec) else []):   # terminal step: state from the done record
            r = rec[k]
            if P.obs_mode != abi.OBS_WORLD:
                infos[i]["radar_debug"] = None   # the terminal observation carries no euler angles in body / los frame
                continue
            e = r["terminal_obs"][9:12].astype(np.float64) * np.pi
            infos[i]["radar_debug"] = rd.radar_debug(P, beam, r["interceptor_pos"], None, r["missile_pos"], int(r["steps"]),
                                                     int(r["flags"]) & 0xff, r["terminal_obs"],
                                                     forward=rd.forward_from_euler(e[0], e[1], e[2]))

    def close(self):
        self.sim.close()

    def seed(self, seed=None):
        """SB3 VecEnv.seed: stored and applied by the next reset() (env i gets seed + i there; here env i is keyed by
        (seed, global env id), which is the same kind of per-env stream).  Re-keying Philox in the middle of running episodes
        would change their draws from one tick to the next."""
        if seed is not None:
            self._pending_seed = int(seed)
        return [None if seed is None else int(seed) + i for i in range(self.num_envs)] if self.num_envs <= 4096 else [seed] * self.num_envs

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        idx = self._indices(indices)
        if method_name == "set_training_step_count":   # scripts/train_flat_ppo.py:173-177, every step
            self.training_step_count = int(method_args[0]) if method_args else int(method_kwargs["step_count"])
            self.sim.set_training_step_count(self.training_step_count)
            return [None for _ in idx]
        if method_name == "get_current_intercept_radius":  # scripts/train_hrl_pretrain.py:157
            r = self.sim.get_current_intercept_radius()
            return [r for _ in idx]
        if method_name == "seed":                           # scripts/compare_policies.py:150
            self.seed(method_args[0] if method_args else method_kwargs.get("seed"))
            return [None for _ in idx]
        raise AttributeError(f"HlynrVecEnv.env_method: '{method_name}' is not part of the accelerated path")

    def get_attr(self, attr_name, indices=None):
        idx = list(self._indices(indices))
        if attr_name == "get_current_intercept_radius":
            return [self.sim.get_current_intercept_radius for _ in idx]
        if attr_name == "observation_generator":
            return [self.observation_generator for _ in idx]
        if attr_name in ("interceptor_state", "missile_state"):   # visualize.py:149-196
            out = []
            for i in idx:
                s = self.sim.export_state(i, 1)
                if attr_name == "interceptor_state":
                    out.append({"position": s["ipos"][0].astype(np.float32), "velocity": s["ivel"][0].astype(np.float32),
                                "orientation": s["quat"][0].astype(np.float32), "angular_velocity": np.zeros(3, np.float32),
                                "fuel": float(s["fuel"][0]), "active": True})
                else:
                    out.append({"position": s["mpos"][0].astype(np.float32), "velocity": s["mvel"][0].astype(np.float32),
                                "orientation": np.array([1, 0, 0, 0], np.float32), "angular_velocity": np.zeros(3, np.float32),
                                "active": True})
            return out
        if attr_name in ("training_step_count", "config", "render_mode"):
            return [getattr(self, attr_name) for _ in idx]
        if attr_name in ("max_steps", "dt"):
            return [getattr(self.sim.params, attr_name) for _ in idx]
        raise AttributeError(f"HlynrVecEnv.get_attr: '{attr_name}' is not exposed")

    def set_attr(self, attr_name, value, indices=None):
        if attr_name == "training_step_count":
            self.env_method("set_training_step_count", value)
            return
        raise AttributeError(f"HlynrVecEnv.set_attr: '{attr_name}' cannot be set")

    def env_is_wrapped(self, wrapper_class, indices=None):
        # Monitor statistics (info['episode']) are produced natively, so evaluate_policy's Monitor check passes
        name = getattr(wrapper_class, "__name__", "")
        return [name == "Monitor" for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None
