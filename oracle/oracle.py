"""TEST INFRASTRUCTURE -- ctypes wrapper of the C oracle (oracle/hlynr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from hlynr_intercept_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libhlynr_oracle.so")
    src = os.path.join(_HERE, "hlynr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(abi.HlynrParams), C.c_int64, C.c_uint64, C.c_int64, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_curriculum.argtypes = [C.c_void_p, C.POINTER(abi.HlynrCurriculum)]
        L.oracle_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_step_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 6 + [
            C.POINTER(abi.HlynrInfoSoA), C.c_void_p, C.c_int]
        L.oracle_get_stats.argtypes = [C.c_void_p, C.POINTER(abi.HlynrStats), C.c_int]
        L.oracle_export_state.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.oracle_kalman_decoupling_error.restype = C.c_double
        L.oracle_kalman_decoupling_error.argtypes = [C.c_void_p]
        L.oracle_params_size.restype = C.c_size_t
        L.oracle_env_state_size.restype = C.c_size_t
        L.oracle_philox.argtypes = [C.c_void_p] * 3
        L.oracle_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_void_p] * 3
        assert L.oracle_params_size() == C.sizeof(abi.HlynrParams)
        assert L.oracle_env_state_size() == C.sizeof(abi.HlynrEnvState)
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleBatch:
    """N oracle envs with the same call surface as oracle.ref_harness.RefBatch and the CUDA HlynrSim."""

    def __init__(self, params, curriculum, n_envs, seed=1234, env_id_offset=0, float64=False, threads=1):
        self.L = lib()
        self.n = int(n_envs)
        self.params = params
        self.h = self.L.oracle_create(C.byref(params), self.n, seed, env_id_offset, int(bool(float64)))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.set_curriculum(curriculum)
        self.threads = max(1, int(threads))
        self._pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None
        self.info = {n: np.zeros((self.n,) + shp, dtype=dt) for n, dt, shp in abi.INFO_FIELDS}
        self._info_struct = abi.HlynrInfoSoA(**{n: self.info[n].ctypes.data for n, _, _ in abi.INFO_FIELDS})

    def set_curriculum(self, cur):
        c = cur.to_struct() if hasattr(cur, "to_struct") else cur
        if isinstance(c, dict):
            c = abi.HlynrCurriculum(**c)
        self.L.oracle_set_curriculum(self.h, C.byref(c))

    def seed(self, s):
        self.L.oracle_seed(self.h, s)

    def reset(self, mask=None):
        obs = np.zeros((self.n, 26), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.oracle_reset(self.h, _ptr(m), _ptr(obs))
        return obs

    def step(self, actions, auto_reset=True, want_info=True):
        n = self.n
        actions = np.ascontiguousarray(actions, dtype=np.float32)
        assert actions.shape == (n, 6)
        obs = np.zeros((n, 26), np.float32)
        reward = np.zeros(n, np.float64)
        term = np.zeros(n, np.uint8)
        trunc = np.zeros(n, np.uint8)
        tobs = np.full((n, 26), np.nan, np.float32)
        margin = np.zeros(n, np.float64)
        info = C.byref(self._info_struct) if want_info else None

        def run(i0, i1):
            self.L.oracle_step_range(self.h, i0, i1, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(term), _ptr(trunc),
                                     _ptr(tobs), info, _ptr(margin), int(auto_reset))

        if self._pool is None:
            run(0, n)
        else:
            cuts = np.linspace(0, n, self.threads + 1).astype(np.int64)
            list(self._pool.map(lambda k: run(int(cuts[k]), int(cuts[k + 1])), range(self.threads)))
        self.margin = margin
        return obs, reward, term, trunc, tobs, ({k: v.copy() for k, v in self.info.items()} if want_info else None)

    def stats(self, zero_after=False):
        s = abi.HlynrStats()
        self.L.oracle_get_stats(self.h, C.byref(s), int(zero_after))
        return {n: getattr(s, n) for n in abi.STATS_FIELDS[:14]}

    def export_state(self):
        arr = np.zeros(self.n, dtype=abi.env_state_numpy_dtype())
        self.L.oracle_export_state(self.h, 0, self.n, _ptr(arr))
        return {k: arr[k].copy() for k in arr.dtype.names}

    def kalman_decoupling_error(self):
        return float(self.L.oracle_kalman_decoupling_error(self.h))

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None
        if self._pool:
            self._pool.shutdown()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox(_ptr(c), _ptr(k), _ptr(out))
    return out


def draws(seed, env_id, episode, step, blk):
    raw = np.zeros(4, np.uint32)
    uni = np.zeros(4, np.float32)
    nrm = np.zeros(4, np.float32)
    lib().oracle_draws(seed, env_id, episode, step, blk, _ptr(raw), _ptr(uni), _ptr(nrm))
    return raw, uni, nrm
