"""Episode JSONL export in the reference's logger format, for the tensor API (second half).

The reference writes one `<log_dir>/episodes/<episode_id>.jsonl` per episode through `UnifiedLogger`
(rl_system/logger.py:148-273): a `header` line, `state` lines per entity and tick (`interceptor`: position / fuel / action,
`missile`: position -- what inference.py:535-548 logs and the Unity replay tooling reads), optional `event` lines and a
`footer` line with the outcome and the metrics of inference.py:588-606.  When the reference's own scripts drive
`HlynrVecEnv` they keep using their own logger on the `infos`; this recorder is for callers of the tensor API, where
nothing per-env ever reaches the host unless asked: it follows ONE env of the batch (lazy: one 400-byte state export per
tick) and writes the same records.
"""
import json
import os
import time

import numpy as np


class EpisodeRecorder:
    def __init__(self, sim, log_dir, env_index=0, metadata=None):
        self.sim, self.env = sim, int(env_index)
        self.dir = os.path.join(log_dir, "episodes")
        os.makedirs(self.dir, exist_ok=True)
        self.metadata = dict(metadata or {})
        self.count, self.file, self._buf = 0, None, []
        self._t0 = self._ret = self._steps = self._min_d = None

    # ---- logger.py:148-273 ----
    def begin_episode(self, episode_id=None):
        self.count += 1
        self.episode_id = episode_id or f"ep_{self.count:06d}"
        self._t0, self._ret, self._steps, self._min_d = time.time(), 0.0, 0, float("inf")
        self.file = os.path.join(self.dir, self.episode_id + ".jsonl")
        with open(self.file, "w") as f:
            f.write(json.dumps({"type": "header", "episode_id": self.episode_id, "start_time": self._t0,
                                "metadata": self.metadata}) + "\n")
        self._buf = []

    def _state(self, entity_id, state):
        self._buf.append({"type": "state", "timestamp": time.time() - self._t0, "entity_id": entity_id, "state": state})
        if len(self._buf) >= 100:
            self._flush()

    def _flush(self):
        if self._buf:
            with open(self.file, "a") as f:
                for e in self._buf:
                    f.write(json.dumps(e) + "\n")
            self._buf = []

    def record_tick(self, action, reward, terminated, truncated, distance, intercepted):
        """Call after every sim.step(auto_reset=False) (or before the auto-reset overwrote the env) with the env's action
        and outputs; reads the env's state from the device.  Returns True when the episode ended (footer written)."""
        if self.file is None:
            self.begin_episode()
        s = self.sim.export_state(self.env, 1)
        self._state("interceptor", {"position": s["ipos"][0].astype(np.float32).tolist(), "fuel": float(s["fuel"][0]),
                                    "action": np.asarray(action, np.float32).tolist()})
        self._state("missile", {"position": s["mpos"][0].astype(np.float32).tolist()})
        self._ret += float(reward)
        self._steps += 1
        self._min_d = min(self._min_d, float(distance))
        if not (terminated or truncated):
            return False
        self._flush()
        end = time.time()
        outcome = "intercepted" if intercepted else "failed"     # inference.py:586
        metrics = {"total_reward": self._ret, "steps": self._steps, "final_distance": float(distance),
                   "min_distance": self._min_d, "fuel_used": float(s["fuel_used"][0])}
        with open(self.file, "a") as f:
            f.write(json.dumps({"type": "footer", "episode_id": self.episode_id, "end_time": end, "duration": end - self._t0,
                                "outcome": outcome, "metrics": metrics}) + "\n")
        self.file = None
        return True
