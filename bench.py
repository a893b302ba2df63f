#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched Hlynr Intercept step on B200 (BASELINE.json metric).

A "step" is one tick (environment.py:605 step + SB3 auto-reset) of every env of the batch = ONE launch of
the sm_100a step kernel through the C ABI (hlynr_step), actions resident in HBM.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--envs-per-gpu 1048576]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, envs sharded by global id, no
                                                           per-step communication; one NCCL all-reduce of the
                                                           episode-statistics block per rollout)
  python bench.py --impl reference ...                     (reference arm: the CPU oracle port of the reference
                                                           step on the box's host cores; the reference itself is
                                                           Python and cannot travel to the GPU box)

Prints ONE JSON line (rank 0).  `value` = env-steps/s with inputs resident in HBM; `e2e` = the same metric
through the numpy VecEnv API (pinned H2D of actions, D2H of obs/reward/dones every step); `roofline` = the
step kernel against the measured HBM copy bandwidth; `cpu_baseline` = the oracle port on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per env-step, fp32 build, API mode (the roofline contract figure)
ALGO_BYTES = {"cfg2": 526, "cfg3": 630, "cfg3_radar": 630, "cfg4": 630, "cfg1": 630}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_traffic(workload, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the committed ncu --set full
    capture (profiles/traffic.json), scaled to n envs; None if no capture matches."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))[workload]
        return t["dram_bytes_per_launch"] * n / t["n_envs"]
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def count_between(self, t0, t1):
        return sum(1 for t, _ in self.rows if t0 <= t <= t1)

    def stop(self, windows=None):
        """windows: [(t0, t1), ...] perf_counter intervals whose samples count (None = all)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for t, r in self.rows:
            if windows is not None and not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(workload, seconds=12.0, n_envs=16384, threads=None):
    """The oracle port (C restatement of the reference step) on the host cores, bounded sample."""
    from hlynr_intercept_b200 import config
    from oracle import draws, oracle

    threads = threads or os.cpu_count() or 1
    P, cur = config.resolve_config(config.baseline_config(workload), warn_dead=False)
    sim = oracle.OracleBatch(P, cur, n_envs, seed=1234, threads=threads)
    sim.reset()
    rng = np.random.default_rng(0)
    acts = [rng.uniform(-1, 1, (n_envs, 6)).astype(np.float32) for _ in range(4)]
    for k in range(3):
        sim.step(acts[k % 4], want_info=False)
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        sim.step(acts[steps % 4], want_info=False)
        steps += 1
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": n_envs * steps / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"{n_envs} envs x {steps} ticks of {workload} ({dt:.1f} s), C oracle port of the reference step "
                      f"(the reference is Python: ~2.4e3 env-steps/s/core measured in the build container)"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation (oracle port) on all host threads, K bounded steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hlynr_intercept_b200 import config
    from oracle import oracle

    threads = os.cpu_count() or 1
    n_envs = args.ref_envs
    P, cur = config.resolve_config(config.baseline_config(args.workload), warn_dead=False)
    sim = oracle.OracleBatch(P, cur, n_envs, seed=1234, threads=threads)
    sim.reset()
    rng = np.random.default_rng(0)
    acts = [rng.uniform(-1, 1, (n_envs, 6)).astype(np.float32) for _ in range(4)]
    for k in range(args.warmup):
        sim.step(acts[k % 4], want_info=False)
    t0 = time.perf_counter()
    for k in range(args.steps):
        sim.step(acts[k % 4], want_info=False)
    dt = time.perf_counter() - t0
    v = n_envs * args.steps / dt
    line = {"impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "envs_per_step_sample": n_envs,
                       "actions": "U(-1,1)^6 float32"},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{n_envs} envs per step (bounded sample of the 2^20-env workload)"},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(w):
    return {"cfg4": "cfg4: medium scenario, physics v2.0 on, in-kernel auto-reset",
            "cfg2": "cfg2: medium scenario, physics v2.0 off",
            "cfg3": "cfg3: hard scenario, physics v2.0 on + domain randomization",
            "cfg1": "cfg1: easy scenario"}.get(w, w)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-envs", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused", type=int, default=64, help="k of the extra fused-rollout measurement (0 = skip)")
    ap.add_argument("--rollout-steps", type=int, default=24, help="n_steps of the extra on-device PPO rollout-collection measurement (0 = skip)")
    ap.add_argument("--rollout-envs", type=int, default=131072)
    ap.add_argument("--post-steps", type=int, default=200, help="steps of the extra step + frame-stack/normalise measurement (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from hlynr_intercept_b200.sim import HlynrSim
    from hlynr_intercept_b200.vec_env import HlynrVecEnv
    from hlynr_intercept_b200 import config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.envs_per_gpu
    K, W = args.steps, max(args.warmup, 3)
    env_cfg = config.baseline_config(args.workload)
    sim = HlynrSim(env_cfg, n_envs=n, device=local_rank, seed=1234, env_id_offset=rank * n, precision=args.precision,
                   warn_dead=False)
    sim.reset()
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    pool = [(torch.rand(n, 6, device=dev, generator=g) * 2 - 1).contiguous() for _ in range(4)]  # U(-1,1)^6 in HBM

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(W):
        sim.step(pool[k % 4], want_terminal_obs=False)
    stats_t = sim.stats_tensor()
    if world > 1:
        dist.all_reduce(stats_t)  # warm NCCL
    sim.stats(zero_after=True)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.perf_counter()
    e0.record()
    for k in range(K):
        sim.step(pool[k % 4], want_terminal_obs=False)
    e1.record()
    stats_t = sim.stats_tensor()  # one stats all-reduce per rollout (NVLink); tiny, latency-bound
    if world > 1:
        dist.all_reduce(stats_t)
    barrier()
    t_region1 = time.perf_counter()
    step_ms_total = e0.elapsed_time(e1)
    launches = sim.launch_count() - launches0   # the timed region's launches (read before any clock probe)
    tmax = torch.tensor([step_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    clock_windows, clock_note = [(t_region0, t_region1)], None
    if rank == 0 and sampler.proc and sampler.count_between(t_region0, t_region1) < 3:
        # a timed region shorter than a few 20 ms sampler periods: sample over an UNTIMED repeat of the same launches
        t_probe0 = time.perf_counter()
        while time.perf_counter() - t_probe0 < 0.4:
            for k in range(50):
                sim.step(pool[k % 4], want_terminal_obs=False)
            torch.cuda.synchronize()
        clock_windows.append((t_probe0, time.perf_counter()))
        clock_note = "timed region shorter than the sampler period: sampled over an untimed repeat of the same launches as well"
    clocks = sampler.stop(clock_windows) if rank == 0 else None
    if clocks is not None and clock_note:
        clocks["note"] = clock_note
    value = world * n * K / (total_ms * 1e-3)
    stats_host = stats_t.cpu().numpy().tolist()

    # extra: fused k-step rollout with the in-kernel random policy (state in registers between ticks)
    fused = None
    if args.fused > 0:
        sim.rollout(args.fused, None, want_obs=False)
        barrier()
        e0.record()
        reps = 3
        for _ in range(reps):
            sim.rollout(args.fused, None, want_obs=False)
        e1.record()
        barrier()
        tf = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fused = {"k": args.fused, "value": world * n * args.fused * reps / (float(tf.item()) * 1e-3), "unit": "env-steps/s",
                 "note": "hlynr_rollout: k ticks per launch, in-kernel Philox random policy, obs written once"}

    # extra: step + on-device VecFrameStack(4) + VecNormalize, the tensor a policy on the same GPU consumes
    post = None
    if args.post_steps > 0:
        from hlynr_intercept_b200.post import HlynrObsPipeline

        pipe = HlynrObsPipeline(sim, n_stack=4, training=True)
        pipe.reset()
        for k in range(5):
            pipe.step(pool[k % 4])
        barrier()
        e0.record()
        for k in range(args.post_steps):
            pipe.step(pool[k % 4])
        e1.record()
        barrier()
        tp = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        ms_post = float(tp.item()) / args.post_steps
        post = {"value": world * n / (ms_post * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_post, "n_stack": 4,
                "note": "hlynr_step writing into the frame ring + hlynr_post_step (stacked, normalised [N,104] output, running "
                        "mean/var updated every step, stacked terminal observations)",
                "algorithmic_bytes_per_env_step_post": 104 + 416 + 416 + 25}
        pipe.close()

    # extra: PPO rollout collection fully on the device (SURVEY 8f rank 4 / BASELINE config 5): policy forward (torch MLP
    # 104 -> 512 -> 512 -> 256, LayerNorm), env step, frame-stack/normalise, TimeLimit bootstrap, GAE; no host
    # synchronisation inside collect()
    roll = None
    if args.rollout_steps > 0:
        from hlynr_intercept_b200.post import HlynrObsPipeline
        from hlynr_intercept_b200.rollout import DeviceRolloutCollector, GaussianMlpPolicy

        n_roll = min(n, args.rollout_envs)
        rsim = HlynrSim(env_cfg, n_envs=n_roll, device=local_rank, seed=4321, env_id_offset=rank * n_roll, precision=args.precision,
                        warn_dead=False)
        rpipe = HlynrObsPipeline(rsim, n_stack=4, training=True)
        torch.manual_seed(rank)
        torch.backends.cuda.matmul.allow_tf32 = True   # the policy GEMMs are the caller's; TF32 tensor cores as PPO users run them
        torch.backends.cudnn.allow_tf32 = True
        pol = GaussianMlpPolicy(104, device=dev)
        col = DeviceRolloutCollector(rpipe, pol, args.rollout_steps)
        col.collect()
        barrier()
        e0.record()
        col.collect()
        e1.record()
        barrier()
        te_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te_, op=dist.ReduceOp.MAX)
        graphed = args.rollout_steps % col.graph_period() == 0
        if graphed:   # the whole collect() as ONE CUDA-graph launch (launch-bound otherwise: ~25 kernels per step)
            col.capture()
            col.replay()
            barrier()
            e0.record()
            col.replay()
            e1.record()
            barrier()
        tr_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr_, op=dist.ReduceOp.MAX)
        # the same loop without the policy network (random actions): what the simulator side costs
        class _NoPolicy(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.dummy = torch.nn.Parameter(torch.zeros(1, device=dev))

            def value(self, obs):
                return obs[:, 0].contiguous()

            def forward(self, obs):
                a = torch.rand(obs.shape[0], 6, device=obs.device) * 2 - 1
                return a, obs[:, 0], obs[:, 1]

        col2 = DeviceRolloutCollector(rpipe, _NoPolicy(), args.rollout_steps)
        col2._started = True
        col2.obs[col2.T].copy_(col.obs[col.T])
        col2.collect()
        barrier()
        e0.record()
        col2.collect()
        e1.record()
        barrier()
        tn_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tn_, op=dist.ReduceOp.MAX)
        roll = {"value": world * n_roll * args.rollout_steps / (float(tr_.item()) * 1e-3), "unit": "env-steps/s",
                "envs_per_gpu": n_roll, "n_steps": args.rollout_steps,
                "policy": "GaussianMlpPolicy 104->512->512->256 (pi and vf), LayerNorm, fp32 weights, TF32 GEMMs (torch / cuBLAS)",
                "cuda_graph": bool(graphed), "eager": world * n_roll * args.rollout_steps / (float(te_.item()) * 1e-3),
                "without_policy_network": world * n_roll * args.rollout_steps / (float(tn_.item()) * 1e-3),
                "timeout_bootstrap_overflow": int(col.overflow.item()),
                "note": "DeviceRolloutCollector.collect(): policy forward + hlynr_step + hlynr_post_step + "
                        "hlynr_bootstrap_timeouts per step, hlynr_gae at the end; buffers [T,N,*] resident in HBM"}
        rpipe.close()
        rsim.close()

    # e2e: the numpy VecEnv API a Stable-Baselines3 user calls (host buffers, H2D + D2H inside the timed region)
    venv = HlynrVecEnv(env_cfg, n_envs=n, device=local_rank, seed=99, env_id_offset=rank * n, precision=args.precision,
                       warn_dead=False, lazy_infos=True)
    venv.reset()
    venv.sim.rollout(1200, None, want_obs=False)  # age the episodes: the timed steps see the steady-state done rate
    rng = np.random.default_rng(rank)
    host_actions = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]  # ordinary (unpinned) numpy arrays
    for k in range(3):
        venv.step(host_actions[k % 2])
    barrier()
    n_done = 0
    t0 = time.perf_counter()
    for k in range(args.e2e_steps):
        _, rew_h, _, infos_h = venv.step(host_actions[k % 2])
        n_done += len(infos_h.records)  # finished episodes (terminal observation + info) arrive as compact records
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": world * n * args.e2e_steps / float(te.item()), "unit": "env-steps/s",
           "h2d_bytes_per_step": n * 6 * 4, "d2h_bytes_per_step": n * (26 * 4 + 4 + 1 + 1 + 1) + 4 + 200 * min(n, 4096),
           "api": "HlynrVecEnv.step(ordinary numpy actions) -> numpy obs, rewards, dones, infos (hlynr_step_host: chunks pipelined "
                  "over 3 streams; actions staged with streaming stores; obs by copy engine into alternating page-locked sets; "
                  "reward/terminated/truncated/dones written by the kernel straight into host memory; done episodes as compact records)",
           "steps": args.e2e_steps, "done_episodes_per_step": n_done / max(args.e2e_steps, 1),
           "ms_per_step": float(te.item()) / args.e2e_steps * 1e3}
    venv.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_step = ALGO_BYTES.get(args.workload, 630) * (1 if args.precision == "fp32" else 1)
        per_launch_ms = step_ms_total / K  # rank-0 kernel time per launch, CUDA events on the launching stream
        achieved = n * bytes_per_step / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "envs_per_gpu": n, "total_envs": n * world,
                       "mode": "API: one hlynr_step launch per tick, actions U(-1,1)^6 float32 resident in HBM",
                       "parallelism": f"env-sharded x{world}, no per-step communication, 1 NCCL stats all-reduce per rollout",
                       "l2": "per-tick working set (~0.6 GB state + I/O) exceeds the 126 MB L2, no flush needed",
                       "seed": 1234},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, n), "peak_source": peak_src, "kernel": "hlynr::step_kernel<float,false>",
                         "algorithmic_bytes_per_env_step": bytes_per_step, "units_per_launch": n,
                         "kernel_us_per_launch": per_launch_ms * 1e3},
            "fused_rollout": fused,
            "obs_pipeline": post,
            "rollout_collection": roll,
            "episode_stats": dict(zip(["episodes", "successes", "return_sum", "length_sum"], stats_host[:4])),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, seconds=args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
