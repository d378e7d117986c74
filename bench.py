#!/usr/bin/env python
"""Rollout benchmark: env-steps/s of the PPO rollout phase (policy act + crowd_sim step +
rollout-storage insert, auto-resets included) on BASELINE.json config[1]:
CrowdSimPred-v0 / const_vel / 20 humans / HH+HR attention / 4096 environments per B200.

    python bench.py --gpus N --steps K --warmup W          # ours (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm (oracle port)

One "step" = one rollout step of every environment of the job.  Prints ONE JSON line (rank 0).
  value    : device-resident rollout (inputs already in HBM, no host round trip), CUDA-event timed
  e2e      : the same loop through the reference-facing VecEnv/Policy API — reward/done/info come
             back to the host every step and masks/rewards go host->device into the storage,
             exactly what the unchanged train.py loop does (train.py:177-191)
  roofline : dominant kernel (QKV projection GEMM), algorithmic FLOPs / CUDA-event time vs the
             measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline : the oracle port (oracle/crowd_env.py + oracle/policy_ref.py) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "env_steps_per_s_ppo_rollout"
UNIT = "env-steps/s"
HUMANS = 20
ENVS_PER_GPU = 4096
ROLLOUT_T = 30
B_ENV = 4 * (23 * HUMANS + 22) + 4 * (2 * 6 * HUMANS + 11) + 2      # 2934 B
B_POL = 4 * (2 * 6 * HUMANS + 10) + 1044                              # 2044 B


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=90)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("CN_GEMM_MODE", "1")),
                    help="1 = tcgen05 3xFP16 GEMMs (default), 0 = fp32 CUDA-core GEMMs")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--burn-in", type=int, default=400,
                    help="untimed rollout steps before the warm-up: all environments start their first episode "
                         "in lock-step (a transient with ~25 %% more work per step, tools/step_series.py); a "
                         "training run lives in the desynchronised steady state reached after ~250 steps")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One host process: k oracle environments + the PyTorch fp32 oracle policy, W + K rollout steps."""
    widx, k, warmup, steps, seed = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    try:                       # numpy's BLAS pool must not oversubscribe the cores (one worker per core)
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle.crowd_env import EnvConfig, OracleVecEnv
    from oracle.policy_ref import PolicyRef
    env = OracleVecEnv(EnvConfig(human_num=HUMANS), k, seed=seed, rank_offset=widx * k, nenv_total=1 << 20)
    torch.manual_seed(425)
    pol = PolicyRef(12)
    obs = env.reset()
    h = torch.zeros(k, 1, 128)
    masks = torch.ones(k, 1)
    std = pol.logstd().exp().detach()
    t0 = None
    for s in range(warmup + steps):
        if s == warmup:
            t0 = time.perf_counter()
        with torch.no_grad():
            value, mean, h = pol({kk: torch.from_numpy(v) for kk, v in obs.items()}, h, masks)
            action = mean + std * torch.randn_like(mean)
        obs, rew, done, infos = env.step(action.numpy())
        masks = torch.from_numpy(1.0 - done.astype(np.float32)).reshape(k, 1)
    return time.perf_counter() - t0


def cpu_rollout_rate(steps, warmup, envs_per_worker=4, workers=None):
    """env-steps/s of the oracle port using every host core (fork workers, like ShmemVecEnv)."""
    import multiprocessing as mp
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    workers = workers or len(os.sched_getaffinity(0)) or (os.cpu_count() or 1)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        times = pool.map(_cpu_worker, [(w, envs_per_worker, warmup, steps, 425) for w in range(workers)])
    total = workers * envs_per_worker * steps
    return total / max(times), workers, envs_per_worker


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__
    subprocess.check_call(["make", "-s", "-C", os.path.join(REPO, "oracle")])
    t0 = time.perf_counter()
    rate, workers, k = cpu_rollout_rate(a.steps, a.warmup)
    wall = time.perf_counter() - t0
    sample = "%d fork workers x %d oracle envs x %d rollout steps (oracle/crowd_env.py + rvo2_ref.cpp + policy_ref.py); no burn-in: the CPU cost per env-step is phase independent (6.8-7.2 ms per 4-env step over steps 0-320)" % (
        workers, k, a.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * workers * k / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CrowdSimPred-v0 const_vel, 20 humans, HH+HR attention rollout (bounded CPU sample: %d envs)" % (workers * k),
                   "parallelism": "%d host processes" % workers},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler(object):
    def __init__(self, index):
        self.path = tempfile.mktemp(suffix=".csv")
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for ln in open(self.path).read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 7 and f[0].replace(".", "").isdigit():
                rows.append(f)
        os.unlink(self.path)
        if not rows:
            return out
        sm = sorted(float(r[0]) for r in rows)
        hot = [x for x in sm if x >= 0.5 * max(sm)] or sm
        out["sm_mhz"] = hot[len(hot) // 2]
        out["sm_max_mhz"] = float(rows[0][1])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        out["samples"] = len(rows)
        try:
            out["power_w_max"] = max(float(r[2]) for r in rows)
        except ValueError:
            pass
        return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from crowdnav_prediction_attngraph_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv, Box
    from crowdnav_prediction_attngraph_b200.policy import Policy
    from crowdnav_prediction_attngraph_b200.storage import RolloutStorage

    N = a.envs_per_gpu
    env = CudaCrowdVecEnv(num_envs=N, nenv_total=N * world, rank_offset=rank * N, seed=425, human_num=HUMANS, device=dev)

    class Args(object):
        num_processes, seq_length, num_mini_batch = N, ROLLOUT_T, 2
    torch.manual_seed(425)
    policy = Policy(env.observation_space.spaces, env.action_space, base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
    policy._cuda = None
    os.environ["CN_GEMM_MODE"] = str(a.gemm_mode)
    rollouts = RolloutStorage(ROLLOUT_T, N, env.observation_space.spaces, env.action_space, 128, 256, device=dev)
    obs = env.reset()
    for k in rollouts.obs:
        rollouts.obs[k][0].copy_(obs[k])
    eng = policy._engine(N, dev)
    launches0 = env.launch_count() + eng.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident rollout (value)
    def device_step():
        rollouts.rollout_step_zero_copy(eng, env)
        if rollouts.step == 0:
            rollouts.after_update()

    for _ in range(a.burn_in):
        device_step()
    for _ in range(a.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = env.launch_count() + eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_env, t_pol = [], []
    e0.record()
    c0 = time.perf_counter()
    for _ in range(a.steps):
        device_step()
    cpu_enqueue_ms = (time.perf_counter() - c0) * 1000.0 / a.steps      # host time to enqueue one step
    e1.record()
    barrier()
    ms_value = e0.elapsed_time(e1)
    launches = env.launch_count() + eng.launch_count() - l0

    # ---------------------------------------------------------------- e2e: reference-facing API, host round trips
    def e2e_step():
        s = rollouts.step
        o = {k: rollouts.obs[k][s] for k in rollouts.obs}
        hx = {'human_node_rnn': rollouts.recurrent_hidden_states['human_node_rnn'][s]}
        with torch.no_grad():
            value, action, logp, hx2 = policy.act(o, hx, rollouts.masks[s])
        nobs, reward, done, infos = env.step(action)          # reward CPU tensor, done numpy, lazy infos
        masks = torch.from_numpy(1.0 - done.astype("float32")).unsqueeze(1).pin_memory()
        bad = torch.ones(N, 1).pin_memory()
        rollouts.insert(nobs, hx2, action, logp, value, reward.pin_memory(), masks, bad)   # H2D copies inside
        if rollouts.step == 0:
            rollouts.after_update()
        return done

    for _ in range(max(3, a.warmup // 2)):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    f0.record()
    for _ in range(a.steps):
        e2e_step()
    f1.record()
    barrier()
    wall_e2e = (time.perf_counter() - w0) * 1000.0
    ms_e2e = max(f0.elapsed_time(f1), 0.0)
    # the timed regions last tens of milliseconds, shorter than nvidia-smi's sampling period: keep the
    # SAME device step running (untimed) until the sampler has seen >= 1 s of load, then read it
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.2:
        for _ in range(50):
            device_step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed regions + 1.2 s of the same device step (sampling period 50 ms)"

    # ---------------------------------------------------------------- per-kernel timing (outside the timed regions)
    lib = eng.lib
    lib.cn_policy_profile(eng._h, 1)
    ns = lib.cn_policy_stage_count()
    import ctypes as C
    names = [lib.cn_policy_stage_name(i).decode() for i in range(ns)]
    acc = [0.0] * ns
    env_ms = 0.0
    reps = 5
    for _ in range(reps):
        s = rollouts.step
        o = {k: rollouts.obs[k][s] for k in rollouts.obs}
        value, action, logp, h_new = eng.act(o, rollouts.recurrent_hidden_states['human_node_rnn'][s], rollouts.masks[s])
        buf = (C.c_float * ns)()
        _capi.check(lib, lib.cn_policy_stage_ms(eng._h, buf, ns), "stage_ms")
        for i in range(ns):
            acc[i] += buf[i] / reps
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        nobs, rew, done, info = env.step_device(action)
        g1.record()
        torch.cuda.synchronize()
        env_ms += g0.elapsed_time(g1) / reps
        rollouts.insert(nobs, {'human_node_rnn': h_new}, action, logp, value, rew, (1.0 - done.float()).unsqueeze(1))
    lib.cn_policy_profile(eng._h, 0)
    stages = dict(zip(names, acc))
    rows_valid = int(lib.cn_policy_last_rows(eng._h))          # compacted human rows of the last act

    # max over ranks
    t = torch.tensor([ms_value, ms_e2e, wall_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e, wall_e2e = [float(x) for x in t]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_envs = N * world
    value = total_envs * a.steps / (ms_value / 1000.0)
    e2e_ms = max(ms_e2e, wall_e2e)
    e2e = total_envs * a.steps / (e2e_ms / 1000.0)
    peaks = {}
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    hbm = peaks.get("hbm_gbs", 6650.0)
    # QKV projection: algorithmic FLOPs of ONE launch = 2 * rows * 1536 * 512 over the rows actually
    # processed (padded humans are compacted out; the dense row count is reported next to it)
    qkv_flops = 2.0 * rows_valid * 1536 * 512
    qkv_ms = stages.get("qkv_gemm", 0.0)
    qkv_tf = qkv_flops / (qkv_ms / 1000.0) / 1e12 if qkv_ms > 0 else 0.0
    gemm_kernel = "cn_gemm_tc_kernel (tcgen05 3xFP16)" if a.gemm_mode == 1 else "cn_gemm_f32_kernel"
    roof_gemm = {"kernel": "%s QKV projection, rows=%d N=1536 K=512" % (gemm_kernel, rows_valid), "bound": "tensor",
                 "achieved": qkv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": qkv_tf / peak_tf if peak_tf else None,
                 "traffic": None, "peak_source": peak_src, "launch_ms": qkv_ms, "algorithmic_flops_per_launch": qkv_flops,
                 "mma_flops_issued_per_launch": qkv_flops * (3 if a.gemm_mode == 1 else 1), "dense_rows": N * HUMANS}
    # environment step: algorithmic bytes per launch = B_env * N / CUDA-event time
    env_gbs = B_ENV * N / (env_ms / 1000.0) / 1e9 if env_ms > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch at N=4096, H=20 (ncu --set full,
    # profiles/r1_env_step_kernel_summary.md, capture r1_env_step_v6): 11.24 MB + 0.45 MB
    env_traffic = 11.69e6 * N / 4096.0 if HUMANS == 20 else None
    roof_env = {"kernel": "cn_env_step_kernel (one rollout step of %d envs; cn_env_event_kernel runs on a side stream)" % N,
                "bound": "hbm",
                "achieved": env_gbs, "peak": hbm, "unit": "GB/s", "frac": env_gbs / hbm, "traffic": env_traffic,
                "traffic_source": "ncu capture profiles/r1_env_step_kernel_summary.md (v6), scaled by N/4096",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s", "launch_ms": env_ms,
                "algorithmic_bytes_per_launch": B_ENV * N,
                "note": "latency/divergence bound by construction (per-human ORCA LP), see DESIGN.md"}
    dominant_is_env = env_ms >= max(stages.values())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_value / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CrowdSimPred-v0 const_vel, 20 humans, HH+HR attention, %d envs per GPU (BASELINE configs[1])" % N,
                   "global_envs": total_envs, "rollout_T": ROLLOUT_T, "parallelism": "env-sharded dp%d" % world,
                   "weights": "random init (orthogonal), seed 425", "gemm_mode": a.gemm_mode,
                   "burn_in_steps": a.burn_in,
                   "state": "desynchronised steady state (episodes at mixed phases, auto-resets every step); the "
                            "lock-step first episodes right after reset() cost up to 0.69 ms/step (profiles/r1_step_series.json)",
                   "l2": "no flush: every step touches a different rollout-storage slot (30 slots x 4.3 MB of observations) plus ~170 MB of policy activations and 19 MB of env state, > the 126 MB L2"},
        "e2e": {"value": e2e, "unit": UNIT,
                "h2d_bytes_per_step": N * 4 * 3,                 # masks + bad_masks + reward into the storage
                "d2h_bytes_per_step": N * (4 + 1 + 4 + 4 + 8 + 4),   # reward, done, info, aux, ep_ret, ep_len
                "ms_per_step": e2e_ms / a.steps, "ms_per_step_cuda_events": ms_e2e / a.steps},
        "gpu_launches": int(launches), "cpu_enqueue_ms_per_step": cpu_enqueue_ms,
        "roofline": roof_env if dominant_is_env else roof_gemm,
        "roofline_other": roof_gemm if dominant_is_env else roof_env,
        "valid_human_rows": rows_valid, "mean_detected_humans": rows_valid / float(N),
        "breakdown_ms": {"env_step_kernel": env_ms, **stages},
        "hbm_roofline": {"bytes_per_env_step": B_ENV + B_POL, "achieved_gbs": value * (B_ENV + B_POL) / 1e9,
                         "peak_gbs": hbm, "frac": value * (B_ENV + B_POL) / 1e9 / hbm},
        "clocks": clocks,
    }
    if world == 1 and not a.no_cpu_baseline:
        # bounded sample: size the run for ~a.cpu_seconds of CPU work
        rate, workers, k = cpu_rollout_rate(steps=max(4, int(a.cpu_seconds / 0.035)), warmup=2)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": "port",
                                "sample": "%d fork workers x %d oracle envs, oracle/crowd_env.py + rvo2_ref.cpp + policy_ref.py (torch fp32, 1 thread each); no burn-in (CPU cost per env-step is phase independent)" % (workers, k)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
