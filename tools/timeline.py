"""Pipelined (no host sync) timeline of the device rollout step: CUDA events between the policy forward and
the env step of every iteration, plus the policy's own stage events of the last iteration.

    python tools/timeline.py [--steps 60] [--warmup 30]
"""
import argparse
import ctypes as C
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--humans", type=int, default=20)
    a = ap.parse_args()
    import torch
    from crowdnav_prediction_attngraph_b200 import _capi
    from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
    from crowdnav_prediction_attngraph_b200.policy import Policy
    from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
    N = a.envs
    dev = torch.device("cuda", 0)
    env = CudaCrowdVecEnv(num_envs=N, nenv_total=N, rank_offset=0, seed=425, human_num=a.humans, device=dev)

    class Args(object):
        num_processes, seq_length, num_mini_batch = N, 30, 2
    torch.manual_seed(425)
    policy = Policy(env.observation_space.spaces, env.action_space, base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
    rollouts = RolloutStorage(30, N, env.observation_space.spaces, env.action_space, 128, 256, device=dev)
    obs = env.reset()
    for k in rollouts.obs:
        rollouts.obs[k][0].copy_(obs[k])
    eng = policy._engine(N, dev)
    noise = torch.randn(N, 2, device=dev)

    def step(ev=None, fixed_noise=False, mode=None):
        s = rollouts.step
        o = {k: v[s] for k, v in rollouts.obs.items()}
        hn = rollouts.recurrent_hidden_states['human_node_rnn']
        if ev:
            ev[0].record()
        nz = noise if fixed_noise else None
        if mode == "inplace":
            noise.normal_()
            nz = noise
        elif mode == "alloc_only":
            _ = torch.empty(N, 2, device=dev)
            nz = noise
        elif mode == "randn_before_event":
            nz = noise
        eng.act(o, hn[s], rollouts.masks[s], noise=nz,
                out=dict(value=rollouts.value_preds[s], action=rollouts.actions[s], log_prob=rollouts.action_log_probs[s],
                         h_out=hn[s + 1]))
        if ev:
            ev[1].record()
        env.step_device(rollouts.actions[s], obs_out={k: v[s + 1] for k, v in rollouts.obs.items()},
                        reward_out=rollouts.rewards[s], not_done_out=rollouts.masks[s + 1])
        if ev:
            ev[2].record()
        rollouts.step = (s + 1) % rollouts.num_steps
        if rollouts.step == 0:
            rollouts.after_update()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    out = {}
    for label, fixed, mode in (("randn_each_step", False, None), ("fixed_noise", True, None),
                               ("normal_inplace", True, "inplace"), ("alloc_only", True, "alloc_only"),
                               ("randn_each_step_again", False, None)):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps + 1)]
        import time
        c0 = time.perf_counter()
        for i in range(a.steps):
            step(evs[i], fixed, mode)
        cpu_ms = (time.perf_counter() - c0) * 1e3 / a.steps
        evs[a.steps][0].record()
        torch.cuda.synchronize()
        act = sum(evs[i][0].elapsed_time(evs[i][1]) for i in range(a.steps)) / a.steps
        envt = sum(evs[i][1].elapsed_time(evs[i][2]) for i in range(a.steps)) / a.steps
        rest = sum(evs[i][2].elapsed_time(evs[i + 1][0]) for i in range(a.steps)) / a.steps
        total = evs[0][0].elapsed_time(evs[a.steps][0]) / a.steps
        out[label] = dict(ms_per_step=total, act=act, env=envt, between=rest, cpu_enqueue_ms=cpu_ms)
    # pipelined stage events of the last act
    lib = eng.lib
    lib.cn_policy_profile(eng._h, 1)
    for _ in range(10):
        step()
    ns = lib.cn_policy_stage_count()
    buf = (C.c_float * ns)()
    _capi.check(lib, lib.cn_policy_stage_ms(eng._h, buf, ns), "stage_ms")
    lib.cn_policy_profile(eng._h, 0)
    out["pipelined_stage_ms"] = {lib.cn_policy_stage_name(i).decode(): round(buf[i], 4) for i in range(ns)}
    out["pipelined_stage_sum"] = sum(buf[i] for i in range(ns))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
