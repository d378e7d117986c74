"""Informational: one full training iteration at the bench shape (30-step device rollout of N envs, GAE, PPO update
in PyTorch on the device = SURVEY rows a22/a23, 'stays in PyTorch').  python tools/bench_update.py [--envs 4096]"""
import argparse, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
from crowdnav_prediction_attngraph_b200.policy import Policy
from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
from crowdnav_prediction_attngraph_b200 import ppo

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--tf32", action="store_true")
ap.add_argument("--no-pack", action="store_true")
a = ap.parse_args()
N, T = a.envs, 30
dev = torch.device("cuda", 0)


class Args(object):
    num_processes, seq_length, num_mini_batch = N, T, 2


torch.manual_seed(425)
env = CudaCrowdVecEnv(num_envs=N, human_num=20, seed=425, device=dev)
policy = Policy(env.observation_space.spaces, env.action_space, base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
ro = RolloutStorage(T, N, env.observation_space.spaces, env.action_space, 128, 256, device=dev)
policy.pack_valid_rows = not a.no_pack
agent = ppo.PPO(policy, 0.2, 5, 2, 0.5, 0.0, lr=4e-5, eps=1e-5, max_grad_norm=0.5,
                matmul_precision='tf32' if a.tf32 else None)
obs = env.reset()
for k in ro.obs:
    ro.obs[k][0].copy_(obs[k])
out = []
for it in range(a.iters):
    eng = policy._engine(N, dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(T):
        ro.rollout_step_zero_copy(eng, env)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with torch.no_grad():
        o = {k: ro.obs[k][-1] for k in ro.obs}
        hx = {k: ro.recurrent_hidden_states[k][-1] for k in ro.recurrent_hidden_states}
        nv = policy.get_value(o, hx, ro.masks[-1]).detach()
    ro.compute_returns(nv, True, 0.99, 0.95, False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    agent.update(ro)
    ro.after_update()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    out.append(dict(rollout_ms=(t1 - t0) * 1e3, returns_ms=(t2 - t1) * 1e3, update_ms=(t3 - t2) * 1e3))
print(json.dumps({"tf32": a.tf32, "pack_valid_rows": not a.no_pack, "envs": N, "rollout_T": T, "ppo_epoch": 5, "num_mini_batch": 2, "iterations": out,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
