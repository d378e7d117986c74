"""Per-step device time of the rollout as the simulation evolves from the synchronized first episodes
to desynchronized steady state (CUDA events, windows of 10 steps).  python tools/step_series.py"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
from crowdnav_prediction_attngraph_b200.policy import Policy
from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
N = 4096
dev = torch.device("cuda", 0)
env = CudaCrowdVecEnv(num_envs=N, nenv_total=N, rank_offset=0, seed=425, human_num=20, device=dev)
class Args(object):
    num_processes, seq_length, num_mini_batch = N, 30, 2
torch.manual_seed(425)
policy = Policy(env.observation_space.spaces, env.action_space, base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
rollouts = RolloutStorage(30, N, env.observation_space.spaces, env.action_space, 128, 256, device=dev)
obs = env.reset()
for k in rollouts.obs:
    rollouts.obs[k][0].copy_(obs[k])
eng = policy._engine(N, dev)
def step():
    rollouts.rollout_step_zero_copy(eng, env)
    if rollouts.step == 0:
        rollouts.after_update()
for _ in range(5):
    step()
W = 10
evs = [torch.cuda.Event(enable_timing=True) for _ in range(81)]
rows, dones = [], []
for w in range(80):
    evs[w].record()
    for _ in range(W):
        step()
    rows.append(None)
evs[80].record()
torch.cuda.synchronize()
series = [round(evs[i].elapsed_time(evs[i + 1]) / W, 4) for i in range(80)]
print(json.dumps({"window_steps": W, "first_step": 5, "ms_per_step": series,
                  "valid_rows_end": int(eng.lib.cn_policy_last_rows(eng._h)),
                  "mean_ep_len_hint": float(torch.as_tensor(env.get_state("step_count")).float().mean())}))
