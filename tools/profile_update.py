"""Where one PPO minibatch pass spends its time at the bench shape (torch profiler, CUDA time by kernel)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from torch.profiler import profile, ProfilerActivity
from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
from crowdnav_prediction_attngraph_b200.policy import Policy
from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
from crowdnav_prediction_attngraph_b200 import ppo
N, T = 4096, 30
dev = torch.device("cuda", 0)
class Args(object):
    num_processes, seq_length, num_mini_batch = N, T, 2
torch.manual_seed(425)
env = CudaCrowdVecEnv(num_envs=N, human_num=20, seed=425, device=dev)
policy = Policy(env.observation_space.spaces, env.action_space, base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
ro = RolloutStorage(T, N, env.observation_space.spaces, env.action_space, 128, 256, device=dev)
agent = ppo.PPO(policy, 0.2, 1, 2, 0.5, 0.0, lr=4e-5, eps=1e-5, max_grad_norm=0.5,
                matmul_precision='tf32' if '--tf32' in sys.argv else None)
obs = env.reset()
for k in ro.obs:
    ro.obs[k][0].copy_(obs[k])
eng = policy._engine(N, dev)
for _ in range(T):
    ro.rollout_step_zero_copy(eng, env)
ro.compute_returns(torch.zeros(N, 1, device=dev), True, 0.99, 0.95, False)
agent.update(ro)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    agent.update(ro)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
