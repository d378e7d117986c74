"""cProfile of the reference-facing (host round trip) rollout step, to see where the host time goes.
python tools/profile_e2e.py"""
import cProfile, io, os, pstats, sys, time
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
def e2e_step():
    s = rollouts.step
    o = {k: rollouts.obs[k][s] for k in rollouts.obs}
    hx = {'human_node_rnn': rollouts.recurrent_hidden_states['human_node_rnn'][s]}
    with torch.no_grad():
        value, action, logp, hx2 = policy.act(o, hx, rollouts.masks[s])
    nobs, reward, done, infos = env.step(action)
    masks = torch.from_numpy(1.0 - done.astype("float32")).unsqueeze(1).pin_memory()
    bad = torch.ones(N, 1).pin_memory()
    rollouts.insert(nobs, hx2, action, logp, value, reward.pin_memory(), masks, bad)
    if rollouts.step == 0:
        rollouts.after_update()
for _ in range(300):
    e2e_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    e2e_step()
torch.cuda.synchronize()
print("e2e ms/step", (time.perf_counter() - t0) * 1e3 / 200)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    e2e_step()
torch.cuda.synchronize()
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:6000])
