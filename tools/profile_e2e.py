"""Host-side profile of the train.py-contract loop (bench.py's e2e leg): where the 0.15 ms/step between the
device-resident loop and the reference-facing API goes.  python tools/profile_e2e.py [steps]"""
import cProfile, io, os, pstats, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
from crowdnav_prediction_attngraph_b200.policy import Policy
from crowdnav_prediction_attngraph_b200.storage import RolloutStorage

N, T, steps = 4096, 30, int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
env = CudaCrowdVecEnv(num_envs=N, seed=425, human_num=20, device=dev)
class Args(object):
    num_processes, seq_length, num_mini_batch = N, T, 2
torch.manual_seed(425)
policy = Policy(env.observation_space.spaces, env.action_space, base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
rollouts = RolloutStorage(T, N, env.observation_space.spaces, env.action_space, 128, 256, device=dev)
obs = env.reset()
for k in rollouts.obs:
    rollouts.obs[k][0].copy_(obs[k])
pin_masks, pin_bad, pin_rew = (torch.zeros(N, 1).pin_memory(), torch.ones(N, 1).pin_memory(), torch.zeros(N, 1).pin_memory())
acc = {"act": 0.0, "env.step": 0.0, "host-staging": 0.0, "insert": 0.0}
def e2e_step():
    t0 = time.perf_counter()
    s = rollouts.step
    o = {k: rollouts.obs[k][s] for k in rollouts.obs}
    hx = {'human_node_rnn': rollouts.recurrent_hidden_states['human_node_rnn'][s]}
    with torch.no_grad():
        value, action, logp, hx2 = policy.act(o, hx, rollouts.masks[s])
    t1 = time.perf_counter()
    nobs, reward, done, infos = env.step(action)
    t2 = time.perf_counter()
    pin_masks.copy_(torch.from_numpy(1.0 - done.astype("float32")).unsqueeze(1))
    pin_rew.copy_(reward)
    t3 = time.perf_counter()
    rollouts.insert(nobs, hx2, action, logp, value, pin_rew, pin_masks, pin_bad)
    if rollouts.step == 0:
        rollouts.after_update()
    t4 = time.perf_counter()
    acc["act"] += t1 - t0; acc["env.step"] += t2 - t1; acc["host-staging"] += t3 - t2; acc["insert"] += t4 - t3
for _ in range(60):
    e2e_step()
torch.cuda.synchronize()
for k in acc: acc[k] = 0.0
w0 = time.perf_counter()
for _ in range(steps):
    e2e_step()
torch.cuda.synchronize()
wall = (time.perf_counter() - w0) / steps * 1e3
print("e2e %.4f ms/step; host segments (ms/step):" % wall, {k: round(v / steps * 1e3, 4) for k, v in acc.items()})
pr = cProfile.Profile(); pr.enable()
for _ in range(steps):
    e2e_step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
