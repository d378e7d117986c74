"""Debug aid: CudaPretextVecEnv in two CN_GST_MODE settings, same actions, first difference."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from crowdnav_prediction_attngraph_b200.vec_env import CudaPretextVecEnv
N, H, T = 4, 20, 80
params = dict(np.load(os.path.join(REPO, "tests", "golden", "gst_params.npz")))
envs = []
for mode in sys.argv[1:3]:
    os.environ["CN_GST_MODE"] = mode
    envs.append(CudaPretextVecEnv(params, num_envs=N, human_num=H, seed=31, device="cuda:0"))
obs = [e.reset() for e in envs]
rng = np.random.RandomState(2)
for t in range(T):
    d = (obs[0]["spatial_edges"] - obs[1]["spatial_edges"]).abs().max().item()
    a = torch.from_numpy(rng.uniform(-1, 1, (N, 2)).astype(np.float32)).cuda()
    res = [e.step(a) for e in envs]
    obs = [r[0] for r in res]
    dr = (res[0][1] - res[1][1]).abs().max().item()
    vm = obs[0]["visible_masks"].sum(1).tolist()
    print("t=%d d_obs(before) %.2e d_rew %.2e done %s visible %s" % (t, d, dr, res[0][2].tolist(), vm))
    if dr > 1e-3 or d > 1e-3:
        print(res[0][1].reshape(-1), res[1][1].reshape(-1))
        e = int((res[0][1] - res[1][1]).abs().reshape(-1).argmax())
        np.set_printoptions(precision=4, suppress=True, linewidth=200)
        for k, o in enumerate(obs):
            print("mode", sys.argv[1 + k], "env", e, "robot", o["robot_node"][e].cpu().numpy().reshape(-1)[:2])
            print(o["spatial_edges"][e].cpu().numpy()[:8])
            print("vis", o["visible_masks"][e].int().cpu().numpy())
        print("max obs diff after step", (obs[0]["spatial_edges"] - obs[1]["spatial_edges"]).abs().max().item())
        break
