"""Debug aid: record the predictor inputs of a CudaPretextVecEnv rollout, replay them through fresh predictor handles in
two CN_GST_MODE settings (python tools/debug_gst_replay.py tcc tc)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import ctypes as C
import numpy as np, torch
from crowdnav_prediction_attngraph_b200 import _capi
from crowdnav_prediction_attngraph_b200.vec_env import CudaPretextVecEnv
N, H, T = 4, 20, 6
params = dict(np.load(os.path.join(REPO, "tests", "golden", "gst_params.npz")))
os.environ["CN_GST_MODE"] = sys.argv[2]
env = CudaPretextVecEnv(params, num_envs=N, human_num=H, seed=31, device="cuda:0")
rec = []
orig = env._process
def spy(obs, reward):
    rec.append((obs['robot_node'].clone(), obs['spatial_edges'].clone(), obs['visible_masks'].clone(),
                None if reward is None else reward.clone()))
    out = orig(obs, reward)
    rec[-1] = rec[-1] + (out['spatial_edges'].clone(), None if reward is None else reward.clone())
    return out
env._process = spy
env.reset()
rng = np.random.RandomState(2)
for t in range(T):
    env.step(torch.from_numpy(rng.uniform(-1, 1, (N, 2)).astype(np.float32)).cuda())
torch.cuda.synchronize()
print("recorded", len(rec), "calls; shapes", [tuple(x.shape) for x in rec[1][:3]], rec[1][2].dtype)

def make(mode):
    os.environ["CN_GST_MODE"] = mode
    lib = _capi.load_library()
    h = C.c_void_p()
    _capi.check(lib, lib.cn_gst_create(N, H, 5, 0.3, 0.3, -20.0, 0, C.byref(h)), "create")
    for k in params:
        a = np.ascontiguousarray(params[k], dtype=np.float32)
        _capi.check(lib, lib.cn_gst_set_param(h, k.encode(), a.ctypes.data, a.size), k)
    _capi.check(lib, lib.cn_gst_finalize(h), "finalize")
    _capi.check(lib, lib.cn_gst_reset(h, None), "reset")
    return lib, h
np.set_printoptions(precision=4, suppress=True, linewidth=200)
hs = [make(m) for m in sys.argv[1:3]]
for t, (r, s, v, rew, out_env, rew_env) in enumerate(rec):
    outs = []
    for lib, h in hs:
        rw = None if rew is None else rew.clone()
        pen = torch.zeros(N, device="cuda"); out = torch.zeros(N, H, 12, device="cuda")
        _capi.check(lib, lib.cn_gst_step(h, r.data_ptr(), s.data_ptr(), v.data_ptr(), None if rw is None else rw.data_ptr(),
                                         pen.data_ptr(), out.data_ptr(), None), "step")
        torch.cuda.synchronize()
        outs.append((out, pen))
    d = (outs[0][0] - outs[1][0]).abs()
    print("call %d: |a-b| %.3e  |b-env| %.3e  pen a %s b %s" % (t, d.max().item(), (outs[1][0] - out_env).abs().max().item(),
                                                               outs[0][1].cpu().numpy(), outs[1][1].cpu().numpy()))
    if d.max().item() > 1e-3:
        e = int(d.reshape(N, -1).max(1)[0].argmax())
        print(" env", e, "robot", r[e].cpu().numpy().reshape(-1)[:2], "vis", v[e].int().cpu().numpy())
        print(" in sp", s[e].cpu().numpy().reshape(H, -1)[:, :2].T)
        print(" a", outs[0][0][e].cpu().numpy()[:6]); print(" b", outs[1][0][e].cpu().numpy()[:6])
        break
