"""One-off results check: run the reference's 500-case test protocol (rl/evaluation.py) for the SHIPPED policy
checkpoint trained_models/GST_predictor_rand/checkpoints/41665.pt on this engine (config 3: CrowdSimPredRealGST-v0 +
GST predictor, randomised humans, random goal changes, phase 'test') and print the metrics next to the shipped log
trained_models/GST_predictor_rand/test/test_41665.pt.log.

    python tools/eval_shipped.py path/to/41665.pt rand
    python tools/eval_shipped.py path/to/41200.pt non_rand      (needs local_ckpt/gst_params_nonrand.npz)
The checkpoint is not part of this repository (10 MB reference artefact)."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from crowdnav_prediction_attngraph_b200 import _capi
from crowdnav_prediction_attngraph_b200.vec_env import Box
from crowdnav_prediction_attngraph_b200.policy import Policy
from crowdnav_prediction_attngraph_b200.evaluation import evaluate_batched

SHIPPED = {
    # trained_models/GST_predictor_rand: randomised humans + goal changes, seed 425, predictor ..._seed_1000_rand
    "rand": dict(log=dict(success_rate=0.88, collision_rate=0.12, timeout_rate=0.00, avg_nav_time=14.14, path_length=20.08,
                          intrusion_ratio=8.35, min_intrusion_dist=0.41), seed=425, randomize=1, goal_changing=1,
                 gst=os.path.join(REPO, "tests", "golden", "gst_params.npz")),
    # trained_models/GST_predictor_non_rand: fixed human attributes, no goal changes, seed 125, predictor ..._seed_1000
    "non_rand": dict(log=dict(success_rate=0.92, collision_rate=0.07, timeout_rate=0.00, avg_nav_time=15.42, path_length=20.96,
                              intrusion_ratio=4.23, min_intrusion_dist=0.44), seed=125, randomize=0, goal_changing=0,
                     gst=os.path.join(REPO, "local_ckpt", "gst_params_nonrand.npz")),
}
ckpt = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else "rand"
cfgw = SHIPPED[which]
SHIPPED_LOG = cfgw["log"]
dev = torch.device("cuda", 0)
sd = torch.load(ckpt, map_location="cpu", weights_only=True)
H = 20


class Args(object):
    num_processes, seq_length, num_mini_batch = 500, 30, 2


spaces = {'robot_node': Box((1, 7)), 'temporal_edges': Box((1, 2)), 'spatial_edges': Box((H, 12)), 'detected_human_num': Box((1,)),
          'visible_masks': Box((H,), np.bool_)}
pol = Policy(spaces, Box((2,)), base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
missing = pol.load_state_dict(sd, strict=False)
print("load_state_dict:", missing)
gst = dict(np.load(cfgw["gst"]))
# trained_models/GST_predictor_rand/configs/config.py: randomised humans, random goal changing, 20 humans, seed 425
d = _capi.default_config_dict(num_envs=500, nenv_total=1, seed=cfgw["seed"], human_num=H, phase=2, test_size=500,
                              randomize_attributes=cfgw["randomize"], random_goal_changing=cfgw["goal_changing"],
                              goal_change_chance=0.5)
t0 = time.time()
out = evaluate_batched(pol, None, "CrowdSimPredRealGST-v0", cfgw["seed"], 500, dev, cfg_dict=d, gst_params=gst)
out["wall_s"] = time.time() - t0
steps = out.pop("episode_steps")
out["mean_episode_steps"] = float(np.mean(steps))
print(json.dumps({"engine": out, "shipped_log": SHIPPED_LOG}))
