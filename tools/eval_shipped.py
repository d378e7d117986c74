"""One-off results check: run the reference's 500-case test protocol (rl/evaluation.py) for the SHIPPED policy
checkpoint trained_models/GST_predictor_rand/checkpoints/41665.pt on this engine (config 3: CrowdSimPredRealGST-v0 +
GST predictor, randomised humans, random goal changes, phase 'test') and print the metrics next to the shipped log
trained_models/GST_predictor_rand/test/test_41665.pt.log.

    python tools/eval_shipped.py path/to/41665.pt
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

SHIPPED_LOG = dict(success_rate=0.88, collision_rate=0.12, timeout_rate=0.00, avg_nav_time=14.14, path_length=20.08,
                   intrusion_ratio=8.35, min_intrusion_dist=0.41)
ckpt = sys.argv[1]
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
gst = dict(np.load(os.path.join(REPO, "tests", "golden", "gst_params.npz")))
# trained_models/GST_predictor_rand/configs/config.py: randomised humans, random goal changing, 20 humans, seed 425
d = _capi.default_config_dict(num_envs=500, nenv_total=1, seed=425, human_num=H, phase=2, test_size=500,
                              randomize_attributes=1, random_goal_changing=1, goal_change_chance=0.5)
t0 = time.time()
out = evaluate_batched(pol, None, "CrowdSimPredRealGST-v0", 425, 500, dev, cfg_dict=d, gst_params=gst)
out["wall_s"] = time.time() - t0
steps = out.pop("episode_steps")
out["mean_episode_steps"] = float(np.mean(steps))
print(json.dumps({"engine": out, "shipped_log": SHIPPED_LOG}))
