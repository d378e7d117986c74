"""GPU: test-phase evaluation.  The batched evaluation (test_size parallel environments) must
report exactly what the reference's sequential protocol reports when both run on the CUDA engine, and the
protocol's seeding quirks (two resets per episode, case counter wrap) must hold."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _policy(N, dev):
    from crowdnav_prediction_attngraph_b200.vec_env import Box
    from crowdnav_prediction_attngraph_b200.policy import Policy, make_reference_like_state_dict

    class Args(object):
        num_processes, seq_length, num_mini_batch = N, 30, 2
    spaces = {'robot_node': Box((1, 7)), 'temporal_edges': Box((1, 2)), 'spatial_edges': Box((20, 12)),
              'detected_human_num': Box((1,))}
    pol = Policy(spaces, Box((2,)), base_kwargs=Args(), base='selfAttn_merge_srnn').to(dev)
    sd = make_reference_like_state_dict(12, seed=5)
    # a goal-seeking bias so that episodes end in all three ways within the time limit
    pol.load_state_dict(sd, strict=False)
    return pol


def test_batched_evaluation_equals_sequential_protocol():
    from crowdnav_prediction_attngraph_b200 import _capi
    from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
    from crowdnav_prediction_attngraph_b200.evaluation import evaluate, evaluate_batched
    dev = torch.device("cuda:0")
    test_size = 7                       # odd on purpose: the case counter wraps at test_size (0,2,4,6,1,3,5)
    d = _capi.default_config_dict(num_envs=1, nenv_total=1, seed=19, human_num=20, phase=2, test_size=test_size,
                                  time_limit=20.0)
    pol = _policy(1, dev)
    env = CudaCrowdVecEnv(device=dev, cfg=d)
    seq = evaluate(pol, env, 1, dev, test_size, None, None, None)
    env.close()
    bat = evaluate_batched(pol, None, "CrowdSimPred-v0", 19, test_size, dev, cfg_dict=d)
    assert seq["episode_steps"] == bat["episode_steps"]
    for k in ("success_rate", "collision_rate", "timeout_rate", "collision_cases", "timeout_cases"):
        assert seq[k] == bat[k], k
    for k in ("avg_nav_time", "path_length", "intrusion_ratio", "mean_episode_reward"):
        assert seq[k] == pytest.approx(bat[k], rel=1e-12, abs=1e-12), k
    if not np.isnan(seq["min_intrusion_dist"]):
        assert seq["min_intrusion_dist"] == pytest.approx(bat["min_intrusion_dist"], rel=1e-9)
    assert abs(seq["success_rate"] + seq["collision_rate"] + seq["timeout_rate"] - 1.0) < 1e-12


def test_single_env_defaults_to_test_phase_and_val_is_rejected():
    from crowdnav_prediction_attngraph_b200 import _capi
    from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
    with pytest.raises(RuntimeError):
        CudaCrowdVecEnv(device="cuda:0", cfg=_capi.default_config_dict(num_envs=2, phase=1))
