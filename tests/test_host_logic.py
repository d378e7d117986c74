"""CPU: host-side logic of the mirrors that needs no GPU -- config snapshot rules, the evaluation summary, the rollout
storage (insert / GAE / minibatch generator) against straightforward restatements of the reference formulas."""
import types

import numpy as np
import pytest
import torch

from crowdnav_prediction_attngraph_b200 import _capi
from crowdnav_prediction_attngraph_b200.evaluation import _summary
from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
from crowdnav_prediction_attngraph_b200.vec_env import Box, LazyInfos, Danger, Collision, config_dict_from_reference


def _ref_config(**over):
    ns = types.SimpleNamespace
    c = ns(sim=ns(human_num=20, human_num_range=0, predict_steps=5, predict_method="const_vel", circle_radius=6 * 2 ** 0.5,
                  arena_size=6),
           action_space=ns(kinematics="holonomic"), humans=ns(policy="orca", radius=0.3, v_pref=1, FOV=2.0,
                                                             random_goal_changing=False, goal_change_chance=0.5,
                                                             end_goal_changing=True),
           robot=ns(visible=False, radius=0.3, v_pref=1, FOV=2, sensor_range=5),
           env=ns(randomize_attributes=False, time_step=0.25, time_limit=50, val_size=100, test_size=500),
           data=ns(pred_timestep=0.25),
           reward=ns(discomfort_dist=0.25, discomfort_penalty_factor=10, success_reward=10, collision_penalty=-20),
           orca=ns(neighbor_dist=10, safety_space=0.15, time_horizon=5), sf=ns(A=2.0, B=1.0, KI=1.0),
           args=ns(sort_humans=True))
    for k, v in over.items():
        setattr(c.sim, k, v)
    return c


def test_config_snapshot_follows_the_reference_phase_rule_and_scope():
    c = _ref_config()
    d = config_dict_from_reference(c, 16, 425, "CrowdSimPred-v0")
    assert d["phase"] == 0 and d["const_vel"] == 1 and d["test_size"] == 500 and d["nenv_total"] == 16
    # rl/networks/envs.py:55-58: a single environment runs in phase 'test'
    assert config_dict_from_reference(c, 1, 425, "CrowdSimPred-v0")["phase"] == 2
    assert config_dict_from_reference(c, 1, 425, "CrowdSimVarNum-v0")["const_vel"] == 0
    assert config_dict_from_reference(c, 8, 425, "CrowdSimPred-v0", phase="test")["phase"] == 2
    with pytest.raises(NotImplementedError):
        config_dict_from_reference(c, 8, 425, "CrowdSimPred-v0", phase="val")
    # sim.human_num_range > 0 and social-force humans are engine features since round 2
    d2 = config_dict_from_reference(_ref_config(human_num_range=2), 8, 425, "CrowdSimPred-v0")
    assert d2["human_num_range"] == 2 and d2["human_num"] == 20
    csf = _ref_config()
    csf.humans.policy = "social_force"
    assert config_dict_from_reference(csf, 8, 425, "CrowdSimPred-v0")["human_policy"] == 1
    cuni = _ref_config()
    cuni.action_space.kinematics = "unicycle"
    with pytest.raises(NotImplementedError):
        config_dict_from_reference(cuni, 8, 425, "CrowdSimPred-v0")
    cus = _ref_config()
    cus.args.sort_humans = False
    with pytest.raises(NotImplementedError):
        config_dict_from_reference(cus, 8, 425, "CrowdSimPred-v0")
    with pytest.raises(NotImplementedError):
        config_dict_from_reference(c, 8, 425, "rosTurtlebot2iEnv-v0")
    # the flat dict maps 1:1 onto the C struct
    cfg = _capi.config_from_dict(d)
    assert cfg.human_num == 20 and cfg.phase == 0 and abs(cfg.circle_radius - 6 * 2 ** 0.5) < 1e-12


def test_lazy_infos_materialise_like_the_reference_dicts():
    infos = LazyInfos(np.array([0, 4, 2]), np.array([0.0, 0.41, 0.0], np.float32), np.array([False, False, True]),
                      np.array([0.0, 0.0, -3.25]), np.array([0, 0, 17]))
    assert len(infos) == 3 and 'episode' not in infos[0]
    assert isinstance(infos[1]['info'], Danger) and abs(infos[1]['info'].min_dist - 0.41) < 1e-6
    assert isinstance(infos[2]['info'], Collision) and infos[2]['episode'] == {'r': -3.25, 'l': 17, 't': 0.0}


def test_evaluation_summary_matches_rl_evaluation_bookkeeping():
    # 3 = ReachGoal, 2 = Collision, 1 = Timeout; nav time averages the successes only, time-outs count time_limit
    out = _summary(4, 50.0, [3, 2, 1, 3], [10.0, 4.0, 50.0, 12.0], [20.0, 5.0, 30.0, 22.0], [0.0, 10.0, 0.0, 5.0],
                   [0.4, 0.5], [1.0, -2.0, 0.0, 3.0])
    assert out["success_rate"] == 0.5 and out["collision_rate"] == 0.25 and out["timeout_rate"] == 0.25
    assert out["avg_nav_time"] == 11.0 and out["collision_cases"] == [1] and out["timeout_cases"] == [2]
    assert out["path_length"] == pytest.approx(19.25) and out["intrusion_ratio"] == pytest.approx(3.75)
    assert out["min_intrusion_dist"] == pytest.approx(0.45)
    assert _summary(1, 50.0, [2], [3.0], [1.0], [0.0], [], [0.0])["avg_nav_time"] == 50.0      # no success: time_limit


def _storage(T=5, N=4, H=3):
    spaces = {'robot_node': Box((1, 7)), 'temporal_edges': Box((1, 2)), 'spatial_edges': Box((H, 12)),
              'detected_human_num': Box((1,))}
    return RolloutStorage(T, N, spaces, Box((2,)), 128, 256, device="cpu"), spaces


def test_rollout_storage_insert_gae_and_generator_on_cpu():
    torch.manual_seed(0)
    T, N, H = 5, 4, 3
    ro, spaces = _storage(T, N, H)
    for t in range(T):
        obs = {k: torch.randn(N, *spaces[k].shape) for k in spaces}
        ro.insert(obs, {'human_node_rnn': torch.randn(N, 1, 128)}, torch.randn(N, 2), torch.randn(N, 1), torch.randn(N, 1),
                  torch.randn(N, 1), (torch.rand(N, 1) > 0.3).float(), torch.ones(N, 1))
        assert torch.equal(ro.obs['spatial_edges'][t + 1], obs['spatial_edges'])
    assert ro.step == 0
    nv = torch.randn(N, 1)
    ro.compute_returns(nv, True, 0.99, 0.95, False)
    # straightforward restatement of rl/networks/storage.py:88-104 (GAE)
    vp = torch.cat([ro.value_preds[:-1], nv.unsqueeze(0)], 0).numpy()
    rew, m = ro.rewards.numpy(), ro.masks.numpy()
    gae = np.zeros((N, 1), np.float32)
    for t in reversed(range(T)):
        delta = rew[t] + 0.99 * vp[t + 1] * m[t + 1] - vp[t]
        gae = delta + 0.99 * 0.95 * m[t + 1] * gae
        np.testing.assert_allclose(ro.returns[t].numpy(), gae + vp[t], rtol=1e-5, atol=1e-6)
    adv = ro.returns[:-1] - ro.value_preds[:-1]
    seen = 0
    for obs_b, hxs_b, act_b, vpred_b, ret_b, masks_b, old_lp_b, adv_b in ro.recurrent_generator(adv, 2):
        n_b = hxs_b['human_node_rnn'].shape[0]
        assert n_b == N // 2 and act_b.shape == (T * n_b, 2) and obs_b['spatial_edges'].shape == (T * n_b, H, 12)
        assert masks_b.shape == (T * n_b, 1) and adv_b.shape == (T * n_b, 1)
        seen += n_b
    assert seen == N
    ro.after_update()
    assert torch.equal(ro.obs['robot_node'][0], ro.obs['robot_node'][-1])
