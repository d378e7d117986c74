"""GPU: the zero-copy device-resident rollout (kernels write straight into the storage slots) fills the
rollout storage exactly like the reference-contract loop act -> envs.step -> rollouts.insert."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(N, H, T):
    from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
    from crowdnav_prediction_attngraph_b200.policy import CudaPolicy, make_reference_like_state_dict
    from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
    env = CudaCrowdVecEnv(num_envs=N, human_num=H, seed=77, device="cuda:0")
    pol = CudaPolicy(N, H, 12, device="cuda:0")
    pol.load_state_dict(make_reference_like_state_dict(12, seed=5))
    ro = RolloutStorage(T, N, env.observation_space.spaces, env.action_space, 128, 256, device="cuda:0")
    obs = env.reset()
    for k in ro.obs:
        ro.obs[k][0].copy_(obs[k])
    return env, pol, ro


def test_zero_copy_rollout_equals_insert_loop():
    N, H, T = 64, 20, 12
    env_a, pol_a, ro_a = _make(N, H, T)
    env_b, pol_b, ro_b = _make(N, H, T)
    torch.manual_seed(3)
    for _ in range(T):
        s = ro_a.step
        o = {k: ro_a.obs[k][s] for k in ro_a.obs}
        v, a, lp, h = pol_a.act(o, ro_a.recurrent_hidden_states['human_node_rnn'][s], ro_a.masks[s])
        nobs, rew, done, info = env_a.step_device(a)
        ro_a.insert(nobs, {'human_node_rnn': h}, a, lp, v, rew, (1.0 - done.float()).unsqueeze(1))
    torch.manual_seed(3)
    for _ in range(T):
        ro_b.rollout_step_zero_copy(pol_b, env_b)
    torch.cuda.synchronize()
    for k in ro_a.obs:
        assert torch.equal(ro_a.obs[k], ro_b.obs[k]), k
    for name in ("rewards", "value_preds", "action_log_probs", "actions", "masks"):
        assert torch.equal(getattr(ro_a, name), getattr(ro_b, name)), name
    assert torch.equal(ro_a.recurrent_hidden_states['human_node_rnn'], ro_b.recurrent_hidden_states['human_node_rnn'])
