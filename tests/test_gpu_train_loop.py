"""GPU: the reference's training loop shape (train.py:144-210) runs on the mirrors end to end —
rollout through the VecEnv contract, GAE, PPO update in PyTorch, parameters re-uploaded to the
CUDA policy engine after every optimiser step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Args(object):
    num_processes, seq_length, num_mini_batch, num_steps = 32, 8, 2, 8
    clip_param, ppo_epoch, value_loss_coef, entropy_coef = 0.2, 2, 0.5, 0.0
    lr, eps, max_grad_norm, gamma, gae_lambda = 4e-5, 1e-5, 0.5, 0.99, 0.95


def test_train_loop_on_mirrors():
    from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
    from crowdnav_prediction_attngraph_b200.policy import Policy
    from crowdnav_prediction_attngraph_b200.storage import RolloutStorage
    from crowdnav_prediction_attngraph_b200 import ppo
    a = Args()
    dev = torch.device("cuda:0")
    torch.manual_seed(425)
    envs = CudaCrowdVecEnv(num_envs=a.num_processes, human_num=20, seed=425, device=dev)
    actor_critic = Policy(envs.observation_space.spaces, envs.action_space, base_kwargs=a, base='selfAttn_merge_srnn').to(dev)
    rollouts = RolloutStorage(a.num_steps, a.num_processes, envs.observation_space.spaces, envs.action_space, 128, 256, device=dev)
    agent = ppo.PPO(actor_critic, a.clip_param, a.ppo_epoch, a.num_mini_batch, a.value_loss_coef, a.entropy_coef,
                    lr=a.lr, eps=a.eps, max_grad_norm=a.max_grad_norm)
    obs = envs.reset()
    for k in rollouts.obs:
        rollouts.obs[k][0].copy_(obs[k])
    w0 = actor_critic.base.actor[0].weight.detach().clone()
    episode_rewards = []
    for j in range(2):
        for step in range(a.num_steps):
            with torch.no_grad():
                o = {k: rollouts.obs[k][step] for k in rollouts.obs}
                hx = {k: rollouts.recurrent_hidden_states[k][step] for k in rollouts.recurrent_hidden_states}
                value, action, logp, hx2 = actor_critic.act(o, hx, rollouts.masks[step])
            obs, reward, done, infos = envs.step(action)
            for info in infos:
                if 'episode' in info.keys():
                    episode_rewards.append(info['episode']['r'])
            masks = torch.FloatTensor([[0.0] if d else [1.0] for d in done])
            bad_masks = torch.FloatTensor([[0.0] if 'bad_transition' in info.keys() else [1.0] for info in infos])
            rollouts.insert(obs, hx2, action, logp, value, reward, masks, bad_masks)
        with torch.no_grad():
            o = {k: rollouts.obs[k][-1] for k in rollouts.obs}
            hx = {k: rollouts.recurrent_hidden_states[k][-1] for k in rollouts.recurrent_hidden_states}
            next_value = actor_critic.get_value(o, hx, rollouts.masks[-1]).detach()
        rollouts.compute_returns(next_value, True, a.gamma, a.gae_lambda, False)
        v_loss, a_loss, ent = agent.update(rollouts)
        rollouts.after_update()
        assert np.isfinite([v_loss, a_loss, ent]).all()
    assert not torch.equal(w0, actor_critic.base.actor[0].weight.detach())
    # the CUDA engine picked up the updated parameters: act == evaluate on the fresh weights
    with torch.no_grad():
        o = {k: rollouts.obs[k][0] for k in rollouts.obs}
        hx = {k: rollouts.recurrent_hidden_states[k][0] for k in rollouts.recurrent_hidden_states}
        value, action, logp, _ = actor_critic.act(o, hx, rollouts.masks[0])
        class One(object):
            pass
        actor_critic.seq_length = 1
        v2, lp2, _, _ = actor_critic.evaluate_actions(o, hx, rollouts.masks[0], action)
    assert (value - v2).abs().max() < 2e-4 and (logp - lp2).abs().max() < 2e-4
    sd = actor_critic.state_dict()
    assert 'base.human_node_final_linear.weight' in sd and 'dist.logstd._bias' in sd and len(sd) == 47
