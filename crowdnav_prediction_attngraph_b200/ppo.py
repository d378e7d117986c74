"""PPO update with the reference's entry point (rl/ppo/ppo.py:8-101), in PyTorch on device.

Multi-GPU: environments are sharded across ranks (one process per GPU); the only collectives are
  * one all-reduce of (sum, sum of squares, count) for the GLOBAL advantage normalisation
    (ppo.py:37-39; torch.std is the unbiased estimator), and
  * one all-reduce of the flat gradient per optimiser step, averaged BEFORE clip_grad_norm_
    (ppo.py:85) so clipping sees the global gradient.
With world_size == 1 no collective is issued and the arithmetic is the reference's."""
import torch
import torch.nn as nn
import torch.optim as optim


def _dist_ready():
    return torch.distributed.is_available() and torch.distributed.is_initialized() and \
        torch.distributed.get_world_size() > 1


def global_advantage_normalize(adv):
    """(adv - mean) / (std + 1e-5) with mean/std over ALL ranks' samples (unbiased std)."""
    if not _dist_ready():
        return (adv - adv.mean()) / (adv.std() + 1e-5)
    a64 = adv.double()
    stats = torch.stack([a64.sum(), (a64 * a64).sum(), torch.tensor(float(adv.numel()), device=adv.device, dtype=torch.float64)])
    torch.distributed.all_reduce(stats)
    s, ss, n = stats[0], stats[1], stats[2]
    mean = s / n
    var = (ss - n * mean * mean) / (n - 1)
    return ((adv - mean.float()) / (var.clamp_min(0).sqrt().float() + 1e-5))


def allreduce_gradients(params, events=None):
    """Average gradients across ranks through one flat fp32 bucket (2.5 M floats = 10 MB).
    `events`: optional list that receives (start, end) CUDA events around the collective (bench timing)."""
    if not _dist_ready():
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    if events is not None and flat.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    torch.distributed.all_reduce(flat)
    if events is not None and flat.is_cuda:
        e1.record()
        events.append((e0, e1))
    flat.div_(torch.distributed.get_world_size())
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel() * flat.element_size()


class PPO(object):
    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef,
                 lr=None, eps=None, max_grad_norm=None, use_clipped_value_loss=True, matmul_precision=None):
        """Same constructor as rl/ppo/ppo.py:8-34.  matmul_precision (extension): None keeps PyTorch's setting
        (fp32 matmuls like the reference); 'tf32' runs the update's matmuls on TF32 tensor cores
        (torch.set_float32_matmul_precision('high') for the duration of update())."""
        self.matmul_precision = matmul_precision
        self.profile = False          # True: time the gradient all-reduces of update() -> self.last_profile
        self.last_profile = None
        self.actor_critic = actor_critic
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.max_grad_norm, self.use_clipped_value_loss = max_grad_norm, use_clipped_value_loss
        self.optimizer = optim.Adam(actor_critic.parameters(), lr=lr, eps=eps)
        if _dist_ready():
            # replicas must start from the same weights (train.py seeds every rank alike; this makes it a guarantee)
            for p in actor_critic.parameters():
                torch.distributed.broadcast(p.data, src=0)

    def update(self, rollouts):
        if self.matmul_precision == 'tf32':
            prev = torch.get_float32_matmul_precision()
            torch.set_float32_matmul_precision('high')
            try:
                return self._update(rollouts)
            finally:
                torch.set_float32_matmul_precision(prev)
        return self._update(rollouts)

    def _update(self, rollouts):
        advantages = rollouts.returns[:-1] - rollouts.value_preds[:-1]
        advantages = global_advantage_normalize(advantages)
        v_sum = a_sum = e_sum = 0.0
        params = [p for p in self.actor_critic.parameters()]
        events = [] if self.profile else None
        ar_bytes = 0
        for _ in range(self.ppo_epoch):
            for sample in rollouts.recurrent_generator(advantages, self.num_mini_batch):
                obs_b, hxs_b, act_b, vpred_b, ret_b, masks_b, old_lp_b, adv_b = sample
                values, lp, entropy, _ = self.actor_critic.evaluate_actions(obs_b, hxs_b, masks_b, act_b)
                ratio = torch.exp(lp - old_lp_b)
                surr1 = ratio * adv_b
                surr2 = torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param) * adv_b
                action_loss = -torch.min(surr1, surr2).mean()
                if self.use_clipped_value_loss:
                    vclip = vpred_b + (values - vpred_b).clamp(-self.clip_param, self.clip_param)
                    value_loss = 0.5 * torch.max((values - ret_b).pow(2), (vclip - ret_b).pow(2)).mean()
                else:
                    value_loss = 0.5 * (ret_b - values).pow(2).mean()
                self.optimizer.zero_grad()
                (value_loss * self.value_loss_coef + action_loss - entropy * self.entropy_coef).backward()
                ar_bytes = allreduce_gradients(params, events) or ar_bytes
                nn.utils.clip_grad_norm_(params, self.max_grad_norm)
                self.optimizer.step()
                v_sum += value_loss.item(); a_sum += action_loss.item(); e_sum += entropy.item()
        n = self.ppo_epoch * self.num_mini_batch
        if self.profile:
            if events:
                torch.cuda.synchronize()
            self.last_profile = {"optimizer_steps": n, "allreduce_calls": len(events or []),
                                 "allreduce_ms": float(sum(a.elapsed_time(b) for a, b in (events or []))),
                                 "allreduce_bytes_per_call": int(ar_bytes),
                                 "world_size": torch.distributed.get_world_size() if _dist_ready() else 1}
        return v_sum / n, a_sum / n, e_sum / n
