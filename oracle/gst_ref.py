"""TEST INFRASTRUCTURE (oracle): CPU restatement of BASELINE config 3's trajectory predictor and wrapper (row a16).

  gst_forward            = CrowdNavPredInterfaceMultiEnv.forward (gst_updated/scripts/wrapper/crowd_nav_interface_parallel.py:45-114)
                           + st_model.forward (gst_updated/src/gumbel_social_transformer/st_model.py:271-455)
                           for the shipped predictor's configuration (checkpoint/args.pickle): full connectivity
                           (spatial_num_heads_edges = 0, no ghost node), one pre-LN node-encoder layer with 8 heads
                           (node_encoder_layer_no_ghost.py:24-66, mha.py:236-242: soft-max over ALL keys, then multiply
                           by the float mask and renormalise with +1e-10), 'faster_lstm' over the 5 observed frames,
                           recursive decoding of 5 steps, sampling=False (the mean is fed back).
  PretextWrapperRef      = VecPretextNormalize.reset / process_obs_rew (rl/vec_env/vec_pretext_normalize.py:85-191).

Pinned against the unmodified reference by tests/test_gst_oracle_golden.py (fixtures from tools/make_golden_gst.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

INVALID = -999.0


def load_params(path):
    g = np.load(path)
    return {k: torch.tensor(g[k]) for k in g.files}


def _encoder_layer(p, x, attn_mask):
    """x: [B, H, 2] displacements, attn_mask: [B, H(target), H(neighbour)] float.  Returns [B, H, 64]."""
    pre = "gumbel_social_transformer.node_encoder_layers.0."
    B, H, _ = x.shape
    x = F.linear(x, p["gumbel_social_transformer.node_embedding.weight"], p["gumbel_social_transformer.node_embedding.bias"])
    ped = (attn_mask.sum(-1) > 0).float().unsqueeze(-1)
    x = F.layer_norm(x, (64,), p[pre + "norm_node.weight"], p[pre + "norm_node.bias"])
    x = x * ped
    qkv = F.linear(x, p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    q = q * (8.0 ** -0.5)
    q = q.reshape(B, H, 8, 8).permute(0, 2, 1, 3)          # [B, head, H, 8]
    k = k.reshape(B, H, 8, 8).permute(0, 2, 1, 3)
    v = v.reshape(B, H, 8, 8).permute(0, 2, 1, 3)
    w = torch.softmax(q @ k.transpose(-1, -2), dim=-1)       # over ALL neighbours
    w = w * attn_mask.unsqueeze(1)
    w = w / (w.sum(-1, keepdim=True) + 1e-10)
    a = (w @ v).permute(0, 2, 1, 3).reshape(B, H, 64)
    x2 = F.linear(a, p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"])
    x = x + x2
    x2 = F.layer_norm(x, (64,), p[pre + "norm1_node.weight"], p[pre + "norm1_node.bias"])
    x2 = F.linear(F.relu(F.linear(x2, p[pre + "linear1.weight"], p[pre + "linear1.bias"])), p[pre + "linear2.weight"],
                  p[pre + "linear2.bias"])
    return x + x2


def _lstm_cell(p, x, h, c):
    g = F.linear(x, p["lstm.weight_ih_l0"], p["lstm.bias_ih_l0"]) + F.linear(h, p["lstm.weight_hh_l0"], p["lstm.bias_hh_l0"])
    i, f, gg, o = g.chunk(4, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def gst_forward(p, in_traj, in_mask):
    """in_traj [N,H,5,2] (world positions, INVALID where unseen), in_mask [N,H,5,1] (0/1).
    Returns out_traj [N,H,5,5] = (mu_x, mu_y, sigma_x, sigma_y, corr) accumulated over the prediction steps
    (positions in the world frame, INVALID where out_mask is 0) and out_mask [N,H,1]."""
    in_traj = torch.as_tensor(in_traj, dtype=torch.float32)
    m = torch.as_tensor(in_mask, dtype=torch.float32)[..., 0]                 # [N,H,5]
    N, H, T = m.shape
    mrel = torch.cat((m[:, :, :1], m[:, :, :-1] * m[:, :, -1:]), dim=2)       # interface.forward: sic
    pos = in_traj                                                             # [N,H,5,2]
    rel = torch.cat((torch.zeros(N, H, 1, 2), pos[:, :, 1:] - pos[:, :, :-1]), dim=2)
    rel = INVALID * (1 - mrel.unsqueeze(-1)) + rel * mrel.unsqueeze(-1)
    x = rel.permute(0, 2, 1, 3).reshape(N * T, H, 2)                          # [N*5, H, 2]
    mt = mrel.permute(0, 2, 1).reshape(N * T, H)
    attn = mt.unsqueeze(2) * mt.unsqueeze(1)                                  # [N*5, H, H]
    xs = _encoder_layer(p, x, attn).reshape(N, T, H, 64)
    xs = xs * mrel.permute(0, 2, 1).unsqueeze(-1)
    h = torch.zeros(N * H, 64)
    c = torch.zeros(N * H, 64)
    for t in range(T):
        h, c = _lstm_cell(p, xs[:, t].reshape(N * H, 64), h, c)
    fp = mrel[:, :, T - 1]                                                    # loss_mask_rel_full_partial [N,H]
    fpf = fp.reshape(-1, 1)
    h = h * fpf
    c = c * fpf
    attn_pred = fp.unsqueeze(2) * fp.unsqueeze(1)                             # [N,H,H]
    mus, sxs, sys_, corrs = [], [], [], []
    for tt in range(5):
        if tt > 0:
            xt = _encoder_layer(p, x_sample, attn_pred).reshape(N * H, 64) * fpf
            hp, cp = _lstm_cell(p, xt, h, c)
            h = hp * fpf + h * (1 - fpf)
            c = cp * fpf + c * (1 - fpf)
        raw = F.linear(h, p["hidden2pos.weight"], p["hidden2pos.bias"]).reshape(N, H, 5)
        mu, sx, sy, corr = raw[..., :2], torch.exp(raw[..., 2:3]), torch.exp(raw[..., 3:4]), torch.tanh(raw[..., 4:5])
        x_sample = mu * fp.unsqueeze(-1)
        mus.append(mu); sxs.append(sx); sys_.append(sy); corrs.append(corr)
    mu = torch.stack(mus, 2).cumsum(2)                                        # [N,H,5,2]
    sx, sy, corr = torch.stack(sxs, 2), torch.stack(sys_, 2), torch.stack(corrs, 2)
    sx2, sy2, cxy = (sx ** 2.).cumsum(2), (sy ** 2.).cumsum(2), (corr * sx * sy).cumsum(2)
    sxc, syc = sx2 ** 0.5, sy2 ** 0.5
    mu = mu + pos[:, :, T - 1:T]
    pm = fp.reshape(N, H, 1, 1)
    mu = mu * pm + INVALID * (1 - pm)
    out = torch.cat((mu, sxc, syc, cxy / (sxc * syc)), dim=3)
    return out, fp.reshape(N, H, 1)


class PretextWrapperRef(object):
    """VecPretextNormalize's observation / reward processing around N RealGST environments."""

    def __init__(self, params, num_envs, human_num, predict_steps=5, robot_radius=0.3, human_radius=0.3,
                 collision_penalty=-20.0):
        self.p, self.N, self.H, self.P = params, num_envs, human_num, predict_steps
        self.thr = robot_radius + human_radius
        self.collision_penalty = collision_penalty
        self.reset()

    def reset(self):
        self.traj = [torch.full((self.N, self.H, 2), INVALID) for _ in range(5)]
        self.mask = [torch.zeros(self.N, self.H, 1) for _ in range(5)]

    def process(self, O, rews=None):
        """O: raw RealGST observation dict of arrays ([N,1,7], [N,1,2], [N,H,12], [N,1], [N,H] bool).
        Returns (obs dict with the predicted, distance-sorted spatial_edges, rews + future-collision penalty)."""
        N, H, P = self.N, self.H, self.P
        robot = torch.as_tensor(O["robot_node"], dtype=torch.float32).reshape(N, 1, 7)
        sp = torch.as_tensor(O["spatial_edges"], dtype=torch.float32).clone()
        vis = torch.as_tensor(O["visible_masks"]).reshape(N, H)
        human_pos = robot[:, :, :2] + sp[:, :, :2]
        self.traj = self.traj[1:] + [human_pos]
        self.mask = self.mask[1:] + [vis.unsqueeze(-1).float()]
        in_traj = torch.stack(self.traj).permute(1, 2, 0, 3)
        in_mask = torch.stack(self.mask).permute(1, 2, 0, 3)
        out_traj, out_mask = gst_forward(self.p, in_traj, in_mask)
        out_mask = out_mask.bool()
        d = out_traj[:, :, :, :2] - robot[:, :, :2].unsqueeze(1)
        coll = (torch.norm(d, dim=-1) < self.thr) & out_mask
        coef = 2. ** torch.arange(2, P + 2).reshape(1, 1, P)
        pen, _ = torch.min((coll.float() * (self.collision_penalty / coef)).reshape(N, -1), dim=1)
        if rews is not None:
            rews = np.asarray(rews, dtype=np.float64).reshape(N, 1) + pen.reshape(N, 1).numpy()
        rel = (out_traj[:, :, :, :2] - robot[:, :, :2].unsqueeze(1)).reshape(N, H, -1)
        om = out_mask.repeat(1, 1, 2 * P)
        sp[:, :, 2:][om] = rel[om]
        order = torch.argsort(torch.linalg.norm(sp[:, :, :2], dim=-1), dim=1, stable=True)
        sp = torch.stack([sp[i][order[i]] for i in range(N)])
        obs = dict(robot_node=robot.numpy(), spatial_edges=sp.numpy(),
                   temporal_edges=np.asarray(O["temporal_edges"], dtype=np.float32).reshape(N, 1, 2),
                   visible_masks=vis.numpy(), detected_human_num=np.asarray(O["detected_human_num"], dtype=np.float32).reshape(N, 1))
        return obs, rews, pen.numpy()
