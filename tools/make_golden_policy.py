#!/usr/bin/env python
"""Golden vectors for the policy forward, generated from the UNMODIFIED reference module
(rl.networks.model.Policy, base selfAttn_merge_srnn) in the build container.

Weights are a deterministic synthetic fill (param_fill below) whose per-tensor scale follows the
shipped checkpoint trained_models/GST_predictor_rand/checkpoints/41665.pt, so fixtures stay
small (no 10 MB weight file in git); inputs are observations recorded in tests/golden/env_*.npz.
A second fixture stores the outputs of the shipped checkpoint itself on the same inputs
(only compared in this container, where the checkpoint exists).
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle", "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

CKPT = "/root/reference/trained_models/GST_predictor_rand/checkpoints/41665.pt"


def param_fill(state_dict, seed, scales):
    """Deterministic fill: keys in sorted order, torch.manual_seed(seed + index), N(0,1) * scale."""
    out = {}
    for i, k in enumerate(sorted(state_dict.keys())):
        g = torch.Generator().manual_seed(seed + i)
        out[k] = torch.randn(state_dict[k].shape, generator=g) * float(scales[k])
    return out


def build_reference_policy(env_name, H, W, nenv):
    sys.argv = ["x", "--no-cuda", "--env-name", env_name, "--num-processes", str(nenv)]
    import gym
    from arguments import get_args
    from rl.networks.model import Policy
    args = get_args()
    obs_space = {"robot_node": gym.spaces.Box(-np.inf, np.inf, (1, 7)),
                 "temporal_edges": gym.spaces.Box(-np.inf, np.inf, (1, 2)),
                 "spatial_edges": gym.spaces.Box(-np.inf, np.inf, (H, W)),
                 "detected_human_num": gym.spaces.Box(-np.inf, np.inf, (1,))}
    act_space = gym.spaces.Box(-np.inf * np.ones(2), np.inf * np.ones(2), dtype=np.float32)
    return Policy(obs_space, act_space, base_kwargs=args, base="selfAttn_merge_srnn")


def main():
    sd_ck = torch.load(CKPT, map_location="cpu")
    scales = {k: float(v.float().std()) if v.numel() > 1 else 1.0 for k, v in sd_ck.items()}
    scales = {k: (s if s > 0 else 0.05) for k, s in scales.items()}
    # output heads boosted so the synthetic policy reaches the checkpoint's output magnitudes
    # (|value| ~ 20, |mean| ~ 10): keeps the absolute 1e-4 tolerance test meaningful
    scales["base.critic_linear.weight"] *= 12.0
    scales["dist.fc_mean.weight"] *= 8.0
    np.savez(os.path.join(REPO, "tests", "golden", "policy_param_scales.npz"),
             keys=np.array(sorted(scales.keys())), scales=np.array([scales[k] for k in sorted(scales.keys())]),
             shapes=np.array([str(tuple(sd_ck[k].shape)) for k in sorted(scales.keys())]))
    for name, env_file, H, W in [("policy_h20", "env_pred_h20", 20, 12), ("policy_h50", "env_pred_h50_rand", 50, 12)]:
        g = np.load(os.path.join(REPO, "tests", "golden", env_file + ".npz"))
        T, N = g["actions"].shape[:2]
        B = 64
        idx = np.random.RandomState(0).choice((T + 1) * N, B, replace=False)
        take = lambda k: torch.from_numpy(g["ob_" + k].reshape((T + 1) * N, *g["ob_" + k].shape[2:])[idx].astype(np.float32))
        obs = {k: take(k) for k in ["robot_node", "temporal_edges", "spatial_edges", "detected_human_num"]}
        gen = torch.Generator().manual_seed(123)
        h = torch.randn(B, 1, 128, generator=gen) * 0.5
        masks = (torch.rand(B, 1, generator=gen) > 0.1).float()
        pol = build_reference_policy("CrowdSimPred-v0", H, W, B)
        out = {}
        for tag, sd in [("synth", param_fill(sd_ck, 1000, scales)), ("ckpt", sd_ck)]:
            pol.load_state_dict(sd)
            rnn = {"human_node_rnn": h.clone(), "human_human_edge_rnn": torch.zeros(B, H + 1, 256)}
            with torch.no_grad():
                value, feat, hx = pol.base({k: v.clone() for k, v in obs.items()}, rnn, masks.clone(), infer=True)
                mean = pol.dist.fc_mean(feat)
            out[tag + "_value"] = value.numpy()
            out[tag + "_mean"] = mean.numpy()
            out[tag + "_h"] = hx["human_node_rnn"].numpy()
            print(name, tag, "value range", float(value.min()), float(value.max()), "mean abs max", float(mean.abs().max()))
        np.savez_compressed(os.path.join(REPO, "tests", "golden", name + ".npz"),
                            h=h.numpy(), masks=masks.numpy(), **{"ob_" + k: v.numpy() for k, v in obs.items()}, **out)


if __name__ == "__main__":
    main()
