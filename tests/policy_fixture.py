"""Shared helpers for the policy parity tests (deterministic synthetic weights + golden loader)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synth_state_dict(template, seed=1000):
    """Same fill as tools/make_golden_policy.py:param_fill (sorted keys, seed + index, N(0,1) * scale)."""
    sc = np.load(os.path.join(GOLD, "policy_param_scales.npz"))
    scales = {str(k): float(s) for k, s in zip(sc["keys"], sc["scales"])}
    keys = sorted(scales.keys())
    out = {}
    for i, k in enumerate(keys):
        if k not in template:
            continue
        g = torch.Generator().manual_seed(seed + i)
        out[k] = torch.randn(tuple(template[k].shape), generator=g) * scales[k]
    assert set(out.keys()) == set(template.keys()), set(template.keys()) ^ set(out.keys())
    return out


def load_policy_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    obs = {k: torch.from_numpy(g["ob_" + k]) for k in ["robot_node", "temporal_edges", "spatial_edges",
                                                       "detected_human_num"]}
    return g, obs, torch.from_numpy(g["h"]), torch.from_numpy(g["masks"])
