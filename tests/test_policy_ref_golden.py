"""Pins oracle/policy_ref.py against outputs of the UNMODIFIED reference Policy module
(tools/make_golden_policy.py).  fp32 torch on both sides -> tight tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle.policy_ref import PolicyRef
from tests.policy_fixture import load_policy_golden, synth_state_dict

CKPT = "/root/reference/trained_models/GST_predictor_rand/checkpoints/41665.pt"


@pytest.mark.parametrize("name", ["policy_h20", "policy_h50"])
def test_policy_ref_matches_reference_synthetic_weights(name):
    g, obs, h, masks = load_policy_golden(name)
    ref = PolicyRef(12)
    ref.load_state_dict(synth_state_dict(ref.state_dict()))
    with torch.no_grad():
        v, m, h1 = ref(obs, h, masks)
    np.testing.assert_allclose(v.numpy(), g["synth_value"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(m.numpy(), g["synth_mean"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(h1.numpy(), g["synth_h"], rtol=0, atol=2e-5)


@pytest.mark.skipif(not os.path.exists(CKPT), reason="shipped checkpoint only exists in the build container")
@pytest.mark.parametrize("name", ["policy_h20", "policy_h50"])
def test_policy_ref_matches_reference_shipped_checkpoint(name):
    g, obs, h, masks = load_policy_golden(name)
    ref = PolicyRef(12)
    missing = ref.load_state_dict(torch.load(CKPT, map_location="cpu"))
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        v, m, h1 = ref(obs, h, masks)
    np.testing.assert_allclose(v.numpy(), g["ckpt_value"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(m.numpy(), g["ckpt_mean"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(h1.numpy(), g["ckpt_h"], rtol=0, atol=5e-5)


def test_update_path_packed_rows_equal_padded():
    """evaluate_actions on the valid (compacted) human rows == on all padded rows: values, log-probs, gradients."""
    from crowdnav_prediction_attngraph_b200.vec_env import Box
    from crowdnav_prediction_attngraph_b200.policy import Policy

    class Args(object):
        num_processes, seq_length, num_mini_batch = 6, 4, 2
    H = 20
    spaces = {'robot_node': Box((1, 7)), 'temporal_edges': Box((1, 2)), 'spatial_edges': Box((H, 12)),
              'detected_human_num': Box((1,))}
    torch.manual_seed(0)
    pol = Policy(spaces, Box((2,)), base_kwargs=Args(), base='selfAttn_merge_srnn')
    T, N = 4, 6
    inp = {'robot_node': torch.randn(T * N, 1, 7), 'temporal_edges': torch.randn(T * N, 1, 2),
           'spatial_edges': torch.randn(T * N, H, 12), 'detected_human_num': torch.randint(1, 9, (T * N, 1)).float()}
    h0 = {'human_node_rnn': torch.randn(N, 1, 128)}
    masks, act = torch.ones(T * N, 1), torch.randn(T * N, 2)
    res = {}
    for flag in (True, False):
        pol.pack_valid_rows = flag
        pol.zero_grad()
        v, lp, ent, _ = pol.evaluate_actions(inp, h0, masks, act)
        (v.sum() + lp.sum() + ent).backward()
        res[flag] = (v.detach().clone(), lp.detach().clone(),
                     torch.cat([p.grad.flatten() for p in pol.parameters() if p.grad is not None]).clone())
    assert torch.allclose(res[True][0], res[False][0], atol=1e-6) and torch.allclose(res[True][1], res[False][1], atol=1e-6)
    assert torch.allclose(res[True][2], res[False][2], atol=1e-5 * float(res[False][2].abs().max()))
