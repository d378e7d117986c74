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
