"""GPU parity of the CUDA policy forward (through the C ABI) against
  (a) golden outputs of the unmodified reference Policy module, and
  (b) the plain PyTorch fp32 oracle (oracle/policy_ref.py) on fresh inputs.
Tolerance: 1e-4 absolute on the action mean (the policy 'logits', north_star) and on the new
hidden state; value head |v| reaches ~20, checked at 1e-4 absolute as well."""
import numpy as np
import pytest
import torch

from tests.policy_fixture import load_policy_golden, synth_state_dict

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _cuda_policy(N, H, sd, gemm_mode=0):   # fp32 CUDA-core path here; tensor-core path in test_gpu_gemm_tc.py
    from crowdnav_prediction_attngraph_b200.policy import CudaPolicy
    pol = CudaPolicy(N, H, 12, device="cuda:0", gemm_mode=gemm_mode)
    pol.load_state_dict(sd)
    return pol


@pytest.mark.parametrize("name,H", [("policy_h20", 20), ("policy_h50", 50)])
def test_cuda_policy_matches_reference_golden(name, H):
    from oracle.policy_ref import PolicyRef
    g, obs, h, masks = load_policy_golden(name)
    sd = synth_state_dict(PolicyRef(12).state_dict())
    N = h.shape[0]
    pol = _cuda_policy(N, H, sd)
    dobs = {k: v.cuda() for k, v in obs.items()}
    value, action, logp, h1, mean = pol.act(dobs, h.cuda(), masks.cuda(), deterministic=True, return_mean=True)
    np.testing.assert_allclose(mean.cpu().numpy(), g["synth_mean"], rtol=0, atol=TOL)
    np.testing.assert_allclose(value.cpu().numpy(), g["synth_value"], rtol=0, atol=TOL)
    np.testing.assert_allclose(h1.cpu().numpy(), g["synth_h"], rtol=0, atol=TOL)
    assert torch.equal(action, mean)        # deterministic = dist.mode()


def test_cuda_policy_matches_oracle_and_samples_like_torch_normal():
    from oracle.policy_ref import PolicyRef
    from crowdnav_prediction_attngraph_b200.policy import make_reference_like_state_dict
    N, H = 300, 20                          # ragged vs the 128-row GEMM tiles
    sd = make_reference_like_state_dict(12, seed=3)
    sd["dist.logstd._bias"] = torch.tensor([[-0.3], [0.2]])
    ref = PolicyRef(12)
    ref.load_state_dict(sd)
    gen = torch.Generator().manual_seed(0)
    n = torch.randint(1, H + 1, (N, 1), generator=gen).float()
    sp = torch.randn(N, H, 12, generator=gen) * 3
    sp[torch.arange(H)[None, :] >= n] = 15.0
    obs = dict(robot_node=torch.randn(N, 1, 7, generator=gen) * 3, temporal_edges=torch.randn(N, 1, 2, generator=gen),
               spatial_edges=sp, detected_human_num=n)
    h = torch.randn(N, 1, 128, generator=gen)
    masks = (torch.rand(N, 1, generator=gen) > 0.2).float()
    with torch.no_grad():
        rv, rm, rh = ref(obs, h, masks)
    pol = _cuda_policy(N, H, sd)
    noise = torch.randn(N, 2, generator=gen)
    dobs = {k: v.cuda() for k, v in obs.items()}
    value, action, logp, h1, mean = pol.act(dobs, h.cuda(), masks.cuda(), noise=noise.cuda(), return_mean=True)
    scale = max(1.0, float(rv.abs().max()))
    assert (value.cpu() - rv).abs().max() < TOL * scale
    assert (mean.cpu() - rm).abs().max() < TOL
    assert (h1.cpu() - rh).abs().max() < TOL
    std = torch.tensor([-0.3, 0.2]).exp()
    exp_action = noise * std + mean.cpu()
    assert torch.equal(action.cpu(), exp_action)
    dist = torch.distributions.Normal(mean.cpu(), std.expand_as(mean.cpu()))
    assert (logp.cpu() - dist.log_prob(action.cpu()).sum(-1, keepdim=True)).abs().max() < 1e-5


def test_policy_module_drop_in_act_and_evaluate():
    """Policy (nn.Module mirror): act through CUDA == evaluate_actions' torch path on the same step."""
    from crowdnav_prediction_attngraph_b200.policy import Policy
    from crowdnav_prediction_attngraph_b200.vec_env import Box

    class A(object):
        num_processes, seq_length, num_mini_batch = 16, 1, 1
    N, H = 16, 20
    spaces = dict(spatial_edges=Box((H, 12)), robot_node=Box((1, 7)), temporal_edges=Box((1, 2)),
                  detected_human_num=Box((1,)))
    torch.manual_seed(1)
    pol = Policy(spaces, Box((2,)), base_kwargs=A(), base='selfAttn_merge_srnn').cuda()
    gen = torch.Generator().manual_seed(0)
    n = torch.randint(1, H + 1, (N, 1), generator=gen).float()
    obs = dict(robot_node=torch.randn(N, 1, 7, generator=gen).cuda(), temporal_edges=torch.randn(N, 1, 2, generator=gen).cuda(),
               spatial_edges=torch.randn(N, H, 12, generator=gen).cuda(), detected_human_num=n.cuda())
    hx = {'human_node_rnn': torch.randn(N, 1, 128, generator=gen).cuda(),
          'human_human_edge_rnn': torch.zeros(N, H + 1, 256).cuda()}
    masks = torch.ones(N, 1).cuda()
    with torch.no_grad():
        value, action, logp, hx2 = pol.act(obs, hx, masks)
        v2, lp2, ent, _ = pol.evaluate_actions(obs, hx, masks, action)
    assert (value - v2).abs().max() < 1e-4 and (logp - lp2).abs().max() < 1e-4
    assert hx2['human_human_edge_rnn'].shape == (N, H + 1, 256)


def test_cuda_policy_varnum_input_size_2_both_gemm_modes():
    """CrowdSimVarNum-v0 policy (BASELINE config 1 shape: 5 humans, spatial_edges width 2)."""
    from oracle.policy_ref import PolicyRef
    from crowdnav_prediction_attngraph_b200.policy import CudaPolicy, make_reference_like_state_dict
    N, H = 70, 5
    sd = make_reference_like_state_dict(2, seed=9)
    ref = PolicyRef(2)
    ref.load_state_dict(sd)
    gen = torch.Generator().manual_seed(1)
    n = torch.randint(1, H + 1, (N, 1), generator=gen).float()
    sp = torch.randn(N, H, 2, generator=gen) * 3
    sp[torch.arange(H)[None, :] >= n] = 15.0
    obs = dict(robot_node=torch.randn(N, 1, 7, generator=gen) * 3, temporal_edges=torch.randn(N, 1, 2, generator=gen),
               spatial_edges=sp, detected_human_num=n)
    h = torch.randn(N, 1, 128, generator=gen)
    masks = torch.ones(N, 1)
    with torch.no_grad():
        rv, rm, rh = ref(obs, h, masks)
    for mode in (0, 1):
        pol = CudaPolicy(N, H, 2, device="cuda:0", gemm_mode=mode)
        pol.load_state_dict(sd)
        value, action, logp, h1, mean = pol.act({k: v.cuda() for k, v in obs.items()}, h.cuda(), masks.cuda(),
                                                deterministic=True, return_mean=True)
        assert (value.cpu() - rv).abs().max() < TOL and (mean.cpu() - rm).abs().max() < TOL and (h1.cpu() - rh).abs().max() < TOL


@pytest.mark.parametrize("H,seed,fused", [(20, 11, "0"), (50, 12, "0"), (20, 13, "1"), (50, 14, "1")])
def test_benchmarked_tensor_core_path_at_full_size_matches_oracle(H, seed, fused, monkeypatch):
    """The configuration bench.py times: gemm_mode = 1 (tcgen05 3xFP16), N = 4096 environments, device-side row
    compaction with random detected_human_num (ragged rows, many 128-row tiles per CTA: persistent tile loop, both TMEM
    accumulators in flight), 3 consecutive calls (the double-buffered outputs and the hidden-state feedback) -- against
    the plain PyTorch fp32 oracle (oracle/policy_ref.py), action mean / hidden state 1e-4, value 1e-4 of its scale.
    fused = "1": the opt-in single-kernel QKV projection + human-human attention (cn_qkv_attn.cuh, CN_FUSE_QKV=1)."""
    from oracle.policy_ref import PolicyRef
    monkeypatch.setenv("CN_FUSE_QKV", fused)          # read by cn_policy_create
    from crowdnav_prediction_attngraph_b200.policy import make_reference_like_state_dict
    N = 4096
    sd = make_reference_like_state_dict(12, seed=seed)
    ref = PolicyRef(12)
    ref.load_state_dict(sd)
    pol = _cuda_policy(N, H, sd, gemm_mode=1)
    gen = torch.Generator().manual_seed(seed)
    h = torch.randn(N, 1, 128, generator=gen) * 0.5
    for it in range(3):
        n = torch.randint(1, H + 1, (N, 1), generator=gen).float()
        if it == 1:
            n[: N // 2] = 1.0                 # half of the batch sees a single human: many tiny segments
        sp = torch.randn(N, H, 12, generator=gen) * 3
        sp[torch.arange(H)[None, :] >= n] = 15.0
        obs = dict(robot_node=torch.randn(N, 1, 7, generator=gen) * 3, temporal_edges=torch.randn(N, 1, 2, generator=gen),
                   spatial_edges=sp, detected_human_num=n)
        masks = (torch.rand(N, 1, generator=gen) > 0.1).float()
        with torch.no_grad():
            rv, rm, rh = ref(obs, h, masks)
        dobs = {k: v.cuda() for k, v in obs.items()}
        value, action, logp, h1, mean = pol.act(dobs, h.cuda(), masks.cuda(), deterministic=True, return_mean=True)
        assert int(pol.lib.cn_policy_last_rows(pol._h)) == int(n.sum())
        scale = max(1.0, float(rv.abs().max()))
        assert float((value.cpu() - rv).abs().max()) < TOL * scale, it
        assert float((mean.cpu() - rm).abs().max()) < TOL, it
        assert float((h1.cpu() - rh).abs().max()) < TOL, it
        h = rh                                 # oracle's state feeds both (no error accumulation across calls)
