"""GPU parity of the PPO-update kernels against plain PyTorch:

  * linear_tc (tcgen05 3xFP16 forward / dgrad / split-K wgrad with dynamic operand scales) vs an fp64 torch reference,
    judged against the error of torch's own fp32 path on the same inputs ("fp32-equivalent");
  * hh_attention_rows (compacted-row attention forward / backward) vs the padded torch formulation in fp64;
  * Policy.evaluate_actions with the kernels on vs off: outputs and every parameter gradient.
Tolerances are written next to each assertion."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,K,N,act,gscale", [(1000, 128, 512, 1, 1.0), (777, 512, 1536, 0, 1e-7), (70001, 512, 256, 1, 3e-4),
                                                (130, 512, 512, 0, 1e3)])
def test_linear_tc_matches_fp64_reference(M, K, N, act, gscale):
    from crowdnav_prediction_attngraph_b200.update_ops import linear_tc
    g = torch.Generator(device=DEV).manual_seed(M + K)
    x = (torch.randn(M, K, device=DEV, generator=g) * torch.rand(M, 1, device=DEV, generator=g) * 3).requires_grad_(True)
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).requires_grad_(True)
    b = (torch.randn(N, device=DEV, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(M, N, device=DEV, generator=g) * gscale * torch.rand(1, N, device=DEV, generator=g)
    y = linear_tc(x, w, b, act)
    y.backward(dy)
    got = [y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone()]
    # fp64 reference and torch's fp32 path.  The ReLU mask is the KERNEL's forward mask for both references: an entry
    # whose pre-activation is within rounding noise of zero would otherwise flip between precisions and dominate the
    # comparison (one flipped entry moves dW by |dz x|, far above any arithmetic error).
    mask = (y.detach() > 0)
    outs = {}
    for dt in (torch.float64, torch.float32):
        xx, ww, bb = [t.detach().to(dt).requires_grad_(True) for t in (x, w, b)]
        yy = torch.nn.functional.linear(xx, ww, bb)
        if act:
            yy = yy * mask.to(dt)
        yy.backward(dy.to(dt))
        outs[dt] = [yy.detach(), xx.grad, ww.grad, bb.grad]
    errs = {name: (_rel(a, r64), _rel(r32, r64)) for name, a, r64, r32 in
            zip(("y", "dx", "dw", "db"), got, outs[torch.float64], outs[torch.float32])}
    print("linear_tc M=%d K=%d N=%d act=%d: (kernel error, torch fp32 error) vs fp64" % (M, K, N, act), errs)
    for name, (e_tc, e_32) in errs.items():
        # fp32-equivalent: within 4x of torch's own fp32 error, or below 2e-6 of the tensor's largest entry -- times
        # sqrt(M / 1000) for the reductions over the M rows (dW, db): the tensor core accumulates the whole K range of a
        # slice sequentially in fp32, so its rounding error grows like a random walk in the reduction length (measured
        # 7e-6 at M = 70 001, where torch's blocked SIMT summation reaches 6e-7)
        floor = 2e-6 * (max(1.0, (M / 1000.0) ** 0.5) if name in ("dw", "db") else 1.0)
        assert e_tc <= max(4 * e_32, floor), (name, errs)


def _segments(B, H, seed):
    rs = np.random.RandomState(seed)
    n = rs.randint(1, H + 1, size=B)
    return torch.from_numpy(n).to(DEV)


@pytest.mark.parametrize("B,H", [(300, 20), (64, 50), (5, 128)])
def test_hh_attention_rows_matches_padded_fp64(B, H):
    from crowdnav_prediction_attngraph_b200.update_ops import hh_attention_rows
    n = _segments(B, H, B + H)
    Mc = int(n.sum())
    row_start = torch.zeros(B + 1, dtype=torch.int32, device=DEV)
    row_start[1:] = torch.cumsum(n, 0)
    row_env = torch.repeat_interleave(torch.arange(B, device=DEV, dtype=torch.int32), n)
    g = torch.Generator(device=DEV).manual_seed(7)
    qkv = (torch.randn(Mc, 1536, device=DEV, generator=g) * 1.5).requires_grad_(True)
    dout = torch.randn(Mc, 512, device=DEV, generator=g)
    out = hh_attention_rows(qkv, row_start, row_env)
    out.backward(dout)
    # padded fp64 reference
    valid = torch.arange(H, device=DEV)[None, :] < n[:, None]
    q64 = qkv.detach().double().requires_grad_(True)
    pad = q64.new_zeros(B, H, 1536)
    pad[valid] = q64
    q, k, v = [t.reshape(B, H, 8, 64).transpose(1, 2) for t in pad.chunk(3, -1)]
    s = torch.matmul(q, k.transpose(-1, -2)) * 0.125
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    o = torch.matmul(torch.softmax(s, -1), v).transpose(1, 2).reshape(B, H, 512)[valid]
    o.backward(dout.double())
    assert _rel(out.detach(), o.detach()) <= 2e-6
    assert _rel(qkv.grad, q64.grad) <= 5e-6


def test_evaluate_actions_kernels_on_equals_off():
    """One minibatch [T=30, N=48] through Policy.evaluate_actions with the update kernels and with plain torch ops:
    value / log-prob / entropy and every parameter gradient of a PPO-like loss."""
    from crowdnav_prediction_attngraph_b200.policy import Policy
    from crowdnav_prediction_attngraph_b200.vec_env import Box
    T, N, H = 30, 48, 20

    class Args(object):
        num_processes, seq_length, num_mini_batch = N, T, 1
    spaces = {'robot_node': Box((1, 7)), 'temporal_edges': Box((1, 2)), 'spatial_edges': Box((H, 12)), 'detected_human_num': Box((1,))}
    torch.manual_seed(3)
    pol = Policy(spaces, Box((2,)), base_kwargs=Args(), base='selfAttn_merge_srnn').to(DEV)
    g = torch.Generator(device=DEV).manual_seed(11)
    B = T * N
    obs = {'robot_node': torch.randn(B, 1, 7, device=DEV, generator=g), 'temporal_edges': torch.randn(B, 1, 2, device=DEV, generator=g),
           'spatial_edges': torch.randn(B, H, 12, device=DEV, generator=g) * 3,
           'detected_human_num': torch.randint(1, H + 1, (B, 1), device=DEV, generator=g).float()}
    hx = {'human_node_rnn': torch.randn(N, 1, 128, device=DEV, generator=g) * 0.3}
    masks = (torch.rand(B, 1, device=DEV, generator=g) > 0.05).float()
    act = torch.randn(B, 2, device=DEV, generator=g)
    import copy
    res = {}
    for tag, module, on in (("tc", pol, True), ("torch32", pol, False), ("fp64", copy.deepcopy(pol).double(), False)):
        module.update_kernels = on
        module.zero_grad()
        dt = torch.float64 if tag == "fp64" else torch.float32
        v, lp, ent, h = module.evaluate_actions(obs, {'human_node_rnn': hx['human_node_rnn'].to(dt)}, masks.to(dt), act.to(dt))
        (0.5 * v.pow(2).mean() - lp.mean() + 0.01 * ent).backward()
        res[tag] = (v.detach().double(), lp.detach().double(), float(ent.detach()),
                    {k: p.grad.double().clone() for k, p in module.named_parameters() if p.grad is not None})
    assert torch.allclose(res["tc"][0], res["fp64"][0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(res["tc"][1], res["fp64"][1], rtol=1e-5, atol=1e-5)
    assert abs(res["tc"][2] - res["fp64"][2]) <= 1e-6
    # Gradients: element-wise comparison against fp64 is meaningless here -- ReLU masks flip for pre-activations within
    # rounding noise of zero, and ONE flipped entry with an outlier upstream gradient moves a weight-gradient entry by
    # orders of magnitude more than any arithmetic error (both fp32 paths sit 3e-3 from fp64 on some tensors, and on
    # k_linear.bias the true gradient is zero).  So: L2 distance between the two fp32 paths relative to the gradient's
    # norm, and the same for each of them against fp64 -- the kernel path must be as close to fp64 as torch's fp32 path.
    report = []
    for k, g64 in res["fp64"][3].items():
        n64 = float(g64.norm())
        if n64 < 1e-12:
            continue
        d_tc32 = float((res["tc"][3][k] - res["torch32"][3][k]).norm()) / n64
        d_tc64 = float((res["tc"][3][k] - g64).norm()) / n64
        d_3264 = float((res["torch32"][3][k] - g64).norm()) / n64
        report.append((d_tc64, d_3264, d_tc32, k))
    report.sort(reverse=True)
    print("relative L2 gradient distances (kernels-fp64, torch32-fp64, kernels-torch32), worst five:", report[:5])
    for d_tc64, d_3264, d_tc32, k in report:
        # measured: torch's fp32 path is up to 2e-3 from fp64 (mask flips), the kernel path up to 2.8e-3 on the same
        # tensors and 1.2e-4 where torch reaches 1e-5 (first embedding layer: the longest chain of fp32-accumulated GEMMs)
        assert d_tc64 <= max(3 * d_3264, 5e-4), (k, d_tc64, d_3264)
        assert d_tc32 <= max(3 * d_3264, 5e-4), (k, d_tc32, d_3264)


def test_gru_sequence_matches_torch_loop():
    """Fused GRU over T = 30 steps with mid-sequence resets vs the eager per-step loop (fp64), forward and all gradients."""
    from crowdnav_prediction_attngraph_b200.update_ops import gru_sequence
    T, N = 30, 77
    g = torch.Generator(device=DEV).manual_seed(5)
    gi = (torch.randn(T, N, 384, device=DEV, generator=g)).requires_grad_(True)
    h0 = (torch.randn(N, 128, device=DEV, generator=g) * 0.5).requires_grad_(True)
    masks = (torch.rand(T, N, device=DEV, generator=g) > 0.1).float()
    whh = (torch.randn(384, 128, device=DEV, generator=g) * 0.1).requires_grad_(True)
    bhh = (torch.randn(384, device=DEV, generator=g) * 0.1).requires_grad_(True)
    dout = torch.randn(T, N, 128, device=DEV, generator=g)
    out = gru_sequence(gi, h0, masks, whh, bhh)
    out.backward(dout)
    got = [out.detach(), gi.grad, h0.grad, whh.grad, bhh.grad]
    gi64, h64, w64, b64 = [t.detach().double().requires_grad_(True) for t in (gi, h0, whh, bhh)]
    h = h64
    outs = []
    for t in range(T):
        h = h * masks[t].double().unsqueeze(-1)
        gh = torch.nn.functional.linear(h, w64, b64)
        ir, iz, inn = gi64[t].chunk(3, -1)
        hr, hz, hn = gh.chunk(3, -1)
        r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
        n = torch.tanh(inn + r * hn)
        h = (1 - z) * n + z * h
        outs.append(h)
    ref = torch.stack(outs, 0)
    ref.backward(dout.double())
    for name, a, r in zip(("out", "d_gi", "d_h0", "d_whh", "d_bhh"), got, (ref.detach(), gi64.grad, h64.grad, w64.grad, b64.grad)):
        assert _rel(a, r) <= 5e-6, (name, _rel(a, r))       # fp32 arithmetic over a 30-step recurrence
