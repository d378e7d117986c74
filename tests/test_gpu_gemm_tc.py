"""GPU test of the tcgen05 / TMEM / TMA GEMM (3xFP16 error-compensated) against an fp64 matmul,
through the internal hook cn_internal_gemm_tc, and of the whole policy forward in gemm_mode=1."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from crowdnav_prediction_attngraph_b200 import _capi
    lib = _capi.load_library()
    lib.cn_internal_gemm_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int]
    return lib, _capi


@pytest.mark.parametrize("M,N,K,act,bn", [(128, 256, 64, 0, 256), (256, 256, 128, 0, 256), (300, 512, 128, 1, 256),
                                          (4096, 1536, 512, 0, 256), (1000, 256, 512, 1, 256),
                                          (128, 64, 64, 0, 64), (4096, 384, 128, 0, 64), (300, 128, 256, 1, 64),
                                          (4096, 512, 256, 2, 64), (777, 256, 320, 0, 64)])
def test_gemm_tc_matches_fp64(M, N, K, act, bn):
    lib, _capi = _lib()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 2).cuda()
    A[:, ::7] = 0                                     # post-ReLU-like zeros
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    b = torch.randn(N, generator=g).cuda()
    Cout = torch.full((M, N), float("nan"), device="cuda")
    _capi.check(lib, lib.cn_internal_gemm_tc(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cout.data_ptr(), M, N, K, act, bn),
                "cn_internal_gemm_tc")
    ref = A.double() @ W.double().T + b.double()
    scale = ref.abs().max().item()                   # magnitude of the accumulated products
    if act == 1:
        ref = ref.clamp_min(0)
    if act == 2:
        ref = torch.tanh(ref)
    err = (Cout.double() - ref).abs().max().item()
    tol = 3e-6 * max(1.0, scale)
    assert err < tol, (err, scale)


@pytest.mark.parametrize("name,H", [("policy_h20", 20), ("policy_h50", 50)])
def test_cuda_policy_tensor_core_mode_matches_reference_golden(name, H):
    from oracle.policy_ref import PolicyRef
    from crowdnav_prediction_attngraph_b200.policy import CudaPolicy
    from tests.policy_fixture import load_policy_golden, synth_state_dict
    g, obs, h, masks = load_policy_golden(name)
    sd = synth_state_dict(PolicyRef(12).state_dict())
    pol = CudaPolicy(h.shape[0], H, 12, device="cuda:0", gemm_mode=1)
    pol.load_state_dict(sd)
    dobs = {k: v.cuda() for k, v in obs.items()}
    value, action, logp, h1, mean = pol.act(dobs, h.cuda(), masks.cuda(), deterministic=True, return_mean=True)
    np.testing.assert_allclose(mean.cpu().numpy(), g["synth_mean"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(value.cpu().numpy(), g["synth_value"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(h1.cpu().numpy(), g["synth_h"], rtol=0, atol=1e-4)
