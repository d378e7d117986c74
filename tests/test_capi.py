"""The C-ABI library loads, exports every symbol include/crowdnav_b200.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest
import torch

from crowdnav_prediction_attngraph_b200 import _capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _capi.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(REPO, "include", "crowdnav_b200.h")).read()
    declared = set(re.findall(r"\b(cn_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cn_abi_version() == 3


def test_struct_layout_matches_header():
    # 16 int32 + 20 double, no padding surprises
    assert C.sizeof(_capi.CnConfig) == 16 * 4 + 20 * 8
    assert C.sizeof(_capi.CnCopySeg) == 3 * 8
    assert C.sizeof(_capi.CnObsPtrs) == 5 * 8 and C.sizeof(_capi.CnStepPtrs) == 7 * 8
    assert C.sizeof(_capi.CnActPtrs) == 12 * 8 and C.sizeof(_capi.CnPolicyConfig) == 5 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_is_a_loud_error(lib):
    cfg = _capi.config_from_dict(_capi.default_config_dict())
    h = C.c_void_p()
    assert lib.cn_env_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.cn_last_error() or b"CUDA" in lib.cn_last_error()
    pc = _capi.CnPolicyConfig(4, 20, 12, 0, 0)
    assert lib.cn_policy_create(C.byref(pc), C.byref(h)) != 0
    with pytest.raises(RuntimeError):
        from crowdnav_prediction_attngraph_b200.vec_env import CudaCrowdVecEnv
        CudaCrowdVecEnv(num_envs=4, device="cpu")
