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


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors == the C structs of include/crowdnav_b200.h: a C program compiled against the header prints
    sizeof / offsetof of every cn_config field, compared with the ctypes layout field by field."""
    import subprocess
    fields = [n for n, _ in _capi.CnConfig._fields_]
    src = tmp_path / "layout.c"
    body = "".join('  printf("%s %%zu\\n", offsetof(cn_config, %s));\n' % (f, f) for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "crowdnav_b200.h"\nint main(void) {\n'
                   '  printf("sizeof %zu %zu %zu %zu %zu %zu\\n", sizeof(cn_config), sizeof(cn_copy_seg), sizeof(cn_obs_ptrs),\n'
                   '         sizeof(cn_step_ptrs), sizeof(cn_act_ptrs), sizeof(cn_policy_config));\n' + body + '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    sizes = [int(x) for x in out[0].split()[1:]]
    assert sizes == [C.sizeof(_capi.CnConfig), C.sizeof(_capi.CnCopySeg), C.sizeof(_capi.CnObsPtrs), C.sizeof(_capi.CnStepPtrs),
                     C.sizeof(_capi.CnActPtrs), C.sizeof(_capi.CnPolicyConfig)], sizes
    for line in out[1:]:
        name, off = line.split()
        assert getattr(_capi.CnConfig, name).offset == int(off), name
    assert C.sizeof(_capi.CnConfig) == 18 * 4 + 23 * 8


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
