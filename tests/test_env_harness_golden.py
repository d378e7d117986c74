"""CPU test of the step kernel's per-phase logic (host build of csrc/cn_env_core.cuh, see
tests/cpu_harness) against the golden vectors recorded from the unmodified reference.
Bit-exact on integers and on the fp32 ORCA velocities; fp64 state within 1e-9."""
import pytest

from tests.golden_util import ENV_CASES, load_env_case, replay
from tests.harness_util import HarnessEnv


@pytest.mark.parametrize("name", ENV_CASES)
def test_kernel_logic_host_build_matches_reference_golden(name):
    g, case, over = load_env_case(name)
    env = HarnessEnv(**over)
    bad = replay(g, case, env.reset, env.step, env.get)
    assert not bad, bad[:5]


def test_mt19937_matches_numpy_legacy():
    import ctypes as C
    import numpy as np
    env = HarnessEnv(num_envs=1, nenv_total=1)
    env.lib.harness_rng_doubles.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    for seed in (0, 1, 2425, 2 ** 32 - 1):
        out = np.zeros(2000)
        env.lib.harness_rng_doubles(seed, 2000, out.ctypes.data)
        assert np.array_equal(out, np.random.RandomState(seed).random_sample(2000))
