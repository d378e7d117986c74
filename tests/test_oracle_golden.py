"""Pins oracle/crowd_env.py (+ oracle/rvo2_ref.cpp) against golden vectors recorded from the
UNMODIFIED reference environment (tools/make_golden.py).

Integer outputs (done, info code, detected_human_num, visibility) must match bit-exactly;
float outputs within 1e-6 (fp32 observations) / 1e-9 (fp64 state) — slack only covers
libm/BLAS last-bit differences between machines, the oracle uses the reference's own call forms.
"""
import ast
import os

import numpy as np
import pytest

from oracle.crowd_env import CrowdEnvOracle, EnvConfig

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["env_pred_h20", "env_pred_h20_rand", "env_pred_h50_rand", "env_varnum_h5", "env_pred_h20_test",
         "env_pred_h10_test_rand", "env_varnum_h5_test"]


def load_case(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s missing" % name)
    g = np.load(path, allow_pickle=False)
    case = ast.literal_eval(str(g["meta"][0]))
    cfg = EnvConfig(human_num=case["human_num"], predict_method=case["predict_method"],
                    randomize_attributes=case["randomize"], random_goal_changing=case["goal_changing"])
    return g, case, cfg


def check_state(st, g, t, k, tol=1e-9):
    np.testing.assert_allclose(st["robot"], g["st_robot"][t, k], rtol=0, atol=tol)
    for key in ("hpx", "hpy", "hvx", "hvy", "hgx", "hgy", "hrad", "hvpref", "belief"):
        np.testing.assert_allclose(st[key], g["st_" + key][t, k], rtol=0, atol=tol, err_msg=key)
    assert np.array_equal(st["vis"], g["st_vis"][t, k])
    assert st["global_time"] == g["st_global_time"][t, k]
    np.testing.assert_allclose(st["potential"], g["st_potential"][t, k], rtol=0, atol=tol)
    np.testing.assert_allclose(st["nd_global"], g["st_nd_global"][t, k], rtol=0, atol=tol)
    assert np.array_equal(st["sim_exists"], g["st_sim_exists"][t, k])
    if g["st_traj"][t, k].size:
        np.testing.assert_allclose(st["traj"], g["st_traj"][t, k], rtol=0, atol=tol)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g, case, cfg = load_case(name)
    T, N = g["actions"].shape[:2]
    obs_keys = [k[3:] for k in g.files if k.startswith("ob_")]
    for k in range(N):
        env = CrowdEnvOracle(cfg, case["seed"] + k, case["nenv"], case.get("phase", "train"))
        ob = env.reset()
        for key in obs_keys:
            np.testing.assert_allclose(ob[key], g["ob_" + key][0, k], rtol=0, atol=1e-6, err_msg=key)
        check_state(env.get_state(), g, 0, k)
        for t in range(T):
            a = g["actions"][t, k].copy()
            ob, rew, done, info = env.worker_step(a)
            assert bool(done) == bool(g["done"][t, k]), (name, k, t)
            assert info["info"] == g["info"][t, k], (name, k, t)
            np.testing.assert_allclose(rew, g["reward"][t, k], rtol=0, atol=1e-9)
            np.testing.assert_allclose(info["min_danger"], g["min_danger"][t, k], rtol=0, atol=1e-9)
            # the fixture reads the per-human simulators after the step = the last ORCA solve (in the test
            # phase that is the final ground-truth look-ahead step)
            ha = np.asarray(env.last_sim_actions, dtype=np.float32)
            ref_ha = g["human_actions"][t, k]
            ok = ~np.isnan(ref_ha[:, 0])
            assert np.array_equal(ha[ok], ref_ha[ok]), (name, k, t)        # bit-exact fp32 ORCA output
            diag = np.asarray(env.last_orca_diag)
            assert np.array_equal(diag[ok, 0], g["orca_nlines"][t, k][ok])
            assert np.array_equal(diag[ok, 1], g["orca_fail"][t, k][ok])
            for key in obs_keys:
                if g["ob_" + key].dtype == bool:
                    assert np.array_equal(ob[key], g["ob_" + key][t + 1, k])
                else:
                    np.testing.assert_allclose(ob[key], g["ob_" + key][t + 1, k], rtol=0, atol=1e-6,
                                               err_msg="%s t=%d" % (key, t))
            check_state(env.get_state(), g, t + 1, k)


RANGE_CASES = ["env_varnum_h5_range2", "env_pred_h6_range3", "env_pred_h8_sf"]


@pytest.mark.parametrize("name", RANGE_CASES)
def test_oracle_variable_human_count_matches_reference_golden(name):
    """SURVEY 8f row 4, oracle only so far (the CUDA engine still rejects these settings): social-force humans
    (env_pred_h8_sf) and sim.human_num_range > 0: humans join and
    leave every 5 s, per-human ORCA simulators are rebuilt when the agent count changes, observations are padded to
    max_human_num.  Fixture arrays are NaN-padded; st_count is the live human count."""
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s missing" % name)
    g = np.load(path, allow_pickle=False)
    case = ast.literal_eval(str(g["meta"][0]))
    cfg = EnvConfig(human_num=case["human_num"], human_num_range=case.get("human_num_range", 0),
                    human_policy=case.get("human_policy", "orca"),
                    predict_method=case["predict_method"], randomize_attributes=case["randomize"],
                    random_goal_changing=case["goal_changing"])
    T, N = g["actions"].shape[:2]
    obs_keys = [k[3:] for k in g.files if k.startswith("ob_")]
    counts = set()
    for k in range(N):
        env = CrowdEnvOracle(cfg, case["seed"] + k, case["nenv"], "train")
        ob = env.reset()
        for t in range(T + 1):
            n = int(g["st_count"][t, k])
            counts.add(n)
            st = env.get_state()
            assert len(st["hpx"]) == n, (name, k, t)
            for key in ("hpx", "hpy", "hgx", "hgy", "hrad", "hvpref"):
                np.testing.assert_allclose(st[key], g["st_" + key][t, k][:n], rtol=0, atol=1e-9, err_msg="%s t=%d" % (key, t))
            np.testing.assert_allclose(st["belief"], g["st_belief"][t, k][:n], rtol=0, atol=1e-9)
            assert np.array_equal(st["vis"], g["st_vis"][t, k][:n])
            for key in obs_keys:
                ref = g["ob_" + key][t, k]
                if ref.dtype == bool:
                    assert np.array_equal(ob[key], ref), (key, t)
                else:
                    np.testing.assert_allclose(ob[key], ref, rtol=0, atol=1e-6, err_msg="%s t=%d" % (key, t))
            if t == T:
                break
            ob, rew, done, info = env.worker_step(g["actions"][t, k].copy())
            assert bool(done) == bool(g["done"][t, k]) and info["info"] == g["info"][t, k], (name, k, t)
            np.testing.assert_allclose(rew, g["reward"][t, k], rtol=0, atol=1e-9)
    if case.get("human_num_range", 0) > 0:
        assert len(counts) >= 3, "the fixture should exercise several human counts: %r" % (counts,)
