"""Replay a golden env fixture (tests/golden/env_*.npz) through any engine exposing
reset()/step(actions)/get(name) and count mismatching steps.  Shared by the host-harness
(CPU) test and the CUDA (gpu) test so both read exactly like the same parity check."""
import ast
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV_CASES = ["env_pred_h20", "env_pred_h20_rand", "env_pred_h50_rand", "env_varnum_h5", "env_pred_h20_test",
             "env_pred_h10_test_rand", "env_varnum_h5_test",
             # sim.human_num_range > 0: humans join / leave every 5 s, observations padded to max_human_num
             "env_varnum_h5_range2", "env_pred_h6_range3",
             # humans.policy = 'social_force' (crowd_nav/policy/social_force.py), randomised attributes + goal changes
             "env_pred_h8_sf"]


def load_env_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    case = ast.literal_eval(str(g["meta"][0]))
    over = dict(num_envs=g["actions"].shape[1], nenv_total=case["nenv"], seed=case["seed"],
                human_num=case["human_num"], const_vel=1 if case["predict_method"] == "const_vel" else 0,
                randomize_attributes=int(case["randomize"]), random_goal_changing=int(case["goal_changing"]),
                phase=2 if case.get("phase", "train") == "test" else 0, human_num_range=int(case.get("human_num_range", 0)),
                human_policy=1 if case.get("human_policy", "orca") == "social_force" else 0)
    return g, case, over


def replay(g, case, reset_fn, step_fn, get_fn, pos_tol=1e-9, obs_tol=1e-5, exact_orca=True):
    """Returns a list of human-readable mismatch strings (empty = parity)."""
    T, N = g["actions"].shape[:2]
    H = case["human_num"] + int(case.get("human_num_range", 0))      # slots = max_human_num; absent humans are NaN in the goldens
    bad = []
    ob = reset_fn()
    for k in ob:
        ref = g["ob_" + k][0].reshape(ob[k].shape)
        if ref.dtype == bool:
            if not np.array_equal(ob[k].astype(bool), ref):
                bad.append("reset obs %s" % k)
        elif np.abs(ob[k].astype(np.float64) - ref).max() > obs_tol:
            bad.append("reset obs %s" % k)
    for t in range(T):
        ob, out = step_fn(g["actions"][t])
        msg = []
        if not np.array_equal(out["done"].astype(bool), g["done"][t]):
            msg.append("done")
        if not np.array_equal(out["info"], g["info"][t]):
            msg.append("info")
        if np.abs(out["reward"] - g["reward"][t]).max() > 1e-5:
            msg.append("reward")
        if "info_aux" in out and np.abs(out["info_aux"] - g["min_danger"][t]).max() > 1e-6:
            msg.append("min_danger")
        ha = np.stack([get_fn("last_hvx").reshape(N, H), get_fn("last_hvy").reshape(N, H)], -1)
        ok = ~np.isnan(g["human_actions"][t][..., 0])
        if exact_orca:
            if not np.array_equal(ha[ok], g["human_actions"][t][ok]):
                msg.append("orca_velocity(bits)")
            if not np.array_equal(get_fn("orca_nlines").reshape(N, H)[ok], g["orca_nlines"][t][ok]):
                msg.append("orca_nlines")
            if not np.array_equal(get_fn("orca_fail").reshape(N, H)[ok], g["orca_fail"][t][ok]):
                msg.append("orca_fail")
        elif np.abs(ha[ok] - g["human_actions"][t][ok]).max() > 1e-5:
            msg.append("orca_velocity")
        for k in ob:
            ref = g["ob_" + k][t + 1].reshape(ob[k].shape)
            if ref.dtype == bool:
                if not np.array_equal(ob[k].astype(bool), ref):
                    msg.append("obs " + k)
            elif np.abs(ob[k].astype(np.float64) - ref).max() > obs_tol:
                msg.append("obs " + k)
        for k in ("hpx", "hpy", "hgx", "hgy", "hrad", "hvpref"):
            if np.abs(get_fn(k).reshape(N, H) - g["st_" + k][t + 1]).max() > pos_tol:
                msg.append("state " + k)
        rob = np.stack([get_fn(k) for k in ("rpx", "rpy")], -1)
        if np.abs(rob - g["st_robot"][t + 1][:, :2]).max() > pos_tol:
            msg.append("state robot")
        if not np.array_equal(get_fn("vis").reshape(N, H).astype(bool), g["st_vis"][t + 1]):
            msg.append("visibility")
        if np.abs(get_fn("potential") - g["st_potential"][t + 1]).max() > pos_tol:
            msg.append("potential")
        if np.abs(get_fn("nd_global") - g["st_nd_global"][t + 1]).max() > pos_tol:
            msg.append("nd_global")
        if "st_count" in g.files and not np.array_equal(get_fn("hn"), g["st_count"][t + 1]):
            msg.append("human count")
        if not np.array_equal(get_fn("sim_exists").reshape(N, H).astype(bool), g["st_sim_exists"][t + 1]):
            msg.append("sim_exists")
        if msg:
            bad.append("t=%d: %s" % (t, ",".join(msg)))
    return bad
