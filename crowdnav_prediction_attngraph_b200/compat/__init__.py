"""Alias packages under the REFERENCE'S module names, so that its own `train.py` / `test.py` run unchanged on the
CUDA engine:

    PYTHONSAFEPATH=1 PYTHONPATH=<this dir>:<reference root> python <reference root>/train.py ...

(PYTHONSAFEPATH keeps Python from putting the script's own directory in front of PYTHONPATH.)  Shadowed:
`rl.networks.envs` (make_vec_envs), `rl.networks.model` (Policy), `rl.networks.storage` (RolloutStorage), `rl.ppo`
(PPO), `crowd_sim` (gym ids + info classes).  Everything else of `rl.*` (network_utils, evaluation, ...) resolves to
the reference's own files: `rl/__init__.py` here appends the reference's `rl` directory to the package path.
Under torchrun (WORLD_SIZE > 1) `make_vec_envs` shards the environments across the ranks and `PPO.update`
all-reduces the gradient over NCCL; train.py itself stays unchanged.

    python -m crowdnav_prediction_attngraph_b200.compat        # prints this directory
"""
import os
import sys

PATH = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """The reference checkout: $CROWDNAV_REFERENCE_ROOT, else the first sys.path entry that holds rl/ppo/ppo.py."""
    env = os.environ.get("CROWDNAV_REFERENCE_ROOT")
    if env:
        return env
    for p in sys.path:
        p = p or os.getcwd()
        if os.path.realpath(p) != os.path.realpath(PATH) and os.path.isfile(os.path.join(p, "rl", "ppo", "ppo.py")) \
                and os.path.isfile(os.path.join(p, "arguments.py")):
            return p
    return None


def extend_with_reference(pkg_path, sub):
    """Append <reference root>/<sub> to a package __path__ so un-shadowed modules fall through to the reference."""
    root = reference_root()
    if root is not None:
        d = os.path.join(root, *sub.split("/"))
        if os.path.isdir(d) and d not in pkg_path:
            pkg_path.append(d)
