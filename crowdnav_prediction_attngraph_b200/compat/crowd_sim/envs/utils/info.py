"""crowd_sim/envs/utils/info.py -> the info classes the engine's infos are instances of (rl/evaluation.py:3,
96-133 tests them with isinstance)."""
from crowdnav_prediction_attngraph_b200.vec_env import Nothing, Timeout, Collision, ReachGoal, Danger  # noqa: F401
