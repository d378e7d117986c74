"""rl/networks/model.py -> the engine's Policy (same constructor, act / get_value / evaluate_actions, state_dict keys)."""
from crowdnav_prediction_attngraph_b200.policy import Policy  # noqa: F401
