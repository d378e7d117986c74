"""rl/ppo -> the engine's PPO (same constructor and update(); all-reduces gradients when torch.distributed is up)."""
from crowdnav_prediction_attngraph_b200.ppo import PPO  # noqa: F401
