"""rl/networks/storage.py -> the engine's device-resident RolloutStorage (same constructor and methods)."""
from crowdnav_prediction_attngraph_b200.storage import RolloutStorage  # noqa: F401
