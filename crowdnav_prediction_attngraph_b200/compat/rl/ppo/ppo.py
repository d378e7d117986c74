from crowdnav_prediction_attngraph_b200.ppo import PPO  # noqa: F401
