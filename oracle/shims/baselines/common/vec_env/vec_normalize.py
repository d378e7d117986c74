from rl.vec_env.vec_normalize import VecNormalize  # noqa: F401
