from rl.vec_env.vec_env import *  # noqa: F401,F403
from rl.vec_env.vec_env import VecEnv, VecEnvWrapper, CloudpickleWrapper, clear_mpi_env_vars  # noqa: F401
