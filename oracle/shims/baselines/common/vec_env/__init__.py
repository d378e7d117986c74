# re-export the reference's own vendored copies (rl/vec_env) when it is importable
from rl.vec_env.vec_env import VecEnv, VecEnvWrapper, CloudpickleWrapper, clear_mpi_env_vars  # noqa: F401
