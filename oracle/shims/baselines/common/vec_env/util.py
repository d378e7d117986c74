from rl.vec_env.util import *  # noqa: F401,F403
from rl.vec_env.util import dict_to_obs, obs_space_info, obs_to_dict  # noqa: F401
