"""ORACLE / TEST INFRASTRUCTURE: stand-in for OpenAI baselines' `baselines.common.vec_env.vec_env` (the reference
depends on the un-vendored, un-pinned `baselines`, README.md:32-37).  Restates the published interface: an abstract
VecEnv (reset / step_async / step_wait, step = async + wait), a VecEnvWrapper that forwards to `.venv` and passes
unknown public attributes through, the cloudpickle carrier and the MPI-environment guard.  Unlike the reference's
vendored rl/vec_env/vec_env.py, OpenAI's classes have NO abstract talk2Env_* methods -- rl/networks/envs.py:193
(VecPyTorch) only instantiates against this shape."""
import contextlib
import os
from abc import ABC, abstractmethod


class VecEnv(ABC):
    closed = False
    viewer = None
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def render(self, mode='human'):
        raise NotImplementedError

    @property
    def unwrapped(self):
        if isinstance(self, VecEnvWrapper):
            return self.venv.unwrapped
        return self


class VecEnvWrapper(VecEnv):
    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        super().__init__(num_envs=venv.num_envs, observation_space=observation_space or venv.observation_space,
                         action_space=action_space or venv.action_space)

    def step_async(self, actions):
        self.venv.step_async(actions)

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def close(self):
        return self.venv.close()

    def render(self, mode='human'):
        return self.venv.render(mode=mode)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.venv, name)


class CloudpickleWrapper(object):
    def __init__(self, x):
        self.x = x

    def __getstate__(self):
        import cloudpickle
        return cloudpickle.dumps(self.x)

    def __setstate__(self, ob):
        import pickle
        self.x = pickle.loads(ob)


@contextlib.contextmanager
def clear_mpi_env_vars():
    removed = {}
    for k, v in list(os.environ.items()):
        if k.startswith('OMPI_') or k.startswith('PMI_'):
            removed[k] = v
            del os.environ[k]
    try:
        yield
    finally:
        os.environ.update(removed)
