"""ORACLE / TEST INFRASTRUCTURE: minimal stand-in for the `gym` package (absent in this image).

Only what the reference's crowd_sim package touches: Env, Wrapper, spaces.Box/Dict,
envs.registration.register/make.  Used solely by tools/make_golden.py to run the UNMODIFIED
reference environment in this container.
"""
from . import spaces  # noqa: F401
from .envs.registration import register, make  # noqa: F401
from . import envs  # noqa: F401


class Env(object):
    metadata = {}
    observation_space = None
    action_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode='human'):
        pass

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    pass
