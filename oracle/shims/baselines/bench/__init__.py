import time


class Monitor(object):
    """Episode return/length bookkeeping like baselines.bench.Monitor (info['episode'] on done)."""

    def __init__(self, env, filename=None, allow_early_resets=False):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space
        self.tstart = time.time()
        self.rewards = []

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kw):
        self.rewards = []
        return self.env.reset(**kw)

    def step(self, action):
        ob, rew, done, info = self.env.step(action)
        self.rewards.append(rew)
        if done:
            info['episode'] = {'r': round(sum(self.rewards), 6), 'l': len(self.rewards),
                               't': round(time.time() - self.tstart, 6)}
        return ob, rew, done, info

    def close(self):
        self.env.close()
