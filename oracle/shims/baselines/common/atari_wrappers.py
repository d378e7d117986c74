def make_atari(*a, **k):
    raise NotImplementedError


def wrap_deepmind(*a, **k):
    raise NotImplementedError
