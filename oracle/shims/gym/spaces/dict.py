from collections import OrderedDict


class Dict(object):
    """Keys are sorted, as in gym.spaces.Dict for plain dict input."""

    def __init__(self, spaces):
        self.spaces = OrderedDict(sorted(spaces.items()))

    def __getitem__(self, k):
        return self.spaces[k]
