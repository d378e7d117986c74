import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low), self.shape)
        self.high = np.broadcast_to(np.asarray(high), self.shape)
