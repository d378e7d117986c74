import contextlib


def log(*a, **k):
    pass


def warn(*a, **k):
    pass


@contextlib.contextmanager
def scoped_configure(*a, **k):
    yield
