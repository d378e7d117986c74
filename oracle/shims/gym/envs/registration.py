import importlib

_registry = {}


def register(id, entry_point, **kwargs):
    _registry[id] = entry_point


def make(id):
    mod, cls = _registry[id].split(':')
    return getattr(importlib.import_module(mod), cls)()
