import functools
import inspect


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class ConfigMixin:
    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        self._internal_dict = _Cfg(cfg)
        return init(self, *args, **kwargs)
    return wrapper
