import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kwargs):
        if not hasattr(self, "_internal_dict"):
            self._internal_dict = FrozenDict(kwargs)
        else:
            self._internal_dict = FrozenDict({**self._internal_dict, **kwargs})

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = list(sig.parameters.items())[1:]
        cfg = {name: p.default for name, p in params if p.default is not inspect.Parameter.empty}
        for (name, _), a in zip(params, args):
            cfg[name] = a
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        init(self, *args, **kwargs)
        self.register_to_config(**cfg)

    return inner_init
