"""The inert plumbing of upstream DiffusionPipeline that the reference pipeline's __init__ / __call__ touch."""
from contextlib import contextmanager

import torch

from ..configuration_utils import FrozenDict


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def register_to_config(self, **kwargs):
        self._internal_dict = FrozenDict({**getattr(self, "_internal_dict", {}), **kwargs})

    @property
    def config(self):
        return self._internal_dict

    @property
    def _execution_device(self):
        return torch.device("cpu")

    @contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def maybe_free_model_hooks(self):
        pass
