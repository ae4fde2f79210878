from contextlib import contextmanager


class CacheMixin:
    @contextmanager
    def cache_context(self, name):
        yield
