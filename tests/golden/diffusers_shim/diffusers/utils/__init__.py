import logging as _logging

USE_PEFT_BACKEND = False


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()


def scale_lora_layers(model, weight):
    pass


def unscale_lora_layers(model, weight=None):
    pass


def deprecate(*args, **kwargs):
    pass


def is_torch_xla_available():
    return False


def is_torch_version(op, version):
    import operator

    import torch
    from packaging import version as V

    ops = {">": operator.gt, ">=": operator.ge, "<": operator.lt, "<=": operator.le, "==": operator.eq}
    return ops[op](V.parse(V.parse(torch.__version__).base_version), V.parse(version))


def is_ftfy_available():
    return False


def replace_example_docstring(example_docstring):
    def wrap(fn):
        return fn

    return wrap
