from . import is_torch_version  # noqa: F401


def maybe_allow_in_graph(cls):
    return cls


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """upstream diffusers.utils.torch_utils.randn_tensor (recalled): drawn on the generator's device, moved to `device`."""
    import torch

    gdev = generator.device if generator is not None and not isinstance(generator, list) else (device or "cpu")
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        return torch.cat([torch.randn(one, generator=g, device=g.device, dtype=dtype) for g in generator]).to(device)
    return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)
