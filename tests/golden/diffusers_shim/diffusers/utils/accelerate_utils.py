def apply_forward_hook(method):
    """diffusers.utils.accelerate_utils.apply_forward_hook: runs accelerate's offload pre-forward hook if one is
    attached; inert here."""
    return method
