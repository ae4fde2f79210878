"""Checkpoint loading for the native transformers: a diffusers-format directory
(``config.json`` + ``diffusion_pytorch_model*.safetensors``, single file or sharded with an index) goes straight into a
``frameino_b200`` model — the way the reference builds its transformer, ``WanTransformer3DModel.from_pretrained(ckpt,
torch_dtype=...)`` (reference app.py:150-156; class from architecture/transformer_wan.py:353), without diffusers.

What the loader does, and why:
  * state-dict key names are diffusers' own (the native modules keep them), so no renaming;
  * dtype policy of ``from_pretrained(torch_dtype=bf16)``: every tensor is cast to ``torch_dtype`` except the modules a
    class lists in ``_keep_in_fp32_modules`` (transformer_wan.py:393), which stay fp32. The reference demo loads its Wan
    transformer as fp16 and then runs the pipeline in bf16 (app.py:156 vs :161); the kernels here compute in bf16, so
    bf16 is the default and anything else is refused at the first forward;
  * tensors are cast and moved to the target device one shard at a time and ASSIGNED to the module (no second copy; the
    5B model never exists in host memory twice); parameters are created without random init;
  * afterwards ``prepare()`` concatenates the q/k/v (self) and k/v (cross) projection weights the fused GEMMs read, so
    the first forward does not pay for it.
``save_pretrained`` writes the same layout (used by the tests; sharded when asked to).
"""
from __future__ import annotations

import json
import os
import re
from contextlib import contextmanager
from typing import Dict, Iterable, Iterator, List, Optional, Union

import torch
from torch import nn

WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
INDEX_NAME = "diffusion_pytorch_model.safetensors.index.json"
CONFIG_NAME = "config.json"


@contextmanager
def no_init_weights() -> Iterator[None]:
    """Module constructors allocate parameters but skip their random initialisation (every value is about to be
    overwritten by the checkpoint; initialising 5 B parameters on the host costs tens of seconds)."""
    classes = (nn.Linear, nn.Conv2d, nn.Conv3d, nn.LayerNorm, nn.Embedding)
    saved = {c: c.reset_parameters for c in classes}
    for c in classes:
        c.reset_parameters = lambda self: None
    try:
        yield
    finally:
        for c, fn in saved.items():
            c.reset_parameters = fn


def checkpoint_files(path: str) -> List[str]:
    """The weight files of a diffusers model directory (or the file itself), in load order."""
    if os.path.isfile(path):
        return [path]
    if not os.path.isdir(path):
        raise FileNotFoundError(f"{path}: no such checkpoint directory (remote hub ids are not resolved here)")
    index = os.path.join(path, INDEX_NAME)
    if os.path.exists(index):
        with open(index) as f:
            weight_map = json.load(f)["weight_map"]
        files = sorted(set(weight_map.values()))
        missing = [fn for fn in files if not os.path.exists(os.path.join(path, fn))]
        if missing:
            raise FileNotFoundError(f"{path}: shards listed in {INDEX_NAME} are missing: {missing}")
        return [os.path.join(path, fn) for fn in files]
    single = os.path.join(path, WEIGHTS_NAME)
    if os.path.exists(single):
        return [single]
    found = sorted(fn for fn in os.listdir(path) if fn.endswith(".safetensors"))
    if not found:
        raise FileNotFoundError(f"{path}: no *.safetensors weights found")
    return [os.path.join(path, fn) for fn in found]


def iter_checkpoint_tensors(files: Iterable[str], device: Union[str, torch.device] = "cpu"):
    """Yields (name, tensor) shard by shard; tensors are read straight onto ``device``."""
    from safetensors import safe_open

    for fn in files:
        with safe_open(fn, framework="pt", device=str(device)) as f:
            for name in f.keys():
                yield name, f.get_tensor(name)


def read_config(path: str) -> Dict:
    """``config.json`` of a diffusers model directory minus the bookkeeping keys (``_class_name`` ...)."""
    cfg_file = os.path.join(path, CONFIG_NAME) if os.path.isdir(path) else os.path.join(os.path.dirname(path), CONFIG_NAME)
    if not os.path.exists(cfg_file):
        raise FileNotFoundError(f"{cfg_file} not found: pass the model directory that holds config.json")
    with open(cfg_file) as f:
        cfg = json.load(f)
    return {k: v for k, v in cfg.items() if not k.startswith("_")}


def target_dtype(model_cls, name: str, tensor: torch.Tensor, torch_dtype: torch.dtype) -> torch.dtype:
    if not tensor.is_floating_point():
        return tensor.dtype
    keep = getattr(model_cls, "_keep_in_fp32_modules", None) or []
    return torch.float32 if any(k in name for k in keep) else torch_dtype


def load_into(model: nn.Module, files: List[str], torch_dtype: torch.dtype, device: Union[str, torch.device]) -> None:
    """Casts, moves and assigns every checkpoint tensor; raises on missing / unexpected / mis-shaped keys."""
    expected = dict(model.state_dict())
    ignore = [re.compile(p) for p in (getattr(model, "_keys_to_ignore_on_load_unexpected", None) or [])]
    loaded: Dict[str, torch.Tensor] = {}
    unexpected: List[str] = []
    for name, t in iter_checkpoint_tensors(files, device):
        if name not in expected:
            if not any(p.search(name) for p in ignore):
                unexpected.append(name)
            continue
        if tuple(t.shape) != tuple(expected[name].shape):
            raise ValueError(f"{name}: checkpoint shape {tuple(t.shape)} != model shape {tuple(expected[name].shape)}")
        loaded[name] = t.to(device=device, dtype=target_dtype(type(model), name, t, torch_dtype)).contiguous()
    missing = [k for k in expected if k not in loaded]
    if missing or unexpected:
        raise ValueError(f"checkpoint does not match {type(model).__name__}: missing keys {missing[:8]}"
                         f"{' ...' if len(missing) > 8 else ''} ({len(missing)}), unexpected keys {unexpected[:8]}"
                         f"{' ...' if len(unexpected) > 8 else ''} ({len(unexpected)})")
    model.load_state_dict(loaded, strict=True, assign=True)
    # what the checkpoint does not hold (computed, non-persistent buffers such as the RoPE tables) follows the weights
    for mod in model.modules():
        for bname, buf in list(mod._buffers.items()):
            if buf is not None and bname in mod._non_persistent_buffers_set and buf.device != torch.device(device):
                mod._buffers[bname] = buf.to(device)


def from_pretrained(model_cls, path: str, subfolder: Optional[str] = None,
                    torch_dtype: torch.dtype = torch.bfloat16, device: Union[str, torch.device, None] = None,
                    prepare: bool = True, **config_overrides):
    """``model_cls.from_pretrained`` (see the module docstring). ``device`` defaults to the current CUDA device when
    there is one (the model cannot run anywhere else), else CPU (inspection / tests)."""
    if subfolder:
        path = os.path.join(path, subfolder)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    device = torch.device(device)
    cfg = read_config(path)
    cfg.update(config_overrides)
    import inspect

    from .modules import compute_dtype

    if not getattr(model_cls, "_stores_any_dtype", False):  # the VAE keeps its checkpoint dtype (fp32, app.py:157)
        torch_dtype = compute_dtype(torch_dtype)  # float16 (reference app.py:156) -> bf16 with a warning; fp32 refused

    accepted = set(inspect.signature(model_cls.__init__).parameters) - {"self"}
    dropped = sorted(k for k in cfg if k not in accepted)
    if dropped:
        from .modules import logger

        logger.warning("config keys %s are not used by %s and were ignored", dropped, model_cls.__name__)
    with no_init_weights():
        model = model_cls(**{k: v for k, v in cfg.items() if k in accepted})
    load_into(model, checkpoint_files(path), torch_dtype, device)
    model.eval()
    if prepare and device.type == "cuda" and hasattr(model, "prepare"):
        model.prepare()
    return model


def shard_state_dict(sd: Dict[str, torch.Tensor], max_shard_bytes: int) -> List[Dict[str, torch.Tensor]]:
    shards: List[Dict[str, torch.Tensor]] = [{}]
    size = 0
    for k in sd:
        n = sd[k].numel() * sd[k].element_size()
        if shards[-1] and size + n > max_shard_bytes:
            shards.append({})
            size = 0
        shards[-1][k] = sd[k]
        size += n
    return shards


def save_pretrained(model: nn.Module, path: str, max_shard_bytes: int = 10 << 30) -> List[str]:
    """Writes ``config.json`` and the weights in the diffusers layout; returns the weight file names."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    cfg = {"_class_name": type(model).__name__, "_frameino_b200": True}
    for k, v in dict(model.config).items():
        cfg[k] = list(v) if isinstance(v, tuple) else v
    with open(os.path.join(path, CONFIG_NAME), "w") as f:
        json.dump(cfg, f, indent=2)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in model.state_dict().items()}
    shards = shard_state_dict(sd, max_shard_bytes)
    if len(shards) == 1:
        save_file(shards[0], os.path.join(path, WEIGHTS_NAME))
        return [WEIGHTS_NAME]
    names, weight_map = [], {}
    for i, shard in enumerate(shards):
        fn = f"diffusion_pytorch_model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(shard, os.path.join(path, fn))
        names.append(fn)
        weight_map.update({k: fn for k in shard})
    total = sum(v.numel() * v.element_size() for v in sd.values())
    with open(os.path.join(path, INDEX_NAME), "w") as f:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=2)
    return names
