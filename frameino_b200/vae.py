"""B200-native ``AutoencoderKLWan`` — drop-in for the reference class of the same name
(reference architecture/autoencoder_kl_wan.py:955-1419): same constructor config, same diffusers
state-dict keys, ``encode(x).latent_dist.mode()`` / ``decode(z, return_dict=False)[0]`` as the FrameINO Wan pipeline calls
them (pipelines/pipeline_wan_i2v_motion_FrameINO.py:464, :478, :502, :926), ``.config`` / ``.dtype`` as it reads them.
Only the Wan2.2 form (``is_residual=True``, the TI2V-5B VAE) is built; tiling / slicing are memory work-arounds the
180 GB part does not need and raise.

How it runs (B200-first, not a translation of the reference's NCTHW / cuDNN path):
  * activations are CHANNELS-LAST bf16 ``[T, H, W, C]``; every 3x3x3 / 3x3 / 3x1x1 convolution is an implicit GEMM on
    the tcgen05 tensor cores (``fino_conv3d_cl_bf16``: TMA box loads per filter tap, zero padding by out-of-bounds fill,
    fp32 accumulation in TMEM, bias and the skip connection in the epilogue); 1x1x1 convolutions are plain GEMMs;
  * WanRMS_norm + SiLU is one row kernel that writes straight into the next convolution's input buffer, behind the two
    cached frames of the reference's ``feat_cache`` (:169-176) — the causal history is a prefix of the same buffer, so
    the convolution is a plain "valid" one along time and nothing is concatenated or padded;
  * the chunking is the reference's (decode: one latent frame at a time, :1208-1216; encode: 1 + 4k frames, :1155-1166),
    so the caches hold exactly what the reference's hold;
  * weights stay fp32 in the state dict (the reference loads its VAE in fp32, app.py:157) and are packed once to bf16
    ``[C_out, taps * C_in(padded to 64)]`` K-major matrices; accumulation, norms and the softmax are fp32.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .modules import ModelBase
from .processors import tensor_key


def _up8(n: int) -> int:
    return (n + 7) // 8 * 8


# ---- parameter holders (diffusers key layout) -----------------------------------------------------------------------
class ConvParams(nn.Module):
    """``weight`` [C_out, C_in, *kernel] + ``bias`` of a WanCausalConv3d / nn.Conv2d; packed bf16 copies for the kernels."""

    def __init__(self, c_in: int, c_out: int, kernel: Tuple[int, ...]):
        super().__init__()
        self.c_in, self.c_out, self.kernel = c_in, c_out, tuple(kernel)
        self.weight = nn.Parameter(torch.empty(c_out, c_in, *kernel))
        self.bias = nn.Parameter(torch.empty(c_out))

    @property
    def kernel3(self) -> Tuple[int, int, int]:
        return self.kernel if len(self.kernel) == 3 else (1, *self.kernel)

    def packed(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(W bf16 [up8(C_out), taps * ceil(C_in/64)*64] tap-major with zero padding, bias bf16 [up8(C_out)])."""
        key = (tensor_key(self.weight), tensor_key(self.bias))
        hit = self.__dict__.get("_fino_cache")
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            w = self.weight.detach().float()
            w = w.reshape(self.c_out, self.c_in, -1).permute(0, 2, 1)  # [C_out, taps, C_in]
            taps = w.shape[1]
            cinp = (self.c_in + 63) // 64 * 64
            wp = torch.zeros(_up8(self.c_out), taps, cinp, dtype=torch.bfloat16, device=w.device)
            wp[: self.c_out, :, : self.c_in] = w.to(torch.bfloat16)
            bp = torch.zeros(_up8(self.c_out), dtype=torch.bfloat16, device=w.device)
            bp[: self.c_out] = self.bias.detach().to(torch.bfloat16)
            wp = wp.reshape(wp.shape[0], taps * cinp).contiguous()
        self.__dict__["_fino_cache"] = (key, wp, bp)
        return wp, bp

    def dense(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """1x1x1 / 1x1 convolution as a GEMM weight: (W bf16 [up8(C_out), up8(C_in)], bias bf16)."""
        key = ("dense", tensor_key(self.weight), tensor_key(self.bias))
        hit = self.__dict__.get("_fino_cache")
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            w = self.weight.detach().reshape(self.c_out, self.c_in)
            wp = torch.zeros(_up8(self.c_out), _up8(self.c_in), dtype=torch.bfloat16, device=w.device)
            wp[: self.c_out, : self.c_in] = w.to(torch.bfloat16)
            bp = torch.zeros(_up8(self.c_out), dtype=torch.bfloat16, device=w.device)
            bp[: self.c_out] = self.bias.detach().to(torch.bfloat16)
        self.__dict__["_fino_cache"] = (key, wp, bp)
        return wp, bp


class RMSNormParams(nn.Module):
    """``gamma`` of WanRMS_norm (:191-199): shape (dim, 1, 1, 1) for video tensors, (dim, 1, 1) for images."""

    def __init__(self, dim: int, images: bool):
        super().__init__()
        self.dim = dim
        self.gamma = nn.Parameter(torch.ones(dim, *((1, 1) if images else (1, 1, 1))))

    def gamma32(self) -> torch.Tensor:
        key = tensor_key(self.gamma)
        hit = self.__dict__.get("_fino_cache")
        if hit is None or hit[0] != key:
            hit = self.__dict__["_fino_cache"] = (key, self.gamma.detach().float().reshape(-1).contiguous())
        return hit[1]


class ResidualBlock(nn.Module):  # :311-382
    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        self.norm1 = RMSNormParams(c_in, images=False)
        self.conv1 = ConvParams(c_in, c_out, (3, 3, 3))
        self.norm2 = RMSNormParams(c_out, images=False)
        self.conv2 = ConvParams(c_out, c_out, (3, 3, 3))
        self.conv_shortcut = ConvParams(c_in, c_out, (1, 1, 1)) if c_in != c_out else None


class AttentionBlock(nn.Module):  # :385-427
    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        self.norm = RMSNormParams(dim, images=True)
        self.to_qkv = ConvParams(dim, 3 * dim, (1, 1))
        self.proj = ConvParams(dim, dim, (1, 1))


class MidBlock(nn.Module):  # :430-466
    def __init__(self, dim: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResidualBlock(dim, dim), ResidualBlock(dim, dim)])
        self.attentions = nn.ModuleList([AttentionBlock(dim)])


class Resample(nn.Module):  # :220-311
    def __init__(self, dim: int, mode: str, out_dim: int):
        super().__init__()
        self.mode = mode
        self.resample = nn.Sequential(nn.Identity(), ConvParams(dim, out_dim, (3, 3)))  # key "resample.1.*"
        if mode == "upsample3d":
            self.time_conv = ConvParams(dim, 2 * dim, (3, 1, 1))
        elif mode == "downsample3d":
            self.time_conv = ConvParams(dim, dim, (3, 1, 1))


class ResidualUpBlock(nn.Module):  # :626-711
    def __init__(self, c_in: int, c_out: int, num_res_blocks: int, temporal: bool, up_flag: bool):
        super().__init__()
        self.c_in, self.c_out, self.temporal, self.up_flag = c_in, c_out, temporal, up_flag
        self.resnets = nn.ModuleList([ResidualBlock(c_in if j == 0 else c_out, c_out) for j in range(num_res_blocks + 1)])
        self.upsampler = Resample(c_out, "upsample3d" if temporal else "upsample2d", c_out) if up_flag else None


class ResidualDownBlock(nn.Module):  # :469-502
    def __init__(self, c_in: int, c_out: int, num_res_blocks: int, temporal: bool, down_flag: bool):
        super().__init__()
        self.c_in, self.c_out, self.temporal, self.down_flag = c_in, c_out, temporal, down_flag
        self.resnets = nn.ModuleList([ResidualBlock(c_in if j == 0 else c_out, c_out) for j in range(num_res_blocks)])
        self.downsampler = Resample(c_out, "downsample3d" if temporal else "downsample2d", c_out) if down_flag else None


class Encoder3d(nn.Module):  # :505-623
    def __init__(self, in_channels: int, dim: int, z_dim: int, dim_mult, num_res_blocks: int, temperal_downsample):
        super().__init__()
        dims = [dim * u for u in [1] + list(dim_mult)]
        n = len(dim_mult)
        self.conv_in = ConvParams(in_channels, dims[0], (3, 3, 3))
        self.down_blocks = nn.ModuleList([
            ResidualDownBlock(ci, co, num_res_blocks, temperal_downsample[i] if i != n - 1 else False, i != n - 1)
            for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:]))])
        self.mid_block = MidBlock(dims[-1])
        self.norm_out = RMSNormParams(dims[-1], images=False)
        self.conv_out = ConvParams(dims[-1], z_dim, (3, 3, 3))


class Decoder3d(nn.Module):  # :783-909
    def __init__(self, dim: int, z_dim: int, dim_mult, num_res_blocks: int, temperal_upsample, out_channels: int):
        super().__init__()
        mult = list(dim_mult)
        dims = [dim * u for u in [mult[-1]] + mult[::-1]]
        n = len(mult)
        self.conv_in = ConvParams(z_dim, dims[0], (3, 3, 3))
        self.mid_block = MidBlock(dims[0])
        self.up_blocks = nn.ModuleList([
            ResidualUpBlock(ci, co, num_res_blocks, temperal_upsample[i] if i != n - 1 else False, i != n - 1)
            for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:]))])
        self.norm_out = RMSNormParams(dims[-1], images=False)
        self.conv_out = ConvParams(dims[-1], out_channels, (3, 3, 3))


# ---- multi-GPU: the frame rows of a decode split across the ranks of a process group -------------------------------
class RowParallel:
    """Every rank decodes a horizontal band of the frames (latent rows ``rows(h)``, the same band scaled by 2 after each
    up-sampling stage). Everything in the decoder is pixel-local except (i) the 3x3 spatial taps of the convolutions —
    each convolution's input buffer carries one halo row above and below the band, refreshed from the neighbouring
    ranks right after the producer has written the band (``exchange``: ONE launch per convolution that pushes the two
    edge rows into the neighbours' peer-mapped mailboxes over NVLink, raises their flags and pulls its own two halo
    rows, ``fino_halo_exchange``; an all-gather of the edge rows where peer memory is not available) — and (ii) the
    mid-block attention, whose keys / values are the whole frame
    (``gather_rows`` of the 1024-channel latent-resolution activations, 7 MB per frame at 44 x 80). Bands at the image
    border keep their outer halo row zero: that IS the convolution's zero padding. The arithmetic per output pixel is
    the un-sharded one, in the same order — the result is bit-identical."""

    SLOT_BYTES = 4 << 20  # one mailbox slot: t x W x C bf16 of one row (1.3 MB at 704x1280x121's widest stage)

    def __init__(self, group=None, peer: bool = True, prims=ops):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("RowParallel needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.prims = prims
        self.seq = 0
        self.own = self.up = self.down = None
        # halo rows travel through peer memory (one fused launch per convolution, fino_halo_exchange) when every rank
        # is a CUDA device of this node; otherwise — the gloo host-logic tests — through an all-gather of the edge rows
        self.peer = bool(peer and self.world > 1 and torch.cuda.is_available() and dist.get_backend(group) == "nccl")
        if self.peer:
            self._open_mailboxes()

    def _open_mailboxes(self) -> None:
        import torch.distributed as dist

        ops_ = self.prims
        self.own = ops_.peer_alloc(ops_.HALO_DATA_OFF + 4 * self.SLOT_BYTES)
        handles = [None] * self.world
        dist.all_gather_object(handles, ops_.peer_export(self.own), group=self.group)
        self._imported = []
        if self.rank > 0:
            self.up = ops_.peer_import(handles[self.rank - 1])
            self._imported.append(self.up)
        if self.rank < self.world - 1:
            self.down = ops_.peer_import(handles[self.rank + 1])
            self._imported.append(self.down)
        dist.barrier(group=self.group)  # every mailbox is mapped before anyone pushes into one

    def close(self) -> None:
        """Unmaps the neighbours' mailboxes and frees this rank's (collective: every rank calls it)."""
        import torch.distributed as dist

        if self.own is None:
            return
        torch.cuda.synchronize()
        words = self.prims.tensor_from_ptr(self.own, (8,), torch.int32).cpu().tolist()
        dist.barrier(group=self.group)
        for p in self._imported:
            self.prims.peer_release(p)
        self._imported = []
        dist.barrier(group=self.group)
        self.prims.peer_free(self.own)
        self.own = self.up = self.down = None
        self.peer = False
        if words[4] or words[5]:
            raise RuntimeError(f"rank {self.rank}: a halo exchange timed out (exchange {words[4] or words[5]}); results "
                               f"after that are invalid")

    @staticmethod
    def split(h: int, world: int) -> List[Tuple[int, int]]:
        """Row ranges [a, b) per rank: the first ``h % world`` ranks hold one row more."""
        if h < world:
            raise ValueError(f"cannot split {h} latent rows over {world} ranks")
        base, extra = divmod(h, world)
        out, a = [], 0
        for r in range(world):
            b = a + base + (1 if r < extra else 0)
            out.append((a, b))
            a = b
        return out

    def rows(self, h: int) -> Tuple[int, int]:
        return self.split(h, self.world)[self.rank]

    def exchange(self, frames: torch.Tensor) -> None:
        """frames: [t, h_loc + 2, W, C] (rows 1..h_loc written): fills row 0 from the rank above, row h_loc + 1 from the
        rank below."""
        import torch.distributed as dist

        if self.world == 1:
            return
        t, hp, w, c = frames.shape
        hl = hp - 2
        if self.peer and frames.is_cuda and t * w * c * 2 <= self.SLOT_BYTES:
            self.seq += 1
            self.prims.halo_exchange(frames, self.own, self.up, self.down, self.seq, self.SLOT_BYTES, self.rank)
            return
        edge = torch.stack((frames[:, 1], frames[:, hl]))  # [2, t, W, C]: my first and last rows
        flat = torch.empty((self.world * 2,) + tuple(edge.shape[1:]), dtype=edge.dtype, device=edge.device)
        dist.all_gather_into_tensor(flat, edge, group=self.group)
        allb = flat.view((self.world,) + tuple(edge.shape))
        if self.rank > 0:
            frames[:, 0].copy_(allb[self.rank - 1, 1])
        if self.rank < self.world - 1:
            frames[:, hl + 1].copy_(allb[self.rank + 1, 0])

    def gather_rows(self, x: torch.Tensor, h_total: int, dim: int = 1) -> torch.Tensor:
        """The bands of all ranks concatenated along ``dim`` (the row axis): every rank gets the whole frames."""
        import torch.distributed as dist

        if self.world == 1:
            return x
        sizes = [b - a for a, b in self.split(h_total, self.world)]
        scale = x.shape[dim] // sizes[self.rank]  # bands grow by 2 per up-sampling stage (and by the patch size)
        assert x.shape[dim] == sizes[self.rank] * scale, "band height does not match the row split"
        hmax = max(sizes) * scale
        if x.shape[dim] != hmax:
            pad_shape = list(x.shape)
            pad_shape[dim] = hmax - x.shape[dim]
            x = torch.cat([x, x.new_zeros(pad_shape)], dim=dim)
        x = x.contiguous()
        flat = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(flat, x, group=self.group)
        allb = flat.view((self.world,) + tuple(x.shape))
        return torch.cat([allb[r].narrow(dim, 0, sizes[r] * scale) for r in range(self.world)], dim=dim)


# ---- run-time state: the causal history of every convolution (the reference's feat_cache) ---------------------------
class _ConvCaches:
    """Per-convolution input buffers: a strip of frames ``[capacity, H, W, C]`` in which every chunk's input is written
    right behind the previous chunk's, so the last ``hist`` frames of the previous chunk ARE the causal history of the
    next one (the reference's ``feat_cache``, :360-367) without moving them; only when the strip is full are the last
    ``hist`` frames copied back to its start (every ``ROOM`` chunks). Frames before the first chunk are zeros.
    With ``halo = 1`` (row-parallel decode) the strips of the spatial convolutions hold one extra row above and below
    every frame (``input`` returns the band, ``frames`` / ``window`` the band with its halo rows)."""

    ROOM = 4  # chunks of the current length a strip holds before it wraps

    def __init__(self, device, halo: int = 0):
        self.device = device
        self.halo = halo
        self.bufs: Dict[int, torch.Tensor] = {}
        self.pos: Dict[int, int] = {}  # first frame of the current chunk's input inside the strip
        self.pad: Dict[int, int] = {}
        self.scratch: Dict[tuple, torch.Tensor] = {}

    def input(self, conv: ConvParams, t: int, h: int, w: int, hist: int = 2, spatial: bool = True) -> torch.Tensor:
        c = _up8(conv.c_in)
        key = id(conv)
        pad = self.halo if spatial else 0
        hp = h + 2 * pad
        buf = self.bufs.get(key)
        if buf is None:
            buf = self.bufs[key] = torch.zeros(hist + self.ROOM * t, hp, w, c, dtype=torch.bfloat16, device=self.device)
            self.pos[key] = hist
            self.pad[key] = pad
        assert tuple(buf.shape[1:]) == (hp, w, c) and self.pad[key] == pad, "canvas changed inside one encode / decode"
        pos = self.pos[key]
        if hist + self.ROOM * t > buf.shape[0]:  # the chunks got longer (first chunk: 1 frame, later 2 or 4): grow
            new = torch.zeros(hist + self.ROOM * t, hp, w, c, dtype=torch.bfloat16, device=self.device)
            new[:hist].copy_(buf[pos - hist:pos])
            buf = self.bufs[key] = new
            pos = self.pos[key] = hist
        elif pos + t > buf.shape[0]:  # strip full: the history goes back to the start
            for i in range(hist):  # ascending: frame pos - hist + i >= i, never overwritten before it is read
                buf[i].copy_(buf[pos - hist + i])
            pos = self.pos[key] = hist
        return buf[pos:pos + t, pad:pad + h] if pad else buf[pos:pos + t]

    def frames(self, conv: ConvParams, t: int) -> torch.Tensor:
        """The current chunk's frames with their halo rows."""
        pos = self.pos[id(conv)]
        return self.bufs[id(conv)][pos:pos + t]

    def window(self, conv: ConvParams, t: int, hist: int = 2) -> torch.Tensor:
        pos = self.pos[id(conv)]
        return self.bufs[id(conv)][pos - hist:pos + t]

    def advance(self, conv: ConvParams, t: int, hist: int = 2) -> None:
        """The chunk's frames become history: the next chunk is written right behind them."""
        self.pos[id(conv)] += t

    def set_history(self, conv: ConvParams, frame: torch.Tensor, hist: int = 1) -> None:
        """Overwrites the most recent history frame (downsample3d's first chunk, :303-305)."""
        self.bufs[id(conv)][self.pos[id(conv)] - 1].copy_(frame)

    def padded(self, conv: ConvParams, t: int, h: int, w: int, c: int) -> torch.Tensor:
        """A zero-initialised [t, h + 2 halo, w, c] buffer kept per (convolution, t): the input of a 2-D convolution."""
        key = (id(conv), t, h, w, c)
        buf = self.scratch.get(key)
        if buf is None:
            buf = self.scratch[key] = torch.zeros(t, h + 2 * self.halo, w, c, dtype=torch.bfloat16, device=self.device)
        return buf


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class DiagonalGaussianDistribution:
    """diffusers.models.autoencoders.vae.DiagonalGaussianDistribution (upstream): the FrameINO pipeline only takes
    ``mode()`` (``retrieve_latents(..., sample_mode="argmax")``, pipeline :464)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator=None) -> torch.Tensor:
        std = torch.exp(0.5 * torch.clamp(self.logvar, -30.0, 20.0))
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + std * noise


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


class AutoencoderKLWan(ModelBase):
    """Drop-in for reference ``architecture.autoencoder_kl_wan.AutoencoderKLWan`` (see the module docstring)."""

    _supports_gradient_checkpointing = False
    _stores_any_dtype = True  # parameters keep the requested dtype (the reference loads the VAE in fp32, app.py:157);
                              # the kernels read bf16 copies packed by prepare()

    def __init__(self, base_dim: int = 96, decoder_base_dim: Optional[int] = None, z_dim: int = 16,
                 dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2, attn_scales=(), temperal_downsample=(False, True, True),
                 dropout: float = 0.0, latents_mean=None, latents_std=None, is_residual: bool = False,
                 in_channels: int = 3, out_channels: int = 3, patch_size: Optional[int] = None,
                 scale_factor_temporal: Optional[int] = 4, scale_factor_spatial: Optional[int] = 8) -> None:
        super().__init__()
        if not is_residual:
            raise NotImplementedError("only the Wan2.2 VAE (is_residual=True, WanResidualDown/UpBlock) is built: the "
                                      "FrameINO Wan2.2-TI2V-5B pipeline's VAE")
        if list(attn_scales):
            raise NotImplementedError("attn_scales: the Wan2.2 VAE has attention in the mid blocks only")
        self._register_config(
            base_dim=base_dim, decoder_base_dim=decoder_base_dim, z_dim=z_dim, dim_mult=list(dim_mult),
            num_res_blocks=num_res_blocks, attn_scales=list(attn_scales), temperal_downsample=list(temperal_downsample),
            dropout=dropout, latents_mean=latents_mean, latents_std=latents_std, is_residual=is_residual,
            in_channels=in_channels, out_channels=out_channels, patch_size=patch_size,
            scale_factor_temporal=scale_factor_temporal, scale_factor_spatial=scale_factor_spatial)
        self.z_dim = z_dim
        self.temperal_downsample = list(temperal_downsample)
        self.temperal_upsample = self.temperal_downsample[::-1]
        self.encoder = Encoder3d(in_channels, base_dim, z_dim * 2, dim_mult, num_res_blocks, self.temperal_downsample)
        self.quant_conv = ConvParams(z_dim * 2, z_dim * 2, (1, 1, 1))
        self.post_quant_conv = ConvParams(z_dim, z_dim, (1, 1, 1))
        self.decoder = Decoder3d(decoder_base_dim or base_dim, z_dim, dim_mult, num_res_blocks, self.temperal_upsample,
                                 out_channels)
        self.spatial_compression_ratio = 2 ** len(self.temperal_downsample)
        self.use_slicing = False
        self.use_tiling = False
        self.row_parallel: Optional[RowParallel] = None

    # memory work-arounds of the reference (:1084-1133). Tiling cannot be mirrored for this VAE because the reference's
    # own tiled path does not run for it: tiled_encode (:1301-1312) feeds the raw 3-channel tile to an encoder whose
    # conv_in expects the 12 patchified channels (patchify happens after the tiling branch, :1148-1153), and
    # tiled_decode (:1373-1397) returns the 12-channel tiles without unpatchify / clamp. No Wan caller enables it
    # (only the CogVideoX scripts call vae.enable_tiling()).
    def enable_tiling(self, *a, **k):
        raise NotImplementedError("tiled encode / decode: the reference's tiled path fails for the Wan2.2 (patch_size 2) "
                                  "VAE (un-patchified tiles into a 12-channel conv_in, autoencoder_kl_wan.py:1148-1153, "
                                  ":1301-1312); the whole-frame path peaks at 32 GB at 704x1280x121, which fits")

    def disable_tiling(self):
        self.use_tiling = False

    def enable_row_parallel(self, group=None, peer: bool = True) -> "RowParallel":
        """``decode`` and ``encode`` split the frame rows over the ranks of ``group``: every rank calls them with the
        same input and gets the whole result, bit-identical to the un-sharded one."""
        self.disable_row_parallel()
        self.row_parallel = RowParallel(group, peer=peer)
        return self.row_parallel

    def disable_row_parallel(self) -> None:
        if self.row_parallel is not None:
            self.row_parallel.close()
        self.row_parallel = None

    def _row_parallel_for(self, latent_rows: int) -> Optional["RowParallel"]:
        """The row split to use for a canvas, or None: one rank, or fewer latent rows than ranks (a band needs at least
        one row) — every rank then runs the whole frames, which is the same result."""
        rp = self.row_parallel
        if rp is None or rp.world == 1 or latent_rows < rp.world:
            return None
        return rp

    def enable_slicing(self):
        self.use_slicing = True  # batch elements are processed one at a time anyway

    def disable_slicing(self):
        self.use_slicing = False

    def prepare(self) -> "AutoencoderKLWan":
        """Packs every convolution's weights (bf16, tap-major, K padded) ahead of the first call."""
        for m in self.modules():
            if isinstance(m, ConvParams):
                if all(k == 1 for k in m.kernel):
                    m.dense()
                else:
                    m.packed()
            elif isinstance(m, RMSNormParams):
                m.gamma32()
        return self

    # ---- building blocks on channels-last chunks [t, H, W, C] -------------------------------------------------------
    @staticmethod
    def _conv1x1(conv: ConvParams, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        w, b = conv.dense()
        t, h, wd, c = x.shape
        x2 = x.reshape(t * h * wd, c)
        if residual is not None:
            y = ops.linear(x2, w, b, epilogue=ops.EPI_GATE_RESIDUAL, residual=residual.reshape(t * h * wd, -1))
        else:
            y = ops.linear(x2, w, b)
        return y.view(t, h, wd, w.shape[0])

    def _causal(self, caches: _ConvCaches, conv: ConvParams, t: int, h: int, w: int,
                residual: Optional[torch.Tensor] = None, exchange: bool = True) -> torch.Tensor:
        """Runs a 3x3x3 causal convolution on the frames its producer has written into ``caches.input(conv, ...)``.
        Row-parallel: the band's halo rows come from the neighbouring ranks first, and the convolution is "valid"
        along H over the band + halo."""
        wp, bp = conv.packed()
        if caches.halo:
            if exchange:
                self.row_parallel.exchange(caches.frames(conv, t))
            y = ops.conv3d_cl(caches.window(conv, t), wp, bp, conv.kernel3, pad_hw=(0, 1), residual=residual)
        else:
            y = ops.conv3d_cl(caches.window(conv, t), wp, bp, conv.kernel3, pad_hw=(1, 1), residual=residual)
        caches.advance(conv, t)
        return y

    @staticmethod
    def _rms_into(x: torch.Tensor, gamma: torch.Tensor, dst: torch.Tensor) -> None:
        """WanRMS_norm + SiLU of x [t, h, w, c] into a convolution's input frames (frame-strided when they carry halo
        rows: one launch per frame then)."""
        if dst.is_contiguous():
            ops.rms_act_cl(x, gamma, silu=True, out=dst)
        else:
            for f in range(x.shape[0]):
                ops.rms_act_cl(x[f], gamma, silu=True, out=dst[f])

    def _res_block(self, caches: _ConvCaches, blk: ResidualBlock, x: torch.Tensor) -> torch.Tensor:
        """WanResidualBlock.forward (:342-382)."""
        t, h, w, _ = x.shape
        skip = x if blk.conv_shortcut is None else self._conv1x1(blk.conv_shortcut, x)
        self._rms_into(x, blk.norm1.gamma32(), caches.input(blk.conv1, t, h, w))
        y = self._causal(caches, blk.conv1, t, h, w)
        self._rms_into(y, blk.norm2.gamma32(), caches.input(blk.conv2, t, h, w))
        return self._causal(caches, blk.conv2, t, h, w, residual=skip)

    def _attention(self, blk: AttentionBlock, x: torch.Tensor, rp: Optional[RowParallel] = None) -> torch.Tensor:
        """WanAttentionBlock.forward (:402-427): per frame, one head of width C over the H*W pixels. Q K^T, the row
        softmax and P V are two tcgen05 GEMMs around a softmax kernel (head_dim = C is far beyond a flash tile); V is
        produced already transposed (V^T = W_v X^T) and its bias is added after P V (softmax rows sum to one)."""
        t, h, w, c = x.shape
        nq = h * w  # this rank's query pixels (the whole frame unless row-parallel: rp is set by a sharded decode only)
        wqkv, bqkv = blk.to_qkv.dense()
        wp, bp = blk.proj.dense()
        out = torch.empty_like(x)
        xn = ops.rms_act_cl(x, blk.norm.gamma32(), silu=False)
        # keys / values are the whole frame: gathered from all bands when the rows are split
        xkv = xn if rp is None else rp.gather_rows(xn, self._latent_rows)
        n = xkv.shape[1] * w
        if n % 8 != 0:
            raise NotImplementedError(f"VAE attention: {xkv.shape[1]} x {w} latent pixels per frame must be a multiple "
                                      f"of 8")
        npad = _up8(n)
        scores = torch.empty(nq, n, dtype=torch.float32, device=x.device)
        probs = torch.empty(nq, npad, dtype=torch.bfloat16, device=x.device)
        v_t = torch.zeros(c, npad, dtype=torch.bfloat16, device=x.device)
        for f in range(t):
            xf = xkv[f].reshape(n, c)
            if rp is None:
                qk = ops.linear(xf, wqkv[: 2 * c], bqkv[: 2 * c])  # [n, 2C]
                q, k = qk[:, :c], qk[:, c:]
            else:
                q = ops.linear(xn[f].reshape(nq, c), wqkv[:c], bqkv[:c].contiguous())
                k = ops.linear(xf, wqkv[c: 2 * c], bqkv[c: 2 * c].contiguous())
            ops.linear(wqkv[2 * c: 3 * c], xf, None, out=v_t[:, :n])  # V^T without bias: [C, n]
            ops.linear(q, k, None, out=scores)  # Q K^T, fp32
            ops.softmax_rows(scores, float(c) ** -0.5, probs)
            o = ops.linear(probs, v_t, bqkv[2 * c: 3 * c].contiguous())  # P V + b_v: [nq, C]
            ops.linear(o, wp, bp, epilogue=ops.EPI_GATE_RESIDUAL, residual=x[f].reshape(nq, c),
                       out=out[f].reshape(nq, c))
        return out

    def _mid(self, caches: _ConvCaches, mid: MidBlock, x: torch.Tensor) -> torch.Tensor:
        x = self._res_block(caches, mid.resnets[0], x)
        x = self._attention(mid.attentions[0], x, self.row_parallel if caches.halo else None)
        return self._res_block(caches, mid.resnets[1], x)

    def _upsample(self, caches: _ConvCaches, up: Resample, x: torch.Tensor, first_chunk: bool) -> torch.Tensor:
        """WanResample.forward, upsample2d / upsample3d (:265-299)."""
        t, h, w, c = x.shape
        if up.mode == "upsample3d" and not first_chunk:  # the first chunk only marks the cache ("Rep", :269-271)
            tc = up.time_conv
            caches.input(tc, t, h, w, spatial=False).copy_(x)
            wp, bp = tc.packed()
            y = torch.empty(2 * t, h, w, c, dtype=x.dtype, device=x.device)
            win = caches.window(tc, t)
            for j in range(2):  # channel half j of time_conv's output is frame 2 tau + j (:291-293)
                ops.conv3d_cl(win, wp[j * c:(j + 1) * c], bp[j * c:(j + 1) * c], (3, 1, 1), out=y[j::2])
            caches.advance(tc, t)
            x = y
        conv = up.resample[1]
        wp, bp = conv.packed()
        if caches.halo:  # the up-sampled band goes between the halo rows of a kept buffer, frame by frame
            t2, h2, w2 = x.shape[0], 2 * x.shape[1], 2 * x.shape[2]
            buf = caches.padded(conv, t2, h2, w2, c)
            for f in range(t2):
                ops.upsample2x_cl(x[f:f + 1], out=buf[f:f + 1, 1:1 + h2])
            self.row_parallel.exchange(buf)
            return ops.conv3d_cl(buf, wp, bp, (1, 3, 3), pad_hw=(0, 1))
        return ops.conv3d_cl(ops.upsample2x_cl(x), wp, bp, (1, 3, 3), pad_hw=(1, 1))

    def _downsample(self, caches: _ConvCaches, ds: Resample, x: torch.Tensor, first_chunk: bool) -> torch.Tensor:
        """WanResample.forward, downsample2d / downsample3d (:296-311)."""
        t, h, w, c = x.shape
        conv = ds.resample[1]
        wp, bp = conv.packed()
        # ZeroPad2d((0, 1, 0, 1)) + Conv2d(stride 2): the pad row / column is the TMA out-of-bounds fill
        if caches.halo:  # row-parallel: output row r reads rows 2r .. 2r+2, i.e. one row of the band below
            buf = caches.padded(conv, t, h, w, c)
            buf[:, 1:1 + h].copy_(x)
            self.row_parallel.exchange(buf)
            x = ops.conv3d_cl(buf[:, 1:], wp, bp, (1, 3, 3), pad_hw=(0, 0), stride_hw=2, out_hw=(h // 2, w // 2))
        else:
            x = ops.conv3d_cl(x, wp, bp, (1, 3, 3), pad_hw=(0, 0), stride_hw=2, out_hw=(h // 2, w // 2))
        if ds.mode == "downsample3d":
            tc = ds.time_conv
            h2, w2 = h // 2, w // 2
            if first_chunk:  # :303-305: the first frame passes through and becomes the cache
                caches.input(tc, t, h2, w2, hist=1, spatial=False)
                caches.set_history(tc, x[t - 1])
            else:
                caches.input(tc, t, h2, w2, hist=1, spatial=False).copy_(x)
                wt, bt = tc.packed()
                y = ops.conv3d_cl(caches.window(tc, t, hist=1), wt, bt, (3, 1, 1), stride_t=2)
                caches.advance(tc, t, hist=1)
                x = y
        return x

    def _conv_in_band(self, caches: _ConvCaches, conv: ConvParams, x: torch.Tensor) -> torch.Tensor:
        """Row-parallel first convolution: x [t, H, W, C] holds WHOLE frames (the input is on every rank), so this rank's
        band — the latent split scaled to x's resolution — and its two halo rows are copied from it directly."""
        t, hfull, w, _ = x.shape
        sc = hfull // self._latent_rows
        a, b = self.row_parallel.rows(self._latent_rows)
        a, b = a * sc, b * sc
        h = b - a
        caches.input(conv, t, h, w)
        lo, hi = max(a - 1, 0), min(b + 1, hfull)
        caches.frames(conv, t)[:, lo - (a - 1):lo - (a - 1) + (hi - lo)].copy_(x[:, lo:hi])
        return self._causal(caches, conv, t, h, w, exchange=False)

    # ---- decode ------------------------------------------------------------------------------------------------------
    def _decode_chunk(self, caches: _ConvCaches, x: torch.Tensor, first_chunk: bool, taps: Optional[dict]) -> torch.Tensor:
        """WanDecoder3d.forward (:874-909) on one latent frame; returns channels-last [t_out, H, W, up8(out_channels)]."""
        dec = self.decoder
        if caches.halo:  # x is the whole latent frame: the band AND its halo rows are at hand, nothing to exchange
            x = self._conv_in_band(caches, dec.conv_in, x)
        else:
            t, h, w, _ = x.shape
            caches.input(dec.conv_in, t, h, w).copy_(x)
            x = self._causal(caches, dec.conv_in, t, h, w)
        x = self._mid(caches, dec.mid_block, x)
        if taps is not None:
            taps.setdefault("mid", []).append(x.clone())
        for i, blk in enumerate(dec.up_blocks):
            x_copy = x
            for res in blk.resnets:
                x = self._res_block(caches, res, x)
            if blk.upsampler is not None:
                x = self._upsample(caches, blk.upsampler, x, first_chunk)
                ops.dupup_add_cl(x, x_copy, 2 if blk.temporal else 1, 2, first_chunk)  # :709-710
            if taps is not None:
                taps.setdefault(f"up{i}", []).append(x.clone())
        t, h, w, _ = x.shape
        self._rms_into(x, dec.norm_out.gamma32(), caches.input(dec.conv_out, t, h, w))
        y = self._causal(caches, dec.conv_out, t, h, w)
        if taps is not None:
            taps.setdefault("head", []).append(y[..., : dec.conv_out.c_out].clone())
        return y

    def _check(self, x: torch.Tensor) -> None:
        if not x.is_cuda:
            raise RuntimeError("frameino_b200 has no CPU path: move the model and inputs to a CUDA device")
        if x.dim() != 5:
            raise ValueError(f"expected a [B, C, T, H, W] tensor, got {tuple(x.shape)}")

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, output_dtype: Optional[torch.dtype] = None):
        """``AutoencoderKLWan.decode`` (:1230-1252 -> _decode :1198-1228). z: [B, z_dim, T, h, w] (fp32 or bf16) ->
        video [B, 3, 1 + 4 (T - 1), h * s, w * s] in [-1, 1], in z's dtype unless ``output_dtype`` says otherwise."""
        self._check(z)
        cfg = self.config
        b, zc, tl, h, w = z.shape
        if zc != cfg.z_dim:
            raise ValueError(f"latent has {zc} channels, the VAE z_dim is {cfg.z_dim}")
        ps = cfg.patch_size or 1
        c_img = cfg.out_channels // (ps * ps)
        n_up = sum(1 for u in self.decoder.up_blocks if u.upsampler is not None)
        n_tup = sum(1 for u in self.decoder.up_blocks if u.upsampler is not None and u.temporal)
        ho, wo = h * (2 ** n_up) * ps, w * (2 ** n_up) * ps
        t_total = 1 + (tl - 1) * (2 ** n_tup)
        dt = output_dtype or (z.dtype if z.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        rp = self._row_parallel_for(h)
        taps = self.__dict__.get("_fino_taps")
        if rp is not None:  # this rank's band of every frame; gathered at the end
            if taps is not None:
                raise NotImplementedError("stage taps are an un-sharded debugging aid")
            a, b_ = rp.rows(h)
            self._latent_rows = h
            ho_full, ho = ho, (b_ - a) * (2 ** n_up) * ps
        out = torch.empty(b, c_img, t_total, ho, wo, dtype=dt, device=z.device)
        zin = z if z.dtype in (torch.float32, torch.bfloat16) else z.float()
        for bi in range(b):
            caches = _ConvCaches(z.device, halo=0 if rp is None else 1)
            x_all = self._conv1x1(self.post_quant_conv, ops.vae_to_cl(zin[bi], 1, _up8(zc)))  # :1207
            f0 = 0
            for i in range(tl):  # :1208-1216
                y = self._decode_chunk(caches, x_all[i:i + 1], i == 0, taps if bi == 0 else None)
                ops.vae_from_cl(y, out[bi, :, f0:f0 + y.shape[0]], c_img, ps, clamp=True)  # :1218-1224
                f0 += y.shape[0]
            assert f0 == t_total
        if rp is not None:
            out = rp.gather_rows(out, h, dim=3)
            assert out.shape[3] == ho_full
        if not return_dict:
            return (out,)
        return DecoderOutput(sample=out)

    # ---- encode ------------------------------------------------------------------------------------------------------
    def _encode_chunk(self, caches: _ConvCaches, x: torch.Tensor, first_chunk: bool, taps: Optional[dict]) -> torch.Tensor:
        """WanEncoder3d.forward (:586-623) on one chunk (1 frame, then 4 at a time); returns [t', h, w, 2 z_dim]."""
        enc = self.encoder
        if caches.halo:  # x holds whole frames (the clip is on every rank): band + halo rows without an exchange
            x = self._conv_in_band(caches, enc.conv_in, x)
        else:
            t, h, w, _ = x.shape
            caches.input(enc.conv_in, t, h, w).copy_(x)
            x = self._causal(caches, enc.conv_in, t, h, w)
        for i, blk in enumerate(enc.down_blocks):
            x_copy = x
            for res in blk.resnets:
                x = self._res_block(caches, res, x)
            if blk.downsampler is not None:
                x = self._downsample(caches, blk.downsampler, x, first_chunk)
            ops.avgdown_add_cl(x, x_copy, 2 if blk.temporal else 1, 2 if blk.down_flag else 1)  # :502
            if taps is not None:
                taps.setdefault(f"down{i}", []).append(x.clone())
        x = self._mid(caches, enc.mid_block, x)
        t, h, w, _ = x.shape
        self._rms_into(x, enc.norm_out.gamma32(), caches.input(enc.conv_out, t, h, w))
        return self._causal(caches, enc.conv_out, t, h, w)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """``AutoencoderKLWan.encode`` (:1172-1196 -> _encode :1145-1170). x: [B, 3, 1 + 4k, H, W] in [-1, 1] ->
        ``latent_dist`` over [B, z_dim, 1 + k, H/s, W/s] (``.mode()`` is what the pipeline takes)."""
        self._check(x)
        cfg = self.config
        b, c, tf, h, w = x.shape
        ps = cfg.patch_size or 1
        if c * ps * ps != cfg.in_channels:
            raise ValueError(f"input has {c} channels x patch {ps}^2, the VAE expects {cfg.in_channels}")
        if (tf - 1) % 4 != 0:
            raise ValueError(f"the VAE encodes 1 + 4k frames, got {tf}")
        n_down = sum(1 for d in self.encoder.down_blocks if d.downsampler is not None)
        hl, wl = h // ps // (2 ** n_down), w // ps // (2 ** n_down)
        tl = 1 + (tf - 1) // 4
        xin = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        dt = xin.dtype
        rp = self._row_parallel_for(hl)
        taps = self.__dict__.get("_fino_taps")
        hl_loc = hl
        if rp is not None:  # this rank's band of the latent rows; gathered at the end
            if taps is not None:
                raise NotImplementedError("stage taps are an un-sharded debugging aid")
            a, b_ = rp.rows(hl)
            self._latent_rows = hl
            hl_loc = b_ - a
        params = torch.empty(b, 2 * cfg.z_dim, tl, hl_loc, wl, dtype=dt, device=x.device)
        for bi in range(b):
            caches = _ConvCaches(x.device, halo=0 if rp is None else 1)
            x_all = ops.vae_to_cl(xin[bi], ps, _up8(cfg.in_channels))  # patchify (:1152-1153) + channels-last
            f0 = 0
            for i in range(tl):  # :1155-1166
                lo, hi = (0, 1) if i == 0 else (1 + 4 * (i - 1), 1 + 4 * i)
                y = self._encode_chunk(caches, x_all[lo:hi], i == 0, taps if bi == 0 else None)
                y = self._conv1x1(self.quant_conv, y)  # :1168 (1x1x1: chunk-wise == on the concatenation)
                ops.vae_from_cl(y.contiguous(), params[bi, :, f0:f0 + y.shape[0]], 2 * cfg.z_dim, 1, clamp=False)
                f0 += y.shape[0]
            assert f0 == tl
        if rp is not None:
            params = rp.gather_rows(params, hl, dim=3)
        dist = DiagonalGaussianDistribution(params)
        if not return_dict:
            return (dist,)
        return AutoencoderKLOutput(latent_dist=dist)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True, generator=None):
        """:1399-1419"""
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        return self.decode(z, return_dict=return_dict)
