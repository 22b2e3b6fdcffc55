"""Noise-conditioned SFNO (``NoiseConditionedSFNO``) whose forward pass runs in libace_b200 (sm_100a CUDA).

Drop-in for the reference's conditional network as ``NoiseConditionedSFNOBuilder`` configures it:

* ``fme/core/models/conditional_sfno/layers.py:33-95``    ``ContextConfig`` / ``Context``
* ``fme/core/models/conditional_sfno/sfnonet.py:46-148``  ``SFNONetConfig`` (fields + defaults)
* ``fme/core/models/conditional_sfno/sfnonet.py:443-824`` ``get_lat_lon_sfnonet`` / ``SphericalFourierNeuralOperatorNet``
* ``fme/ace/registry/stochastic_sfno.py:21-172``          ``isotropic_noise`` / ``NoiseConditionedModel``

Same parameter names, shapes and creation order as the reference (``state_dict()`` round-trips with reference checkpoints,
including the two older dhconv weight layouts that ``SpectralConvS2`` upgrades on load, ``s2convolutions.py:282-357``;
construction under a seed draws the same initial weights).  The torch sub-modules only *hold* parameters; the device library
keeps its own re-laid-out copy, refreshed when a parameter's storage or version counter changes.  Inference only.

Random draws (the N(0,1) numbers behind the noise) come from ``torch.randn`` on the input's device, as the reference's come
from ``fme.core.rand.randn``; everything downstream (isotropic scaling, inverse SHT, the network) runs in the library.

Options that only re-parameterise the per-degree spectral operator or a 1x1 convolution are folded into the parameters the
device library receives (``_sync_params``; exact identities, evaluated in fp64 once per parameter change, never per step):

* ``filter_num_groups = G``      -> block-diagonal dense ``[L, C, C]`` operator (s2convolutions.py:229-236, :388);
* ``spectral_lora_rank``         -> ``W_l + (alpha / r) B_l A_l``                  (s2convolutions.py:94-115, :393-410);
* ``filter_preserves_global_mean`` -> ``W_0 = I``                                  (s2convolutions.py:411-418);
* ``spectral_ratio < 1``         -> ``Q W'_l P`` with the pre / post projections  (s2convolutions.py:210-228: channel-wise linear
  maps commute with the transforms; the transform then runs at full width, i.e. the option's speed-up is not realised);
* ``lora_rank`` (``LoRAConv2d``) -> ``W + (alpha / r) W_up W_down``                (lora.py:131-141).

``filter_residual`` / ``filter_output`` (SHT round trips of the residual streams / the output) and the eval branch of
``clip_latent_global_means`` (the latent's per-channel mean clamped into the ``_gm_min`` / ``_gm_max`` envelope buffers of the
checkpoint, sfnonet.py:792-812; the envelope itself is only updated by training, which this module does not do) run on the device.

Unsupported reference options raise ``NotImplementedError`` at construction (no fallback): ``filter_type`` other than
``"linear"``, ``global_layer_norm``, local (DISCO) blocks, ``spectral_ratio < 1`` together with a round-trip residual
(``filter_residual`` or a non-Gaussian data grid), ``use_mlp=False``, ``encoder_layers != 1``,
dropout, activation other than GELU.
"""
import ctypes
import dataclasses
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .sfno import trunc_normal_
from .sht import InverseRealSHT, RealSHT, ShtPlan


@dataclasses.dataclass
class ContextConfig:
    embed_dim_scalar: int = 0
    embed_dim_labels: int = 0
    embed_dim_noise: int = 0
    embed_dim_pos: int = 0


@dataclasses.dataclass
class Context:
    embedding_scalar: Optional[torch.Tensor] = None
    embedding_pos: Optional[torch.Tensor] = None
    labels: Optional[torch.Tensor] = None
    noise: Optional[torch.Tensor] = None

    def __post_init__(self):
        if self.embedding_scalar is not None and self.noise is not None and self.noise.ndim != self.embedding_scalar.ndim + 2:
            raise ValueError("noise must have 2 more dimensions than embedding_scalar")
        if self.labels is not None and self.labels.ndim != 2:
            raise ValueError("labels must have 2 dimensions")


@dataclasses.dataclass
class SFNONetConfig:
    """Fields and defaults of the reference's ``SFNONetConfig`` (sfnonet.py:46-148)."""

    embed_dim: int = 256
    filter_type: str = "linear"
    scale_factor: int = 1
    global_layer_norm: bool = False
    num_layers: int = 12
    use_mlp: bool = True
    mlp_ratio: float = 2.0
    activation_function: str = "gelu"
    encoder_layers: int = 1
    pos_embed: bool = True
    drop_rate: float = 0.0
    drop_path_rate: float = 0.0
    hard_thresholding_fraction: float = 1.0
    big_skip: bool = True
    checkpointing: int = 0
    filter_num_groups: int = 1
    filter_residual: bool = False
    filter_output: bool = False
    local_blocks: Optional[list] = None
    normalize_big_skip: bool = False
    affine_norms: bool = False
    lora_rank: int = 0
    lora_alpha: Optional[float] = None
    spectral_lora_rank: int = 0
    spectral_lora_alpha: Optional[float] = None
    filter_preserves_global_mean: bool = False
    spectral_ratio: float = 1.0
    clip_latent_global_means: bool = False


# ---------------------------------------------------------------------------------------------- parameter holders
class _ChannelLayerNorm(nn.Module):
    def __init__(self, n_channels, elementwise_affine):
        super().__init__()
        if elementwise_affine:
            self.weight = nn.Parameter(torch.ones(n_channels))
            self.bias = nn.Parameter(torch.zeros(n_channels))


class _ConditionalLayerNorm(nn.Module):
    """Holder for ConditionalLayerNorm (layers.py:143-283): same layers in the same order, conditioning initialised to identity."""

    def __init__(self, n_channels, cc: ContextConfig, elementwise_affine):
        super().__init__()
        if cc.embed_dim_scalar > 0:
            self.W_scale = nn.Linear(cc.embed_dim_scalar, n_channels)
            self.W_bias = nn.Linear(cc.embed_dim_scalar, n_channels)
        if cc.embed_dim_labels > 0:
            self.W_scale_labels = nn.Linear(cc.embed_dim_labels, n_channels)
            self.W_bias_labels = nn.Linear(cc.embed_dim_labels, n_channels)
        if cc.embed_dim_noise > 0:
            self.W_scale_2d = nn.Conv2d(cc.embed_dim_noise, n_channels, 1, bias=False)
            self.W_bias_2d = nn.Conv2d(cc.embed_dim_noise, n_channels, 1, bias=False)
        if cc.embed_dim_pos > 0:
            self.W_scale_pos = nn.Conv2d(cc.embed_dim_pos, n_channels, 1, bias=False)
            self.W_bias_pos = nn.Conv2d(cc.embed_dim_pos, n_channels, 1, bias=False)
        self.norm = _ChannelLayerNorm(n_channels, elementwise_affine)
        with torch.no_grad():
            for name, b in (("W_scale", 1.0), ("W_bias", 0.0), ("W_scale_labels", 0.0), ("W_bias_labels", 0.0)):
                if hasattr(self, name):
                    getattr(self, name).weight.zero_()
                    getattr(self, name).bias.fill_(b)
            for name in ("W_scale_2d", "W_bias_2d", "W_scale_pos", "W_bias_pos"):
                if hasattr(self, name):
                    getattr(self, name).weight.zero_()


def _lora_scaling(rank, alpha):
    return (float(alpha) if alpha is not None else float(rank)) / float(rank)


class _LoRAConv2d(nn.Conv2d):
    """Holder for LoRAConv2d (lora.py:9-141), 1x1: ``weight``, ``bias``, ``lora_down.weight``, ``lora_up.weight`` created and
    initialised in the reference's order (base conv, down, up, then Kaiming re-draw of down and zeroed up)."""

    def __init__(self, in_channels, out_channels, bias=True, lora_rank=0, lora_alpha=None):
        super().__init__(in_channels, out_channels, 1, 1, bias=bias)
        if lora_rank < 0:
            raise ValueError(f"lora_rank must be >= 0, got {lora_rank}")
        self.lora_rank = int(lora_rank)
        self.lora_scaling = 0.0
        if self.lora_rank > 0:
            self.lora_down = nn.Conv2d(in_channels, self.lora_rank, 1, bias=False)
            self.lora_up = nn.Conv2d(self.lora_rank, out_channels, 1, 1, bias=False)
            self.lora_scaling = _lora_scaling(lora_rank, lora_alpha)
            nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_up.weight)

    def sources(self):
        return [self.weight, self.lora_down.weight, self.lora_up.weight]

    def effective_weight(self):
        """``W + (alpha / r) W_up W_down`` [O, I, 1, 1] fp32: lora.py:131-141 is linear in x, so the update merges exactly."""
        with torch.no_grad():
            up, down = self.lora_up.weight.double().flatten(1), self.lora_down.weight.double().flatten(1)
            w = self.weight.double().flatten(1) + self.lora_scaling * (up @ down)
            return w.float().reshape(self.weight.shape).contiguous()


class _SpectralConvS2(nn.Module):
    """Holder for SpectralConvS2 (s2convolutions.py:138-280): ``pre_proj`` / ``post_proj`` (``spectral_ratio < 1``), ``weight``
    [G, L, C_s/G, C_s/G, 2], ``lora_A`` / ``lora_B``, ``bias`` [1, C, 1, 1], in the reference's creation order."""

    def __init__(self, channels, modes_lat, num_groups=1, lora_rank=0, lora_alpha=None, preserve_global_mean=False, spectral_ratio=1.0):
        super().__init__()
        sc = validate_spectral_ratio(spectral_ratio, channels, num_groups, channels_name="in_channels", num_groups_name="num_groups")
        if channels % num_groups != 0:
            raise ValueError(f"in_channels={channels} is not divisible by num_groups={num_groups}")
        self.modes_lat, self.channels, self.spectral_channels, self.num_groups = modes_lat, channels, sc, num_groups
        self.preserve_global_mean = bool(preserve_global_mean)
        if spectral_ratio < 1.0:
            self.pre_proj = nn.Conv2d(channels, sc, kernel_size=1, bias=False)
            self.post_proj = nn.Conv2d(sc, channels, kernel_size=1, bias=False)
        else:
            self.pre_proj = self.post_proj = None
        scale = math.sqrt(1 / sc) * torch.ones(modes_lat, 1, 1, 2)
        scale[0, :] *= math.sqrt(2.0)
        self.weight = nn.Parameter(scale * torch.randn(num_groups, modes_lat, sc // num_groups, sc // num_groups, 2))
        self.lora_rank = int(lora_rank)
        if self.lora_rank > 0:
            self.lora_A = nn.Parameter(scale * torch.randn(num_groups, modes_lat, lora_rank, sc // num_groups, 2))
            self.lora_B = nn.Parameter(torch.zeros(num_groups, modes_lat, sc // num_groups, lora_rank, 2))
            self.lora_scaling = _lora_scaling(lora_rank, lora_alpha)
        else:
            self.lora_A = self.lora_B = None
            self.lora_scaling = 0.0
        self.bias = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.register_load_state_dict_pre_hook(self._upgrade_old_weight_layouts)

    @staticmethod
    def _upgrade_old_weight_layouts(module, state_dict, prefix, *unused):
        """s2convolutions.py:282-357: [I, O, L, 2] (no group axis) and [G, I, O, L, 2] checkpoints are re-laid out on load."""
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is None:
            return
        g, lat, c = module.num_groups, module.modes_lat, module.weight.shape[2]
        if tuple(w.shape) == (c * g, c * g, lat, 2):
            w = w.view(1, *w.shape)
        if w.ndim == 5 and tuple(w.shape) == (g, c, c, lat, 2) and tuple(w.shape) != tuple(module.weight.shape):
            w = w.permute(0, 3, 2, 1, 4)
        state_dict[key] = w

    @property
    def is_plain(self):
        """True when ``weight`` already is the dense [1, L, C, C, 2] operator the device library takes."""
        return self.num_groups == 1 and self.lora_rank == 0 and not self.preserve_global_mean and self.pre_proj is None

    @property
    def is_grouped_native(self):
        """True when the grouped operator is multiplied as its diagonal blocks on the device (1 / G of the dense operator's weight
        bytes and multiplications -- the property the reference's only in-tree performance assertion checks,
        fme/core/models/conditional_sfno/test_sfnonet.py:262-284): groups of 64 or 128 channels, no channel bottleneck."""
        return self.num_groups > 1 and self.pre_proj is None and (self.spectral_channels // self.num_groups) in (64, 128)

    def grouped_weight(self):
        """[G, L, C/G, C/G, 2] fp32: the diagonal blocks with the LoRA update merged (per group) and, for
        ``preserve_global_mean``, the l = 0 blocks replaced by the identity."""
        with torch.no_grad():
            w = self.weight.detach()
            if self.lora_rank > 0:
                a = torch.view_as_complex(self.lora_A.double().contiguous())
                b = torch.view_as_complex(self.lora_B.double().contiguous())
                w = torch.view_as_real(torch.view_as_complex(w.double().contiguous()) + self.lora_scaling * torch.einsum("gxor,gxri->gxoi", b, a))
            w = w.float().clone() if w.data_ptr() == self.weight.data_ptr() else w.float()
            if self.preserve_global_mean:
                cg = w.shape[2]
                w[:, 0] = 0.0
                w[:, 0, :, :, 0] = torch.eye(cg, device=w.device)
            return w.contiguous()

    def sources(self):
        return [p for p in (self.weight, self.lora_A, self.lora_B, getattr(self.pre_proj, "weight", None),
                            getattr(self.post_proj, "weight", None)) if p is not None]

    def effective_weight(self):
        """The dense per-degree operator [1, L, C, C, 2] fp32 equivalent to this layer's grouped / LoRA / mean-preserving /
        bottlenecked contraction (module docstring); fp64 arithmetic, a few degrees at a time to bound the scratch memory."""
        G, L, cg = self.num_groups, self.modes_lat, self.spectral_channels // self.num_groups
        C, sc = self.channels, self.spectral_channels
        with torch.no_grad():
            dev = self.weight.device
            out = torch.empty(L, C, C, 2, dtype=torch.float32, device=dev)
            P = self.pre_proj.weight.double().flatten(1).to(torch.complex128) if self.pre_proj is not None else None    # [sc, C]
            Q = self.post_proj.weight.double().flatten(1).to(torch.complex128) if self.post_proj is not None else None  # [C, sc]
            step = max(1, min(L, (1 << 22) // max(1, C * C)))
            for l0 in range(0, L, step):
                l1 = min(L, l0 + step)
                w = torch.view_as_complex(self.weight[:, l0:l1].double().contiguous())  # [G, l, o, i]
                if self.lora_rank > 0:
                    a = torch.view_as_complex(self.lora_A[:, l0:l1].double().contiguous())  # [G, l, r, i]
                    b = torch.view_as_complex(self.lora_B[:, l0:l1].double().contiguous())  # [G, l, o, r]
                    w = w + self.lora_scaling * torch.einsum("gxor,gxri->gxoi", b, a)
                dense = torch.zeros(l1 - l0, sc, sc, dtype=torch.complex128, device=dev)
                for g in range(G):
                    dense[:, g * cg:(g + 1) * cg, g * cg:(g + 1) * cg] = w[g]
                if self.preserve_global_mean and l0 == 0:
                    dense[0] = torch.eye(sc, dtype=torch.complex128, device=dev)
                if P is not None:
                    dense = Q @ dense @ P
                out[l0:l1] = torch.view_as_real(dense).float()
            return out.unsqueeze(0).contiguous()


def validate_spectral_ratio(spectral_ratio, channels, num_groups, *, channels_name="embed_dim", num_groups_name="filter_num_groups",
                            filter_type=None, local_blocks=False):
    """s2convolutions.py:34-91: the spectral channel count ``round(channels * spectral_ratio)`` and the reference's errors."""
    if not 0.0 < spectral_ratio <= 1.0:
        raise ValueError(f"spectral_ratio must be in (0, 1], got {spectral_ratio}.")
    spectral_channels = round(channels * spectral_ratio)
    if spectral_ratio < 1.0:
        if filter_type is not None and filter_type != "linear":
            raise NotImplementedError(f"spectral_ratio < 1 is only supported for filter_type='linear', got filter_type='{filter_type}'.")
        if local_blocks:
            raise NotImplementedError("spectral_ratio < 1 is not supported with local_blocks, since local (DISCO) blocks have no "
                                      "spectral filter to bottleneck.")
        if spectral_channels < 1:
            raise ValueError(f"spectral_ratio={spectral_ratio} with {channels_name}={channels} produces fewer than 1 spectral channel.")
        if spectral_channels % num_groups != 0:
            raise ValueError(f"spectral_ratio={spectral_ratio} with {channels_name}={channels} yields {spectral_channels} spectral "
                             f"channels, which is not divisible by {num_groups_name}={num_groups}.")
    return spectral_channels


class _FilterLayer(nn.Module):
    def __init__(self, channels, modes_lat, **kw):
        super().__init__()
        self.filter = _SpectralConvS2(channels, modes_lat, **kw)


class _MLP(nn.Module):
    def __init__(self, channels, hidden, lora_rank=0, lora_alpha=None):
        super().__init__()
        self.fwd = nn.Sequential(_LoRAConv2d(channels, hidden, True, lora_rank, lora_alpha), nn.GELU(),
                                 _LoRAConv2d(hidden, channels, True, lora_rank, lora_alpha))


class _Block(nn.Module):
    """Holder for FourierNeuralOperatorBlock (sfnonet.py:262-374), same registration order."""

    def __init__(self, channels, hidden, modes_lat, cc, p):
        super().__init__()
        self.norm0 = _ConditionalLayerNorm(channels, cc, p.affine_norms)
        self.filter = _FilterLayer(channels, modes_lat, num_groups=p.filter_num_groups, lora_rank=p.spectral_lora_rank,
                                   lora_alpha=p.spectral_lora_alpha, preserve_global_mean=p.filter_preserves_global_mean,
                                   spectral_ratio=p.spectral_ratio)
        self.inner_skip = _LoRAConv2d(channels, channels, True, p.lora_rank, p.lora_alpha)
        self.norm1 = _ConditionalLayerNorm(channels, cc, p.affine_norms)
        self.mlp = _MLP(channels, hidden, p.lora_rank, p.lora_alpha)


# ---------------------------------------------------------------------------------------------- the conditional network
class SphericalFourierNeuralOperatorNet(nn.Module):
    """``forward(x [B, C_in, H, W], context: Context) -> [B, C_out, H, W]`` (sfnonet.py:773-824)."""

    def __init__(self, params: SFNONetConfig, img_shape: Tuple[int, int], in_chans: int, out_chans: int,
                 context_config: ContextConfig = ContextConfig(), data_grid: str = "equiangular"):
        super().__init__()
        p = params
        unsupported = []
        if p.filter_type != "linear":
            unsupported.append(f"filter_type={p.filter_type!r}")
        if p.scale_factor != 1:
            raise NotImplementedError("scale factor must be 1 as it is not implemented for conditional layer normalization")
        validate_spectral_ratio(p.spectral_ratio, p.embed_dim, p.filter_num_groups, filter_type=p.filter_type,
                                local_blocks=bool(p.local_blocks))  # SFNONetConfig.__post_init__ (sfnonet.py:139-146)
        for name, off in (("global_layer_norm", False), ("encoder_layers", 1), ("drop_rate", 0.0),
                          ("drop_path_rate", 0.0), ("use_mlp", True)):
            if getattr(p, name) != off:
                unsupported.append(f"{name}={getattr(p, name)!r}")
        if p.local_blocks:
            unsupported.append("local_blocks")
        if p.spectral_ratio < 1.0 and (p.filter_residual or data_grid != "legendre-gauss"):
            unsupported.append("spectral_ratio < 1 with a round-trip residual (filter_residual or a non-Gaussian data grid)")
        if p.activation_function != "gelu":
            if p.activation_function not in ("relu", "silu"):
                raise ValueError(f"Unknown activation function {p.activation_function}")
            unsupported.append(f"activation_function={p.activation_function!r}")
        if unsupported:
            raise NotImplementedError("ace_b200 NoiseConditionedSFNO does not implement: " + ", ".join(unsupported))
        self.params, self.context_config, self.data_grid = p, context_config, data_grid
        self.img_shape = tuple(img_shape)
        self.in_chans, self.out_chans, self.embed_dim, self.num_layers = in_chans, out_chans, p.embed_dim, p.num_layers
        self.big_skip, self.affine_norms = p.big_skip, p.affine_norms
        self.filter_residual, self.filter_output = bool(p.filter_residual), bool(p.filter_output)
        h, w = self.img_shape
        self.modes_lat = int(h * p.hard_thresholding_fraction)
        self.modes_lon = int((w // 2 + 1) * p.hard_thresholding_fraction)
        self.mlp_hidden = int(p.embed_dim * p.mlp_ratio)
        kw = dict(lmax=self.modes_lat, mmax=self.modes_lon)
        self.trans_down = RealSHT(h, w, grid=data_grid, **kw)
        self.itrans_up = InverseRealSHT(h, w, grid=data_grid, **kw)
        self.trans = RealSHT(h, w, grid="legendre-gauss", **kw)
        self.itrans = InverseRealSHT(h, w, grid="legendre-gauss", **kw)
        C = p.embed_dim
        lora = (p.lora_rank, p.lora_alpha)
        self.encoder = nn.Sequential(_LoRAConv2d(in_chans, C, True, *lora), nn.GELU(), _LoRAConv2d(C, C, False, *lora))
        self.blocks = nn.ModuleList([_Block(C, self.mlp_hidden, self.modes_lat, context_config, p) for _ in range(p.num_layers)])
        self.decoder = nn.Sequential(_LoRAConv2d(C + p.big_skip * in_chans, C, True, *lora), nn.GELU(), _LoRAConv2d(C, out_chans, False, *lora))
        if p.pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, C, h, w))
            self.pos_embed.is_shared_mp = ["matmul"]
            trunc_normal_(self.pos_embed, std=0.02)
        else:
            self.pos_embed = None
        if p.normalize_big_skip:
            self.norm_big_skip = _ConditionalLayerNorm(in_chans, context_config, p.affine_norms)
        self.clip_latent_global_means = bool(p.clip_latent_global_means)
        if self.clip_latent_global_means:  # sfnonet.py:730-746: the envelope travels in the checkpoint as two buffers
            self.register_buffer("_gm_min", torch.full((1, C, 1, 1), float("inf")))
            self.register_buffer("_gm_max", torch.full((1, C, 1, 1), float("-inf")))
        self._net = None
        self._net_device = None
        self._plans = None
        self._uploaded = {}

    # ------------------------------------------------------------------ device object management
    def _config(self):
        cc = self.context_config
        return _lib.CsfnoConfig(
            img_h=self.img_shape[0], img_w=self.img_shape[1], in_chans=self.in_chans, out_chans=self.out_chans, embed_dim=self.embed_dim,
            num_layers=self.num_layers, lmax=self.modes_lat, mmax=self.modes_lon, mlp_hidden=self.mlp_hidden,
            pos_embed=int(self.pos_embed is not None), big_skip=int(bool(self.big_skip)), normalize_big_skip=int(hasattr(self, "norm_big_skip")),
            affine_norms=int(bool(self.affine_norms)), embed_dim_scalar=cc.embed_dim_scalar, embed_dim_labels=cc.embed_dim_labels,
            embed_dim_noise=cc.embed_dim_noise, embed_dim_pos=cc.embed_dim_pos, norm_eps=1e-5,
            filter_residual=int(self.filter_residual), filter_output=int(self.filter_output),
            clip_latent_global_means=int(self.clip_latent_global_means))

    def _release(self):
        if getattr(self, "_net", None) is not None:
            try:
                _lib.load().ace_csfno_destroy(self._net)
            except Exception:  # noqa: BLE001
                pass
        self._net, self._net_device, self._plans, self._uploaded = None, None, None, {}

    def __del__(self):
        try:
            self._release()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def _ensure_net(self, device):
        if self._net is not None and self._net_device == device:
            return
        self._release()
        outer = ShtPlan.get(*self.img_shape, self.modes_lat, self.modes_lon, self.data_grid)
        inner = ShtPlan.get(*self.img_shape, self.modes_lat, self.modes_lon, "legendre-gauss")
        cfg = self._config()
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().ace_csfno_create(ctypes.byref(cfg), outer.handle, inner.handle, ctypes.byref(handle)))
        self._net, self._net_device, self._plans = handle, device, (outer, inner)

    def device_parameters(self):
        """``[(key, sources, build)]``: every parameter the device library takes, under the reference's ``state_dict`` key of the
        un-adapted layer, with the torch parameters it is derived from.  ``build`` is ``None`` when the single source is passed
        as stored; otherwise it returns the folded fp32 tensor (module docstring: LoRA merges, grouped / mean-preserving /
        bottlenecked spectral operators)."""
        derived, hidden = {}, set()
        for prefix, mod in self.named_modules():
            if isinstance(mod, _LoRAConv2d) and mod.lora_rank > 0:
                derived[prefix + ".weight"] = (mod.sources(), mod.effective_weight)
                hidden.update((prefix + ".lora_down.weight", prefix + ".lora_up.weight"))
            elif isinstance(mod, _SpectralConvS2) and not mod.is_plain:
                derived[prefix + ".weight"] = (mod.sources(), mod.grouped_weight if mod.is_grouped_native else mod.effective_weight)
                hidden.update(prefix + "." + n for n in ("lora_A", "lora_B", "pre_proj.weight", "post_proj.weight"))
        out = []
        for name, prm in self.named_parameters():
            if name in hidden or (not self.big_skip and name.startswith("norm_big_skip.")):
                continue  # (the reference still creates norm_big_skip without a big skip; it is never used)
            sources, build = derived.get(name, ([prm], None))
            out.append((name, sources, build))
        out += [(name, [buf], None) for name, buf in self.named_buffers() if name in ("_gm_min", "_gm_max")]
        return out

    def request_latent_global_mean_envelope_reset(self) -> None:
        """sfnonet.py:754-762: asks the next TRAINING forward to restart the envelope; inference never updates it, so nothing to do."""

    def _sync_params(self, stream):
        lib = _lib.load()
        dirty = False
        for name, sources, build in self.device_parameters():
            key = tuple((p.data_ptr(), p._version) for p in sources)
            if self._uploaded.get(name) == key:
                continue
            for p in sources:
                if p.device != self._net_device and not (getattr(self, "_offloaded", False) and p.device.type == "cpu"):
                    raise _lib.AceError(f"parameter {name} is on {p.device}, input is on {self._net_device}")
            t = sources[0].detach() if build is None else build()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            if t.device != self._net_device:  # offload_parameters(): an edited host copy, staged through a temporary device tensor
                t = t.to(self._net_device)
            _lib.check(lib.ace_csfno_set_param(self._net, name.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), stream))
            self._uploaded[name] = key
            dirty = True
        if dirty:
            _lib.check(lib.ace_csfno_finalize(self._net, stream))

    def offload_parameters(self):
        """Single weight residency on the GPU (same contract as the deterministic network's ``offload_parameters``): once the device
        library holds its split-plane copies, the torch fp32 parameters of this network move to host memory (-3 GB for the ERA5
        baseline).  ``state_dict`` / ``load_state_dict`` / in-place edits keep working on the host copies; an edited parameter (or a
        fold that depends on it) is rebuilt and re-uploaded before the next forward.  The envelope buffers stay on the device."""
        if self._net is None:
            raise _lib.AceError("offload_parameters(): run one forward first (the device library has no parameters yet)")
        with torch.cuda.device(self._net_device):
            self._sync_params(_lib.current_stream_ptr())
            torch.cuda.current_stream().synchronize()
        for prm in self.parameters():
            prm.data = prm.data.cpu()
        for name, sources, _ in self.device_parameters():
            self._uploaded[name] = tuple((p.data_ptr(), p._version) for p in sources)
        self._offloaded = True
        return self

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, context: Context):
        if not x.is_cuda:
            raise _lib.AceError("ace_b200 NoiseConditionedSFNO: input must be a CUDA tensor (there is no CPU path)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise _lib.AceError("ace_b200 NoiseConditionedSFNO is inference-only: call under torch.no_grad() (or requires_grad_(False))")
        if x.dim() != 4 or x.shape[1] != self.in_chans or tuple(x.shape[-2:]) != self.img_shape:
            raise ValueError(f"expected input [B, {self.in_chans}, {self.img_shape[0]}, {self.img_shape[1]}], got {tuple(x.shape)}")
        cc = self.context_config
        B = x.shape[0]

        def ctx(t, width, what, spatial):
            if width == 0:
                return None
            if t is None:
                raise ValueError(f"{what} must be provided")
            shape = (B, width, *self.img_shape) if spatial else (B, width)
            if tuple(t.shape) != shape:
                raise ValueError(f"{what}: expected {shape}, got {tuple(t.shape)}")
            return t.to(device=x.device, dtype=torch.float32).contiguous()

        scalar = ctx(context.embedding_scalar, cc.embed_dim_scalar, "embedding_scalar", False)
        labels = ctx(context.labels, cc.embed_dim_labels, "labels", False)
        noise = ctx(context.noise, cc.embed_dim_noise, "noise", True)
        pos = ctx(context.embedding_pos, cc.embed_dim_pos, "embedding_pos", True)
        dtype = x.dtype
        x = x.float().contiguous()
        y = torch.empty(B, self.out_chans, *self.img_shape, dtype=torch.float32, device=x.device)
        if B == 0:
            return y.to(dtype)

        def ptr(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None

        with torch.cuda.device(x.device):
            self._ensure_net(x.device)
            stream = _lib.current_stream_ptr()
            self._sync_params(stream)
            _lib.check(_lib.load().ace_csfno_forward(self._net, ptr(x), ptr(scalar), ptr(labels), ptr(noise), ptr(pos), ptr(y), B, stream))
        return y.to(dtype)


def _native_handle(self):
    """The ace_csfno* behind this module (after at least one forward); used by the fused stepper."""
    return self._net


SphericalFourierNeuralOperatorNet.native_handle = _native_handle


def get_lat_lon_sfnonet(params: SFNONetConfig, in_chans: int, out_chans: int, img_shape: Tuple[int, int], data_grid: str = "equiangular",
                        context_config: ContextConfig = ContextConfig()) -> SphericalFourierNeuralOperatorNet:
    """sfnonet.py:443-493."""
    return SphericalFourierNeuralOperatorNet(params, img_shape, in_chans, out_chans, context_config, data_grid)


# ---------------------------------------------------------------------------------------------- noise + wrapper
def isotropic_noise(leading_shape, lmax: int, mmax: int, isht: InverseRealSHT, device, normals=None) -> torch.Tensor:
    """stochastic_sfno.py:21-47 on the device: ``normals = (real, imag)`` overrides the two N(0,1) draws (tests)."""
    shape = (*leading_shape, lmax, mmax)
    if normals is None:
        real = torch.randn(shape, dtype=torch.float32, device=device)
        imag = torch.randn(shape, dtype=torch.float32, device=device)
    else:
        real, imag = (t.to(device=device, dtype=torch.float32).contiguous() for t in normals)
    if isht.lmax != lmax or isht.mmax != mmax:
        raise ValueError("isotropic_noise: lmax / mmax must be the inverse transform's")
    nf = 1
    for s in leading_shape:
        nf *= int(s)
    scratch = torch.empty(*shape, 2, dtype=torch.float32, device=device)
    out = torch.empty(*leading_shape, isht.nlat, isht.nlon, dtype=torch.float32, device=device)
    if nf == 0:
        return out
    with torch.cuda.device(device):
        _lib.check(_lib.load().ace_isotropic_noise(isht.plan().handle, ctypes.c_void_p(real.data_ptr()), ctypes.c_void_p(imag.data_ptr()),
                                                   ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(out.data_ptr()), nf,
                                                   _lib.current_stream_ptr()))
    return out


class NoiseConditionedModel(nn.Module):
    """stochastic_sfno.py:50-172: draws the noise, builds the Context, calls the conditional network.

    ``forward(x, labels=None, noise=None)``; ``noise`` (not in the reference's signature) overrides the draw for parity tests."""

    def __init__(self, conditional_model: nn.Module, img_shape: Tuple[int, int], embed_dim_noise: int = 256, embed_dim_pos: int = 0,
                 n_labels: int = 0, label_embed_dim: int = 0, inverse_sht=None, lmax: int = 0, mmax: int = 0):
        super().__init__()
        self.conditional_model = conditional_model
        self.embed_dim, self.img_shape = embed_dim_noise, tuple(img_shape)
        self._inverse_sht, self._lmax, self._mmax = inverse_sht, lmax, mmax
        if label_embed_dim > 0 and n_labels == 0:
            raise ValueError("label_embed_dim > 0 requires n_labels > 0")
        if label_embed_dim > 0:
            self.label_embedding = nn.Linear(n_labels, label_embed_dim)
            effective_label_dim = label_embed_dim
        else:
            self.label_embedding = None
            effective_label_dim = n_labels
        self.label_pos_embed = None
        if embed_dim_pos != 0:
            self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim_pos, *self.img_shape))
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
            if effective_label_dim > 0:
                self.label_pos_embed = nn.Parameter(torch.zeros(effective_label_dim, embed_dim_pos, *self.img_shape))
                nn.init.trunc_normal_(self.label_pos_embed, std=0.02)
        else:
            self.pos_embed = None

    def draw_noise(self, batch: int, device, normals=None) -> torch.Tensor:
        if self._inverse_sht is not None:
            return isotropic_noise((batch, self.embed_dim), self._lmax, self._mmax, self._inverse_sht, device, normals)
        return torch.randn(batch, self.embed_dim, *self.img_shape, device=device, dtype=torch.float32)

    def forward(self, x: torch.Tensor, labels: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = x.reshape(-1, *x.shape[-3:])
        B = x.shape[0]
        if noise is None and self.embed_dim > 0:
            noise = self.draw_noise(B, x.device)
        lib, vp = _lib.load(), ctypes.c_void_p
        if labels is not None and (self.label_embedding is not None or self.label_pos_embed is not None):
            if not x.is_cuda:
                raise _lib.AceError("ace_b200 NoiseConditionedSFNO: input must be a CUDA tensor (there is no CPU path)")
            labels = labels.to(device=x.device, dtype=torch.float32).contiguous()
        if labels is not None and self.label_embedding is not None and B > 0:
            lin = self.label_embedding  # stochastic_sfno.py:152-153
            if labels.dim() != 2 or labels.shape != (B, lin.in_features):
                raise ValueError(f"labels: expected {(B, lin.in_features)}, got {tuple(labels.shape)}")
            w, b = lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous()
            emb = torch.empty(B, lin.out_features, dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                _lib.check(lib.ace_label_embed(vp(labels.data_ptr()), vp(w.data_ptr()), vp(b.data_ptr()), B, lin.in_features, lin.out_features,
                                               vp(emb.data_ptr()), _lib.current_stream_ptr()))
            labels = emb
        pos = None
        if self.pos_embed is not None:
            if self.label_pos_embed is not None and labels is not None and B > 0:  # stochastic_sfno.py:157-165
                n = self.label_pos_embed.shape[0]
                if labels.dim() != 2 or labels.shape != (B, n):
                    raise ValueError(f"labels: expected {(B, n)}, got {tuple(labels.shape)}")
                base = self.pos_embed.detach().float().contiguous()
                lpe = self.label_pos_embed.detach().float().contiguous()
                pos = torch.empty(B, *base.shape[1:], dtype=torch.float32, device=x.device)
                with torch.cuda.device(x.device):
                    _lib.check(lib.ace_label_pos_embed(vp(base.data_ptr()), vp(labels.data_ptr()), vp(lpe.data_ptr()), B, n, base.numel(),
                                                       vp(pos.data_ptr()), _lib.current_stream_ptr()))
            else:
                pos = self.pos_embed.expand(B, -1, -1, -1)
        return self.conditional_model(x, Context(embedding_scalar=None, embedding_pos=pos, labels=labels, noise=noise if self.embed_dim > 0 else None))


NoiseConditionedSFNO = NoiseConditionedModel  # the reference's alias (stochastic_sfno.py:175-176)
