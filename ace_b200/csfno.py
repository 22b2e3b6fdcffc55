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

Unsupported reference options raise ``NotImplementedError`` at construction (no fallback): ``filter_type`` other than
``"linear"``, ``filter_num_groups != 1``, ``global_layer_norm``, ``filter_residual`` / ``filter_output``, local (DISCO) blocks,
LoRA, ``spectral_ratio != 1``, ``filter_preserves_global_mean``, ``clip_latent_global_means``, ``use_mlp=False``,
``encoder_layers != 1``, dropout, activation other than GELU, noise + positional context wider than 64 channels, learned label
embeddings / label-position interaction.
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


class _SpectralConvS2(nn.Module):
    """Holder for SpectralConvS2 (s2convolutions.py:162-280): ``weight`` [1, L, C, C, 2] and ``bias`` [1, C, 1, 1]."""

    def __init__(self, channels, modes_lat):
        super().__init__()
        self.modes_lat, self.channels = modes_lat, channels
        scale = math.sqrt(1 / channels) * torch.ones(modes_lat, 1, 1, 2)
        scale[0, :] *= math.sqrt(2.0)
        self.weight = nn.Parameter(scale * torch.randn(1, modes_lat, channels, channels, 2))
        self.bias = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.register_load_state_dict_pre_hook(self._upgrade_old_weight_layouts)

    @staticmethod
    def _upgrade_old_weight_layouts(module, state_dict, prefix, *unused):
        """s2convolutions.py:282-357: [I, O, L, 2] (no group axis) and [G, I, O, L, 2] checkpoints are re-laid out on load."""
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is None:
            return
        c, lat = module.channels, module.modes_lat
        if tuple(w.shape) == (c, c, lat, 2):
            w = w.view(1, c, c, lat, 2)
        if tuple(w.shape) == (1, c, c, lat, 2) and tuple(w.shape) != tuple(module.weight.shape):
            w = w.permute(0, 3, 2, 1, 4)
        state_dict[key] = w


class _FilterLayer(nn.Module):
    def __init__(self, channels, modes_lat):
        super().__init__()
        self.filter = _SpectralConvS2(channels, modes_lat)


class _MLP(nn.Module):
    def __init__(self, channels, hidden):
        super().__init__()
        self.fwd = nn.Sequential(nn.Conv2d(channels, hidden, 1, bias=True), nn.GELU(), nn.Conv2d(hidden, channels, 1, bias=True))


class _Block(nn.Module):
    """Holder for FourierNeuralOperatorBlock (sfnonet.py:262-374), same registration order."""

    def __init__(self, channels, hidden, modes_lat, cc, affine_norms):
        super().__init__()
        self.norm0 = _ConditionalLayerNorm(channels, cc, affine_norms)
        self.filter = _FilterLayer(channels, modes_lat)
        self.inner_skip = nn.Conv2d(channels, channels, 1, 1)
        self.norm1 = _ConditionalLayerNorm(channels, cc, affine_norms)
        self.mlp = _MLP(channels, hidden)


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
        for name, off in (("global_layer_norm", False), ("filter_residual", False), ("filter_output", False), ("lora_rank", 0),
                          ("spectral_lora_rank", 0), ("filter_preserves_global_mean", False), ("clip_latent_global_means", False),
                          ("filter_num_groups", 1), ("spectral_ratio", 1.0), ("encoder_layers", 1), ("drop_rate", 0.0),
                          ("drop_path_rate", 0.0), ("use_mlp", True)):
            if getattr(p, name) != off:
                unsupported.append(f"{name}={getattr(p, name)!r}")
        if p.local_blocks:
            unsupported.append("local_blocks")
        if p.activation_function != "gelu":
            if p.activation_function not in ("relu", "silu"):
                raise ValueError(f"Unknown activation function {p.activation_function}")
            unsupported.append(f"activation_function={p.activation_function!r}")
        if context_config.embed_dim_noise + context_config.embed_dim_pos > 64:
            unsupported.append("noise + positional context wider than 64 channels")
        if unsupported:
            raise NotImplementedError("ace_b200 NoiseConditionedSFNO does not implement: " + ", ".join(unsupported))
        self.params, self.context_config, self.data_grid = p, context_config, data_grid
        self.img_shape = tuple(img_shape)
        self.in_chans, self.out_chans, self.embed_dim, self.num_layers = in_chans, out_chans, p.embed_dim, p.num_layers
        self.big_skip, self.affine_norms = p.big_skip, p.affine_norms
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
        self.encoder = nn.Sequential(nn.Conv2d(in_chans, C, 1, bias=True), nn.GELU(), nn.Conv2d(C, C, 1, bias=False))
        self.blocks = nn.ModuleList([_Block(C, self.mlp_hidden, self.modes_lat, context_config, p.affine_norms) for _ in range(p.num_layers)])
        self.decoder = nn.Sequential(nn.Conv2d(C + p.big_skip * in_chans, C, 1, bias=True), nn.GELU(), nn.Conv2d(C, out_chans, 1, bias=False))
        if p.pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, C, h, w))
            self.pos_embed.is_shared_mp = ["matmul"]
            trunc_normal_(self.pos_embed, std=0.02)
        else:
            self.pos_embed = None
        if p.normalize_big_skip:
            self.norm_big_skip = _ConditionalLayerNorm(in_chans, context_config, p.affine_norms)
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
            embed_dim_noise=cc.embed_dim_noise, embed_dim_pos=cc.embed_dim_pos, norm_eps=1e-5)

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

    def _sync_params(self, stream):
        lib = _lib.load()
        dirty = False
        skip_big = not self.big_skip  # the reference still creates norm_big_skip without a big skip; it is never used
        for name, prm in self.named_parameters():
            if skip_big and name.startswith("norm_big_skip."):
                continue
            key = (prm.data_ptr(), prm._version)
            if self._uploaded.get(name) == key:
                continue
            if prm.device != self._net_device:
                raise _lib.AceError(f"parameter {name} is on {prm.device}, input is on {self._net_device}")
            t = prm.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            _lib.check(lib.ace_csfno_set_param(self._net, name.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), stream))
            self._uploaded[name] = key
            dirty = True
        if dirty:
            _lib.check(lib.ace_csfno_finalize(self._net, stream))

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
        if label_embed_dim > 0:
            raise NotImplementedError("ace_b200 NoiseConditionedSFNO does not implement learned label embeddings (label_embed_dim > 0)")
        if embed_dim_pos != 0 and n_labels > 0:
            raise NotImplementedError("ace_b200 NoiseConditionedSFNO does not implement the label-position interaction embedding")
        self.conditional_model = conditional_model
        self.embed_dim, self.img_shape = embed_dim_noise, tuple(img_shape)
        self._inverse_sht, self._lmax, self._mmax = inverse_sht, lmax, mmax
        self.label_embedding = None
        self.label_pos_embed = None
        if embed_dim_pos != 0:
            self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim_pos, *self.img_shape))
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
        else:
            self.pos_embed = None

    def draw_noise(self, batch: int, device, normals=None) -> torch.Tensor:
        if self._inverse_sht is not None:
            return isotropic_noise((batch, self.embed_dim), self._lmax, self._mmax, self._inverse_sht, device, normals)
        return torch.randn(batch, self.embed_dim, *self.img_shape, device=device, dtype=torch.float32)

    def forward(self, x: torch.Tensor, labels: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = x.reshape(-1, *x.shape[-3:])
        if noise is None and self.embed_dim > 0:
            noise = self.draw_noise(x.shape[0], x.device)
        pos = self.pos_embed.expand(x.shape[0], -1, -1, -1) if self.pos_embed is not None else None
        return self.conditional_model(x, Context(embedding_scalar=None, embedding_pos=pos, labels=labels, noise=noise if self.embed_dim > 0 else None))


NoiseConditionedSFNO = NoiseConditionedModel  # the reference's alias (stochastic_sfno.py:175-176)
