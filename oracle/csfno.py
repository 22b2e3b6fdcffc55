"""Noise-conditioned SFNO (``NoiseConditionedSFNO``) -- oracle restatement on torch-CPU.  TEST INFRASTRUCTURE.

Restates, for filter_type="linear" and scale_factor=1 (no local (DISCO) blocks, no dropout) -- i.e. the configuration family
of the ACE2-ERA5 / stochastic baselines plus the grouped, LoRA, global-mean-preserving and spectral_ratio variants:

  /root/reference/fme/core/models/conditional_sfno/layers.py:33-95     ContextConfig / Context
  /root/reference/fme/core/models/conditional_sfno/layers.py:95-141    ChannelLayerNorm
  /root/reference/fme/core/models/conditional_sfno/layers.py:143-320   ConditionalLayerNorm
  /root/reference/fme/core/models/conditional_sfno/layers.py:363-415   MLP
  /root/reference/fme/core/models/conditional_sfno/lora.py:9-141       LoRAConv2d
  /root/reference/fme/core/models/conditional_sfno/s2convolutions.py:118-433  _contract_dhconv, SpectralConvS2
  /root/reference/fme/core/models/conditional_sfno/sfnonet.py:262-436  FourierNeuralOperatorBlock
  /root/reference/fme/core/models/conditional_sfno/sfnonet.py:443-824  get_lat_lon_sfnonet, SphericalFourierNeuralOperatorNet
  /root/reference/fme/ace/registry/stochastic_sfno.py:21-172           isotropic_noise, NoiseConditionedModel

Submodules are created in the reference's order with the same torch layers, so ``state_dict()`` keys / shapes are the
reference's (its stored checkpoint golden loads with ``load_state_dict``) and construction under ``torch.manual_seed`` draws
the same numbers.  Pinned in tests/test_oracle_csfno.py against the reference's stored goldens and against the live
reference modules (oracle/refload.py:load_csfno) when /root/reference is present.
"""
import dataclasses
import math
from typing import Optional

import torch
import torch.nn as nn

from .sfno import trunc_normal_
from .sht import InverseRealSHT, RealSHT


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


class ChannelLayerNorm(nn.Module):
    """Per-pixel normalisation over the channel axis, biased variance (layers.py:95-141)."""

    def __init__(self, n_channels, eps=1e-5, elementwise_affine=False):
        super().__init__()
        self.eps = eps
        if elementwise_affine:
            self.weight = nn.Parameter(torch.ones(n_channels))
            self.bias = nn.Parameter(torch.zeros(n_channels))
        else:
            self.weight = self.bias = None

    def forward(self, x):
        mean = x.mean(dim=-3, keepdim=True)
        var = x.var(dim=-3, keepdim=True, unbiased=False)
        y = (x - mean) * torch.rsqrt(var + self.eps)
        if self.weight is not None:
            y = y * self.weight.view(1, -1, 1, 1) + self.bias.view(1, -1, 1, 1)
        return y


class ConditionalLayerNorm(nn.Module):
    """layers.py:143-320: LayerNorm whose scale / bias are affine functions of the context."""

    def __init__(self, n_channels, img_shape, context_config, global_layer_norm=False, epsilon=1e-5, elementwise_affine=False):
        super().__init__()
        c = context_config
        self.W_scale = nn.Linear(c.embed_dim_scalar, n_channels) if c.embed_dim_scalar > 0 else None
        self.W_bias = nn.Linear(c.embed_dim_scalar, n_channels) if c.embed_dim_scalar > 0 else None
        self.W_scale_labels = nn.Linear(c.embed_dim_labels, n_channels) if c.embed_dim_labels > 0 else None
        self.W_bias_labels = nn.Linear(c.embed_dim_labels, n_channels) if c.embed_dim_labels > 0 else None
        self.W_scale_2d = nn.Conv2d(c.embed_dim_noise, n_channels, 1, bias=False) if c.embed_dim_noise > 0 else None
        self.W_bias_2d = nn.Conv2d(c.embed_dim_noise, n_channels, 1, bias=False) if c.embed_dim_noise > 0 else None
        self.W_scale_pos = nn.Conv2d(c.embed_dim_pos, n_channels, 1, bias=False) if c.embed_dim_pos > 0 else None
        self.W_bias_pos = nn.Conv2d(c.embed_dim_pos, n_channels, 1, bias=False) if c.embed_dim_pos > 0 else None
        if global_layer_norm:
            self.norm = nn.LayerNorm((n_channels, img_shape[0], img_shape[1]), eps=epsilon, elementwise_affine=elementwise_affine)
        else:
            self.norm = ChannelLayerNorm(n_channels, eps=epsilon, elementwise_affine=elementwise_affine)
        # reset_parameters (layers.py:254-283): the conditioning starts as the identity
        with torch.no_grad():
            for lin, b in ((self.W_scale, 1.0), (self.W_bias, 0.0), (self.W_scale_labels, 0.0), (self.W_bias_labels, 0.0)):
                if lin is not None:
                    lin.weight.zero_()
                    lin.bias.fill_(b)
            for conv in (self.W_scale_2d, self.W_bias_2d, self.W_scale_pos, self.W_bias_pos):
                if conv is not None:
                    conv.weight.zero_()

    def forward(self, x, context):
        vec = lambda t: t.unsqueeze(-1).unsqueeze(-1)  # noqa: E731
        shape = list(x.shape[:-2]) + [1, 1]
        scale = vec(self.W_scale(context.embedding_scalar)) if self.W_scale is not None else torch.ones(shape, dtype=x.dtype, device=x.device)
        if self.W_scale_2d is not None:
            scale = scale + self.W_scale_2d(context.noise)
        bias = vec(self.W_bias(context.embedding_scalar)) if self.W_bias is not None else torch.zeros(shape, dtype=x.dtype, device=x.device)
        if self.W_scale_labels is not None:
            scale = scale + vec(self.W_scale_labels(context.labels))
        if self.W_bias_labels is not None:
            bias = bias + vec(self.W_bias_labels(context.labels))
        if self.W_bias_2d is not None:
            bias = bias + self.W_bias_2d(context.noise)
        if self.W_scale_pos is not None:
            scale = scale + self.W_scale_pos(context.embedding_pos)
        if self.W_bias_pos is not None:
            bias = bias + self.W_bias_pos(context.embedding_pos)
        return self.norm(x) * scale + bias


class LoRAConv2d(nn.Conv2d):
    """lora.py:9-141 for 1x1, ungrouped, dropout-free convolutions: y = conv(x) + (alpha / r) * up(down(x)).

    Construction order (it fixes the seeded draws): the base Conv2d, ``lora_down``, ``lora_up`` (each with torch's default
    initialisation), then ``reset_lora_parameters`` re-draws ``lora_down`` (Kaiming) and zeroes ``lora_up``."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, bias=True, lora_rank=0, lora_alpha=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, bias=bias)
        self.lora_rank = int(lora_rank)
        self.lora_scaling = 0.0
        if self.lora_rank > 0:
            self.lora_down = nn.Conv2d(in_channels, self.lora_rank, 1, bias=False)
            self.lora_up = nn.Conv2d(self.lora_rank, out_channels, kernel_size, stride, bias=False)
            self.lora_scaling = (float(lora_alpha) if lora_alpha is not None else float(lora_rank)) / float(self.lora_rank)
            nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_up.weight)

    def forward(self, x):
        y = super().forward(x)
        if self.lora_rank == 0:
            return y
        return y + self.lora_up(self.lora_down(x)) * self.lora_scaling


class MLP(nn.Module):
    """layers.py:363-415 without dropout / checkpointing: fc1 (bias), activation, fc2 (bias)."""

    def __init__(self, in_features, hidden_features, act_layer=nn.GELU, lora_rank=0, lora_alpha=None):
        super().__init__()
        self.fwd = nn.Sequential(LoRAConv2d(in_features, hidden_features, 1, bias=True, lora_rank=lora_rank, lora_alpha=lora_alpha),
                                 act_layer(),
                                 LoRAConv2d(hidden_features, in_features, 1, bias=True, lora_rank=lora_rank, lora_alpha=lora_alpha))

    def forward(self, x):
        return self.fwd(x)


class SpectralConvS2(nn.Module):
    """s2convolutions.py:138-433, dense weights [G, L, O/G, I/G, 2]; out[b,g,o,l,m] = sum_i x[b,g,i,l,m] w[g,l,o,i]
    (+ the LoRA update B A x, the l = 0 pass-through of ``preserve_global_mean``, the pre / post projections of
    ``spectral_ratio < 1``)."""

    def __init__(self, forward_transform, inverse_transform, channels, num_groups=1, bias=True, filter_residual=False,
                 preserve_global_mean=False, lora_rank=0, lora_alpha=None, spectral_ratio=1.0):
        super().__init__()
        if not 0.0 < spectral_ratio <= 1.0:
            raise ValueError(f"spectral_ratio must be in (0, 1], got {spectral_ratio}.")
        sc = round(channels * spectral_ratio)  # validate_spectral_ratio (s2convolutions.py:34-91)
        if spectral_ratio < 1.0 and (sc < 1 or sc % num_groups != 0):
            raise ValueError(f"spectral_ratio={spectral_ratio} with in_channels={channels} yields {sc} spectral channels")
        assert channels % num_groups == 0
        self.num_groups, self.spectral_channels = num_groups, sc
        self.forward_transform, self.inverse_transform = forward_transform, inverse_transform
        self.modes_lat, self.modes_lon = inverse_transform.lmax, inverse_transform.mmax
        self._round_trip_residual = filter_residual or (
            forward_transform.nlat != inverse_transform.nlat or forward_transform.nlon != inverse_transform.nlon
            or forward_transform.grid != inverse_transform.grid)
        self._preserve_global_mean = preserve_global_mean
        if spectral_ratio < 1.0:
            self.pre_proj = nn.Conv2d(channels, sc, kernel_size=1, bias=False)
            self.post_proj = nn.Conv2d(sc, channels, kernel_size=1, bias=False)
        else:
            self.pre_proj = self.post_proj = None
        scale = math.sqrt(1 / sc) * torch.ones(self.modes_lat, 1, 1, 2)
        scale[0, :] *= math.sqrt(2.0)
        self.weight = nn.Parameter(scale * torch.randn(num_groups, self.modes_lat, sc // num_groups, sc // num_groups, 2))
        self.lora_scaling = 0.0
        if lora_rank > 0:
            self.lora_A = nn.Parameter(scale * torch.randn(num_groups, self.modes_lat, lora_rank, sc // num_groups, 2))
            self.lora_B = nn.Parameter(torch.zeros(num_groups, self.modes_lat, sc // num_groups, lora_rank, 2))
            self.lora_scaling = (lora_alpha if lora_alpha is not None else lora_rank) / lora_rank
        else:
            self.lora_A = self.lora_B = None
        if bias:
            self.bias = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.register_load_state_dict_pre_hook(self._upgrade_old_weight_layouts)

    @staticmethod
    def _upgrade_old_weight_layouts(module, state_dict, prefix, *unused):
        """s2convolutions.py:282-357: checkpoints written as [I, O, L, 2] (no group axis) or [G, I, O, L, 2] are re-laid out."""
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

    def forward(self, x):
        residual = x
        x = x.float()
        if self.pre_proj is not None:
            x = self.pre_proj(x)
        x = self.forward_transform(x)
        if self._round_trip_residual:
            residual = self.inverse_transform(x.contiguous())
            if self.post_proj is not None:
                residual = self.post_proj(residual)
        B, C, H, W = x.shape
        x = x.reshape(B, self.num_groups, C // self.num_groups, H, W)
        xs = x[..., : self.modes_lat, : self.modes_lon]
        xp = torch.zeros_like(x)
        xp[..., : self.modes_lat, : self.modes_lon] = torch.einsum("bgixy,gxoi->bgoxy", xs, torch.view_as_complex(self.weight))
        if self.lora_A is not None:
            tmp = torch.einsum("gxri,bgixy->bgxry", torch.view_as_complex(self.lora_A), xs)
            xp = xp + self.lora_scaling * torch.einsum("gxor,bgxry->bgoxy", torch.view_as_complex(self.lora_B), tmp)
        if self._preserve_global_mean:
            xp = torch.cat([x[..., :1, :], xp[..., 1:, :]], dim=-2)
        x = self.inverse_transform(xp.reshape(B, C, H, W).contiguous())
        if self.post_proj is not None:
            x = self.post_proj(x)
        if hasattr(self, "bias"):
            x = x + self.bias
        return x, residual


class _Filter(nn.Module):
    """SpectralFilterLayer (sfnonet.py:178-260): only the attribute nesting (``filter.filter.weight``) matters here."""

    def __init__(self, conv):
        super().__init__()
        self.filter = conv

    def forward(self, x):
        return self.filter(x)


class FourierNeuralOperatorBlock(nn.Module):
    """sfnonet.py:262-436 with inner_skip="linear", outer_skip="identity", concat_skip=False:
        xn = norm0(x, ctx); y, r = filter(xn); y = act(y + inner_skip(r)); y = mlp(norm1(y, ctx)); return y + r."""

    def __init__(self, forward_transform, inverse_transform, embed_dim, img_shape, context_config, global_layer_norm=False,
                 mlp_ratio=2.0, act_layer=nn.GELU, use_mlp=True, filter_residual=False, affine_norms=False, filter_num_groups=1,
                 filter_preserves_global_mean=False, lora_rank=0, lora_alpha=None, spectral_lora_rank=0, spectral_lora_alpha=None,
                 spectral_ratio=1.0, outer_skip="identity"):
        super().__init__()
        self.norm0 = ConditionalLayerNorm(embed_dim, img_shape, context_config, global_layer_norm, elementwise_affine=affine_norms)
        self.filter = _Filter(SpectralConvS2(forward_transform, inverse_transform, embed_dim, num_groups=filter_num_groups, bias=True,
                                             filter_residual=filter_residual, preserve_global_mean=filter_preserves_global_mean,
                                             lora_rank=spectral_lora_rank, lora_alpha=spectral_lora_alpha, spectral_ratio=spectral_ratio))
        self.inner_skip = LoRAConv2d(embed_dim, embed_dim, 1, 1, lora_rank=lora_rank, lora_alpha=lora_alpha)
        self.act_layer = act_layer()
        self.norm1 = ConditionalLayerNorm(embed_dim, img_shape, context_config, global_layer_norm, elementwise_affine=affine_norms)
        if use_mlp:
            self.mlp = MLP(embed_dim, int(embed_dim * mlp_ratio), act_layer, lora_rank, lora_alpha)
        if outer_skip == "identity":  # the network's choice (sfnonet.py:648); the block's own default is None (:283): no outer skip
            self.outer_skip = nn.Identity()
        elif outer_skip is not None:
            raise NotImplementedError(f"outer_skip={outer_skip!r}")

    def forward(self, x, context):
        x, residual = self.filter(self.norm0(x, context))
        x = self.act_layer(x + self.inner_skip(residual))
        x = self.norm1(x, context)
        if hasattr(self, "mlp"):
            x = self.mlp(x)
        if hasattr(self, "outer_skip"):
            x = x + self.outer_skip(residual)
        return x


class SphericalFourierNeuralOperatorNet(nn.Module):
    """sfnonet.py:496-824, eval path (``clip_latent_global_means``: the envelope buffers are only applied, never updated)."""

    def __init__(self, img_shape, in_chans, out_chans, context_config=ContextConfig(), embed_dim=256, num_layers=12,
                 global_layer_norm=False, use_mlp=True, mlp_ratio=2.0, activation_function="gelu", encoder_layers=1, pos_embed=True,
                 big_skip=True, filter_residual=False, filter_output=False, normalize_big_skip=False, affine_norms=False,
                 filter_num_groups=1, filter_preserves_global_mean=False, data_grid="equiangular", hard_thresholding_fraction=1.0,
                 lora_rank=0, lora_alpha=None, spectral_lora_rank=0, spectral_lora_alpha=None, spectral_ratio=1.0,
                 clip_latent_global_means=False):
        super().__init__()
        h, w = img_shape
        modes_lat, modes_lon = int(h * hard_thresholding_fraction), int((w // 2 + 1) * hard_thresholding_fraction)
        # get_lat_lon_sfnonet (sfnonet.py:443-493): first / last transforms on the data grid, the rest on Legendre-Gauss
        self.trans_down = RealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid=data_grid)
        self.itrans_up = InverseRealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid=data_grid)
        self.trans = RealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid="legendre-gauss")
        self.itrans = InverseRealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid="legendre-gauss")
        self.big_skip, self.filter_residual, self.filter_output = big_skip, filter_residual, filter_output
        act_layer = {"relu": nn.ReLU, "gelu": nn.GELU, "silu": nn.SiLU}[activation_function]

        def stack(cin, cout):
            mods, cur = [], cin
            for _ in range(encoder_layers):
                mods += [LoRAConv2d(cur, embed_dim, 1, bias=True, lora_rank=lora_rank, lora_alpha=lora_alpha), act_layer()]
                cur = embed_dim
            mods.append(LoRAConv2d(cur, cout, 1, bias=False, lora_rank=lora_rank, lora_alpha=lora_alpha))
            return nn.Sequential(*mods)

        self.encoder = stack(in_chans, embed_dim)
        self.blocks = nn.ModuleList([
            FourierNeuralOperatorBlock(self.trans_down if i == 0 else self.trans, self.itrans_up if i == num_layers - 1 else self.itrans,
                                       embed_dim, img_shape, context_config, global_layer_norm, mlp_ratio, act_layer, use_mlp,
                                       filter_residual, affine_norms, filter_num_groups, filter_preserves_global_mean, lora_rank,
                                       lora_alpha, spectral_lora_rank, spectral_lora_alpha, spectral_ratio)
            for i in range(num_layers)])
        self.decoder = stack(embed_dim + big_skip * in_chans, out_chans)
        if pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim, h, w))
            trunc_normal_(self.pos_embed, std=0.02)
        else:
            self.pos_embed = None
        if normalize_big_skip:
            self.norm_big_skip = ConditionalLayerNorm(in_chans, img_shape, context_config, global_layer_norm, elementwise_affine=affine_norms)
        else:
            self.norm_big_skip = None
        self._clip_latent_global_means = clip_latent_global_means
        if clip_latent_global_means:  # sfnonet.py:730-746
            self.register_buffer("_gm_min", torch.full((1, embed_dim, 1, 1), float("inf")))
            self.register_buffer("_gm_max", torch.full((1, embed_dim, 1, 1), float("-inf")))

    def forward(self, x, context):
        if self.big_skip:
            residual = self.itrans_up(self.trans_down(x)) if self.filter_residual else x
            if self.norm_big_skip is not None:
                residual = self.norm_big_skip(residual, context)
        x = self.encoder(x)
        if self.pos_embed is not None:
            x = x + self.pos_embed
        if self._clip_latent_global_means and torch.isfinite(self._gm_max).all():  # eval branch of sfnonet.py:792-812
            global_means = x.mean(dim=(-2, -1), keepdim=True)
            x = x + (torch.clamp(global_means, min=self._gm_min, max=self._gm_max) - global_means)
        for blk in self.blocks:
            x = blk(x, context)
        if self.big_skip:
            x = torch.cat((x, residual), dim=1)
        x = self.decoder(x)
        if self.filter_output:
            x = self.itrans_up(self.trans_down(x))
        return x


def isotropic_noise_from_normals(real, imag, lmax, isht):
    """stochastic_sfno.py:21-47 given its two N(0,1) draws ``real``, ``imag`` [..., lmax, mmax] (the draws are the caller's)."""
    real, imag = real.clone(), imag.clone()
    imag[..., :, 0] = 0.0
    real[..., :, 1:] /= math.sqrt(2.0)
    imag[..., :, 1:] /= math.sqrt(2.0)
    return isht((real + 1j * imag) * (math.sqrt(4.0 * math.pi) / lmax))


class NoiseConditionedModel(nn.Module):
    """stochastic_sfno.py:50-172: draws the noise, builds the Context, calls the conditional net.

    ``forward(x, labels=None, noise=None)``: ``noise`` overrides the draw (tests inject the same noise on both sides)."""

    def __init__(self, conditional_model, img_shape, embed_dim_noise=256, embed_dim_pos=0, n_labels=0, label_embed_dim=0,
                 isotropic=False):
        super().__init__()
        self.conditional_model = conditional_model
        self.embed_dim, self.img_shape, self.isotropic = embed_dim_noise, img_shape, isotropic
        self.label_embedding = nn.Linear(n_labels, label_embed_dim) if label_embed_dim > 0 else None
        eff = label_embed_dim if label_embed_dim > 0 else n_labels
        self.label_pos_embed = None
        if embed_dim_pos != 0:
            self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim_pos, *img_shape))
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
            if eff > 0:
                self.label_pos_embed = nn.Parameter(torch.zeros(eff, embed_dim_pos, *img_shape))
                nn.init.trunc_normal_(self.label_pos_embed, std=0.02)
        else:
            self.pos_embed = None

    def draw_noise(self, batch):
        if self.isotropic:
            isht = self.conditional_model.itrans_up
            shape = (batch, self.embed_dim, isht.lmax, isht.mmax)
            return isotropic_noise_from_normals(torch.randn(shape), torch.randn(shape), isht.lmax, isht)
        return torch.randn(batch, self.embed_dim, *self.img_shape)

    def forward(self, x, labels=None, noise=None):
        x = x.reshape(-1, *x.shape[-3:])
        if noise is None:
            noise = self.draw_noise(x.shape[0])
        if labels is not None and self.label_embedding is not None:
            labels = self.label_embedding(labels)
        pos = None
        if self.pos_embed is not None:
            pos = self.pos_embed.repeat(x.shape[0], 1, 1, 1)
            if self.label_pos_embed is not None and labels is not None:
                pos = pos + torch.einsum("bl,lpxy->bpxy", labels, self.label_pos_embed)
        return self.conditional_model(x, Context(embedding_scalar=None, embedding_pos=pos, labels=labels, noise=noise))
