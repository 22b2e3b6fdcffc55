"""SphericalFourierNeuralOperatorNet forward -- oracle restatement on torch-CPU.

TEST INFRASTRUCTURE.  Restates, for the configuration family ACE/ACE2 uses
(spectral_transform="sht", filter_type="linear", operator_type in
{"diagonal", "dhconv"}, instance_norm / none, use_mlp, pos_embed, big_skip,
scale_factor=1, residual_filter_factor=1):

  /root/reference/fme/ace/models/modulus/sfnonet.py:123-252   (block)
  /root/reference/fme/ace/models/modulus/sfnonet.py:341-749   (net init + forward)
  /root/reference/fme/ace/models/modulus/s2convolutions.py:47-197 (SpectralConvS2)
  /root/reference/fme/ace/models/modulus/contractions.py:170-195 (diagonal / dhconv)
  /root/reference/fme/ace/models/modulus/layers.py:97-137      (MLP)
  /root/reference/fme/ace/models/modulus/initialization.py     (trunc_normal_)

Submodules are created in the same order as the reference creates them, with the
same torch layers, so that (a) ``state_dict()`` keys/shapes are identical to the
reference's and (b) construction under a given ``torch.manual_seed`` draws the
same random numbers -- which is what lets ``tests/`` check this file against the
reference's stored golden ``test_sfnonet_output_is_unchanged.pt`` without the
reference being present.
"""
import math

import torch
import torch.nn as nn

from .sht import InverseRealSHT, RealSHT


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Truncated normal by inverse-CDF of a uniform draw (initialization.py:23-76)."""

    def cdf(v):
        return (1.0 + math.erf(v / math.sqrt(2.0))) / 2.0

    with torch.no_grad():
        lo, hi = cdf((a - mean) / std), cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.0))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


def contract_dhconv(x, w):
    """out[b,o,l,m] = sum_i x[b,i,l,m] * w[i,o,l]  (contractions.py:184-195); w real [...,2]."""
    return torch.einsum("bixy,iox->boxy", x, torch.view_as_complex(w))


def contract_diagonal(x, w):
    """out[b,o,l,m] = sum_i x[b,i,l,m] * w[i,o,l,m]  (contractions.py:170-180)."""
    return torch.einsum("bixy,ioxy->boxy", x, torch.view_as_complex(w))


class SpectralConvS2(nn.Module):
    """s2convolutions.py:47-197, dense complex weights only."""

    def __init__(self, forward_transform, inverse_transform, in_channels, out_channels, operator_type, bias=True):
        super().__init__()
        scale = 1 / (in_channels * out_channels)
        self.forward_transform = forward_transform
        self.inverse_transform = inverse_transform
        self.modes_lat = inverse_transform.lmax
        self.modes_lon = inverse_transform.mmax
        self.scale_residual = (
            forward_transform.nlat != inverse_transform.nlat
            or forward_transform.nlon != inverse_transform.nlon
            or forward_transform.grid != inverse_transform.grid
        )
        self.operator_type = operator_type
        shape = [in_channels, out_channels]
        if operator_type == "diagonal":
            shape += [self.modes_lat, self.modes_lon]
        elif operator_type == "dhconv":
            shape += [self.modes_lat]
        else:
            raise ValueError(f"Unsupported operator type f{operator_type}")
        self.weight = nn.Parameter(scale * torch.randn(*shape, 2))
        if bias:
            self.bias = nn.Parameter(scale * torch.zeros(1, out_channels, 1, 1))

    def forward(self, x):
        dtype = x.dtype
        residual = x
        x = self.forward_transform(x.float())
        if self.scale_residual:
            residual = self.inverse_transform(x.contiguous()).to(dtype)
        contract = contract_dhconv if self.operator_type == "dhconv" else contract_diagonal
        xp = torch.zeros_like(x)
        xp[..., : self.modes_lat, : self.modes_lon] = contract(x[..., : self.modes_lat, : self.modes_lon], self.weight)
        x = self.inverse_transform(xp.contiguous())
        if hasattr(self, "bias"):
            x = x + self.bias
        return x.type(dtype), residual


class _FilterLayer(nn.Module):
    """SpectralFilterLayer, linear branch (sfnonet.py:45-120); keeps the ``filter.filter`` nesting."""

    def __init__(self, forward_transform, inverse_transform, embed_dim, operator_type):
        super().__init__()
        self.filter = SpectralConvS2(forward_transform, inverse_transform, embed_dim, embed_dim, operator_type, bias=True)

    def forward(self, x):
        return self.filter(x)


class _MLP(nn.Module):
    """layers.py:97-137 (no dropout, no checkpointing)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        fc1 = nn.Conv2d(in_features, hidden_features, 1, bias=True)
        fc2 = nn.Conv2d(hidden_features, in_features, 1, bias=True)
        self.fwd = nn.Sequential(fc1, nn.GELU(), fc2)

    def forward(self, x):
        return self.fwd(x)


class Block(nn.Module):
    """FourierNeuralOperatorBlock (sfnonet.py:123-252), inner_skip="linear", outer_skip="identity"."""

    def __init__(self, forward_transform, inverse_transform, embed_dim, operator_type, mlp_ratio, norm_layer, use_mlp):
        super().__init__()
        self.norm0 = norm_layer()
        self.filter = _FilterLayer(forward_transform, inverse_transform, embed_dim, operator_type)
        self.inner_skip = nn.Conv2d(embed_dim, embed_dim, 1, 1)
        self.act_layer = nn.GELU()
        self.norm1 = norm_layer()
        if use_mlp:
            self.mlp = _MLP(embed_dim, int(embed_dim * mlp_ratio))

    def forward(self, x):
        x_norm = self.norm0(x)
        x, residual = self.filter(x_norm)
        x = x + self.inner_skip(residual)
        x = self.act_layer(x)
        x = self.norm1(x)
        if hasattr(self, "mlp"):
            x = self.mlp(x)
        return x + residual


class SphericalFourierNeuralOperatorNet(nn.Module):
    """Oracle of the modulus SFNO as configured by fme/ace/registry/sfno.py:21-61."""

    def __init__(
        self,
        img_shape,
        in_chans,
        out_chans,
        embed_dim=256,
        num_layers=12,
        operator_type="diagonal",
        scale_factor=1,
        hard_thresholding_fraction=1.0,
        normalization_layer="instance_norm",
        use_mlp=True,
        mlp_ratio=2.0,
        encoder_layers=1,
        pos_embed=True,
        big_skip=True,
        data_grid="legendre-gauss",
    ):
        super().__init__()
        assert scale_factor == 1, "oracle covers scale_factor == 1 only (all ACE configs)"
        self.img_shape = tuple(img_shape)
        self.big_skip = big_skip
        h, w = self.img_shape
        modes_lat = int(h * hard_thresholding_fraction)
        modes_lon = int((w // 2 + 1) * hard_thresholding_fraction)

        # sfnonet.py:499-515
        self.trans_down = RealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid=data_grid)
        self.itrans_up = InverseRealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid=data_grid)
        self.trans = RealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid="legendre-gauss")
        self.itrans = InverseRealSHT(h, w, lmax=modes_lat, mmax=modes_lon, grid="legendre-gauss")

        # sfnonet.py:562-577
        enc, cur = [], in_chans
        for _ in range(encoder_layers):
            enc += [nn.Conv2d(cur, embed_dim, 1, bias=True), nn.GELU()]
            cur = embed_dim
        enc.append(nn.Conv2d(cur, embed_dim, 1, bias=False))
        self.encoder = nn.Sequential(*enc)

        if normalization_layer == "instance_norm":
            def norm_layer():
                return nn.InstanceNorm2d(num_features=embed_dim, eps=1e-6, affine=True, track_running_stats=False)
        elif normalization_layer == "none":
            norm_layer = nn.Identity
        else:
            raise NotImplementedError(normalization_layer)

        # sfnonet.py:603-657
        self.blocks = nn.ModuleList()
        for i in range(num_layers):
            fwd = self.trans_down if i == 0 else self.trans
            inv = self.itrans_up if i == num_layers - 1 else self.itrans
            self.blocks.append(Block(fwd, inv, embed_dim, operator_type, mlp_ratio, norm_layer, use_mlp))

        # sfnonet.py:659-671
        dec, cur = [], embed_dim + big_skip * in_chans
        for _ in range(encoder_layers):
            dec += [nn.Conv2d(cur, embed_dim, 1, bias=True), nn.GELU()]
            cur = embed_dim
        dec.append(nn.Conv2d(cur, out_chans, 1, bias=False))
        self.decoder = nn.Sequential(*dec)

        # sfnonet.py:673-685
        if pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim, h, w))
            trunc_normal_(self.pos_embed, std=0.02)

        # sfnonet.py:687-697
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        residual = x  # residual_filter_factor == 1 -> identity filters (sfnonet.py:481-483, :716)
        x = self.encoder(x)
        if hasattr(self, "pos_embed"):
            x = x + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        if self.big_skip:
            x = torch.cat((x, residual), dim=1)
        return self.decoder(x)
