"""``SphericalFourierNeuralOperatorNet`` whose forward pass runs in libace_b200 (sm_100a CUDA).

Drop-in for ``/root/reference/fme/ace/models/modulus/sfnonet.py:255-749`` as it is configured by
``SphericalFourierNeuralOperatorBuilder`` (``fme/ace/registry/sfno.py:14-61``):

* same constructor signature and ``params``-overrides-kwargs rule (sfnonet.py:341-463);
* same parameter names, shapes, dtypes and creation order, so ``state_dict()`` round-trips with
  reference checkpoints and construction under a seed draws the same initial weights;
* ``forward(x[B, C_in, H, W]) -> [B, C_out, H, W]`` fp32, any batch size.

The torch sub-modules below only *hold* the parameters (they are never called); the device
library keeps its own re-laid-out copy (split-bf16 planes, real-ified dhconv weights), which is
refreshed whenever a parameter's storage or version counter changes (``load_state_dict``,
``.to()``, in-place edits).  Inference only: autograd through the kernels is not implemented, so
calling with gradients enabled on trainable parameters raises instead of silently detaching.

Unsupported reference options raise ``NotImplementedError`` at construction (no fallback):
``spectral_transform="fft"``, ``filter_type="non-linear"``, ``scale_factor != 1``,
``residual_filter_factor != 1``, factorized / separable weights, ``layer_norm``,
``use_mlp=False``, ``encoder_layers != 1``, dropout, activation other than GELU.
"""
import ctypes
import math
from typing import Any, Tuple

import torch
import torch.nn as nn

from . import _lib
from .sht import InverseRealSHT, RealSHT, ShtPlan


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Inverse-CDF truncated normal, same draw sequence as the reference's initialization.py:23-76."""
    lo = (1.0 + math.erf((a - mean) / std / math.sqrt(2.0))) / 2.0
    hi = (1.0 + math.erf((b - mean) / std / math.sqrt(2.0))) / 2.0
    with torch.no_grad():
        tensor.uniform_(2 * lo - 1, 2 * hi - 1).erfinv_().mul_(std * math.sqrt(2.0)).add_(mean).clamp_(min=a, max=b)
    return tensor


class _SpectralConvS2(nn.Module):
    """Parameter holder for SpectralConvS2 (s2convolutions.py:47-160): ``weight`` and ``bias``."""

    def __init__(self, channels, modes_lat, modes_lon, operator_type):
        super().__init__()
        scale = 1 / (channels * channels)
        shape = [channels, channels, modes_lat] + ([modes_lon] if operator_type == "diagonal" else [])
        self.weight = nn.Parameter(scale * torch.randn(*shape, 2))
        self.bias = nn.Parameter(scale * torch.zeros(1, channels, 1, 1))


class _FilterLayer(nn.Module):
    def __init__(self, channels, modes_lat, modes_lon, operator_type):
        super().__init__()
        self.filter = _SpectralConvS2(channels, modes_lat, modes_lon, operator_type)


class _MLP(nn.Module):
    def __init__(self, channels, hidden):
        super().__init__()
        self.fwd = nn.Sequential(nn.Conv2d(channels, hidden, 1, bias=True), nn.GELU(), nn.Conv2d(hidden, channels, 1, bias=True))


class _Block(nn.Module):
    """Parameter holder for FourierNeuralOperatorBlock (sfnonet.py:126-215), same registration order."""

    def __init__(self, channels, hidden, modes_lat, modes_lon, operator_type, norm_layer):
        super().__init__()
        self.norm0 = norm_layer()
        self.filter = _FilterLayer(channels, modes_lat, modes_lon, operator_type)
        self.inner_skip = nn.Conv2d(channels, channels, 1, 1)
        self.norm1 = norm_layer()
        self.mlp = _MLP(channels, hidden)


def _pick(params, name, default):
    return getattr(params, name) if hasattr(params, name) else default


class SphericalFourierNeuralOperatorNet(nn.Module):
    def __init__(
        self,
        params,
        spectral_transform: str = "sht",
        filter_type: str = "linear",
        operator_type: str = "diagonal",
        img_shape: Tuple[int, int] = (721, 1440),
        scale_factor: int = 1,
        residual_filter_factor: int = 1,
        in_chans: int = 2,
        out_chans: int = 2,
        embed_dim: int = 256,
        num_layers: int = 12,
        use_mlp: int = True,
        mlp_ratio: float = 2.0,
        activation_function: str = "gelu",
        encoder_layers: int = 1,
        pos_embed: bool = True,
        drop_rate: float = 0.0,
        drop_path_rate: float = 0.0,
        num_blocks: int = 16,
        sparsity_threshold: float = 0.0,
        normalization_layer: str = "instance_norm",
        hard_thresholding_fraction: float = 1.0,
        use_complex_kernels: bool = True,
        big_skip: bool = True,
        rank: float = 1.0,
        factorization: Any = None,
        separable: bool = False,
        complex_network: bool = True,
        complex_activation: str = "real",
        spectral_layers: int = 3,
        checkpointing: int = 0,
    ):
        super().__init__()
        self.params = params
        p = params
        self.spectral_transform = _pick(p, "spectral_transform", spectral_transform)
        self.filter_type = _pick(p, "filter_type", filter_type)
        self.operator_type = _pick(p, "operator_type", operator_type)
        self.img_shape = (
            (p.img_shape_x, p.img_shape_y) if hasattr(p, "img_shape_x") and hasattr(p, "img_shape_y") else tuple(img_shape)
        )
        self.scale_factor = _pick(p, "scale_factor", scale_factor)
        self.residual_filter_factor = _pick(p, "residual_filter_factor", residual_filter_factor)
        self.in_chans = _pick(p, "N_in_channels", in_chans)
        self.out_chans = _pick(p, "N_out_channels", out_chans)
        self.embed_dim = self.num_features = _pick(p, "embed_dim", embed_dim)
        self.num_layers = _pick(p, "num_layers", num_layers)
        self.hard_thresholding_fraction = _pick(p, "hard_thresholding_fraction", hard_thresholding_fraction)
        self.normalization_layer = _pick(p, "normalization_layer", normalization_layer)
        self.use_mlp = _pick(p, "use_mlp", use_mlp)
        activation = _pick(p, "activation_function", activation_function)
        self.encoder_layers = _pick(p, "encoder_layers", encoder_layers)
        use_pos_embed = _pick(p, "pos_embed", pos_embed)
        self.big_skip = _pick(p, "big_skip", big_skip)
        self.factorization = _pick(p, "factorization", factorization)
        self.separable = _pick(p, "separable", separable)
        self.checkpointing = _pick(p, "checkpointing", checkpointing)
        data_grid = _pick(p, "data_grid", "equiangular")
        self.data_grid = data_grid

        unsupported = []
        if self.spectral_transform != "sht":
            unsupported.append(f"spectral_transform={self.spectral_transform!r}")
        if self.filter_type != "linear":
            unsupported.append(f"filter_type={self.filter_type!r}")
        if self.operator_type not in ("diagonal", "dhconv"):
            raise ValueError(f"Unsupported operator type f{self.operator_type}")
        if self.scale_factor != 1:
            unsupported.append(f"scale_factor={self.scale_factor}")
        if self.residual_filter_factor != 1:
            unsupported.append(f"residual_filter_factor={self.residual_filter_factor}")
        if self.factorization is not None or self.separable:
            unsupported.append("factorized / separable spectral weights")
        if self.normalization_layer not in ("instance_norm", "none"):
            if self.normalization_layer == "layer_norm":
                unsupported.append("normalization_layer='layer_norm'")
            else:
                raise NotImplementedError(f"Error, normalization {self.normalization_layer} not implemented.")
        if not self.use_mlp:
            unsupported.append("use_mlp=False")
        if activation != "gelu":
            if activation not in ("relu", "silu"):
                raise ValueError(f"Unknown activation function {activation}")
            unsupported.append(f"activation_function={activation!r}")
        if self.encoder_layers != 1:
            unsupported.append(f"encoder_layers={self.encoder_layers}")
        if drop_rate > 0.0 or drop_path_rate > 0.0:
            unsupported.append("dropout")
        if unsupported:
            raise NotImplementedError("ace_b200 SFNO does not implement: " + ", ".join(unsupported))

        self.h = int(self.img_shape[0] // self.scale_factor)
        self.w = int(self.img_shape[1] // self.scale_factor)
        self.modes_lat = int(self.h * self.hard_thresholding_fraction)
        self.modes_lon = int((self.w // 2 + 1) * self.hard_thresholding_fraction)
        self.mlp_hidden = int(self.embed_dim * mlp_ratio)

        # transforms (attribute bags; sfnonet.py:499-515).  trans_down/itrans_up live on the data grid,
        # trans/itrans on the Gauss grid; their device tables are created lazily on first forward.
        kw = dict(lmax=self.modes_lat, mmax=self.modes_lon)
        self.trans_down = RealSHT(*self.img_shape, grid=data_grid, **kw).float()
        self.itrans_up = InverseRealSHT(*self.img_shape, grid=data_grid, **kw).float()
        self.trans = RealSHT(self.h, self.w, grid="legendre-gauss", **kw).float()
        self.itrans = InverseRealSHT(self.h, self.w, grid="legendre-gauss", **kw).float()

        C = self.embed_dim
        self.encoder = nn.Sequential(nn.Conv2d(self.in_chans, C, 1, bias=True), nn.GELU(), nn.Conv2d(C, C, 1, bias=False))

        if self.normalization_layer == "instance_norm":
            def norm_layer():
                return nn.InstanceNorm2d(num_features=C, eps=1e-6, affine=True, track_running_stats=False)
        else:
            norm_layer = nn.Identity

        self.blocks = nn.ModuleList(
            [_Block(C, self.mlp_hidden, self.modes_lat, self.modes_lon, self.operator_type, norm_layer) for _ in range(self.num_layers)]
        )

        dec_in = C + self.big_skip * self.in_chans
        self.decoder = nn.Sequential(nn.Conv2d(dec_in, C, 1, bias=True), nn.GELU(), nn.Conv2d(C, self.out_chans, 1, bias=False))

        if use_pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, C, self.img_shape[0], self.img_shape[1]))
            self.pos_embed.is_shared_mp = ["matmul"]
            trunc_normal_(self.pos_embed, std=0.02)

        # sfnonet.py:687-697
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

        self._net = None          # ctypes handle of the ace_sfno object
        self._net_device = None
        self._plans = None
        self._uploaded = {}       # param name -> (data_ptr, version) last sent to the library
        self._offloaded = False   # offload_parameters(): the torch parameters live on the host, the library's copies on the GPU

    # ------------------------------------------------------------------ device object management
    def _config(self):
        return _lib.SfnoConfig(
            img_h=self.img_shape[0], img_w=self.img_shape[1], in_chans=self.in_chans, out_chans=self.out_chans,
            embed_dim=self.embed_dim, num_layers=self.num_layers, lmax=self.modes_lat, mmax=self.modes_lon,
            mlp_hidden=self.mlp_hidden, operator_type=1 if self.operator_type == "dhconv" else 0,
            normalization=1 if self.normalization_layer == "instance_norm" else 0,
            pos_embed=1 if hasattr(self, "pos_embed") else 0, big_skip=1 if self.big_skip else 0, norm_eps=1e-6,
        )

    def _release(self):
        if getattr(self, "_net", None) is not None:
            try:
                _lib.load().ace_sfno_destroy(self._net)
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
        lib = _lib.load()
        outer = ShtPlan.get(*self.img_shape, self.modes_lat, self.modes_lon, self.data_grid)
        inner = ShtPlan.get(self.h, self.w, self.modes_lat, self.modes_lon, "legendre-gauss")
        cfg = self._config()
        handle = ctypes.c_void_p()
        _lib.check(lib.ace_sfno_create(ctypes.byref(cfg), outer.handle, inner.handle, ctypes.byref(handle)))
        self._net, self._net_device, self._plans = handle, device, (outer, inner)

    def _sync_params(self, stream):
        lib = _lib.load()
        dirty = False
        for name, prm in self.named_parameters():
            key = (prm.data_ptr(), prm._version)
            if self._uploaded.get(name) == key:
                continue
            if prm.device != self._net_device:
                if not (self._offloaded and prm.device.type == "cpu"):
                    raise _lib.AceError(f"parameter {name} is on {prm.device}, input is on {self._net_device}")
                # offload_parameters(): the edited host copy is staged through a temporary device tensor (stream-ordered free)
                t = prm.detach().to(self._net_device)
            else:
                t = prm.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            _lib.check(lib.ace_sfno_set_param(self._net, name.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), stream))
            self._uploaded[name] = key
            dirty = True
        if dirty:
            _lib.check(lib.ace_sfno_finalize(self._net))

    def offload_parameters(self):
        """Single weight residency on the GPU: after the parameters have reached the device library (split-bf16 planes, the form
        every kernel reads), move the torch fp32 parameters to host memory.  ``state_dict`` / ``load_state_dict`` / in-place edits keep
        working on the host copies (an edited parameter is re-uploaded before the next forward); ``.cuda()`` brings them back.
        Frees the fp32 parameter bytes on the device: 1.8 GB at ACE2 1 degree, 6.9 GB at 0.25 degree."""
        if self._net is None:
            raise _lib.AceError("offload_parameters(): run one forward first (the device library has no parameters yet)")
        with torch.cuda.device(self._net_device):
            self._sync_params(_lib.current_stream_ptr())
            torch.cuda.current_stream().synchronize()
        for name, prm in self.named_parameters():
            prm.data = prm.data.cpu()
            self._uploaded[name] = (prm.data_ptr(), prm._version)
        self._offloaded = True
        return self

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        if not x.is_cuda:
            raise _lib.AceError("ace_b200 SFNO: input must be a CUDA tensor (there is no CPU path)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise _lib.AceError(
                "ace_b200 SFNO is inference-only: call under torch.no_grad() (or requires_grad_(False) on the module)"
            )
        if x.dim() != 4 or x.shape[1] != self.in_chans or tuple(x.shape[-2:]) != tuple(self.img_shape):
            raise ValueError(f"expected input [B, {self.in_chans}, {self.img_shape[0]}, {self.img_shape[1]}], got {tuple(x.shape)}")
        dtype = x.dtype
        x = x.float().contiguous()
        B = x.shape[0]
        y = torch.empty(B, self.out_chans, *self.img_shape, dtype=torch.float32, device=x.device)
        if B == 0:
            return y.to(dtype)
        with torch.cuda.device(x.device):
            self._ensure_net(x.device)
            stream = _lib.current_stream_ptr()
            self._sync_params(stream)
            _lib.check(_lib.load().ace_sfno_forward(self._net, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), B, stream))
        return y.to(dtype)

    def native_handle(self):
        """The ace_sfno* behind this module (after at least one forward); used by the fused stepper."""
        return self._net
