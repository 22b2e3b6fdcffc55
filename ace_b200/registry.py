"""Module-builder registry boundary (``ModuleSelector`` / ``ModuleConfig``) for the B200 SFNO.

Mirrors the reference's plugin API for this path:

* ``fme/core/registry/registry.py:13-59``  -- ``Registry.register(name)`` / ``get(name, config)``
* ``fme/core/registry/module.py:18-58``    -- ``ModuleConfig`` (dataclass with ``build`` / ``from_state``)
* ``fme/core/registry/module.py:121-210``  -- ``ModuleSelector(type, config).build(n_in, n_out, dataset_info)``
* ``fme/ace/registry/sfno.py:14-61``       -- ``SphericalFourierNeuralOperatorBuilder`` (field names + defaults)

``B200SphericalFourierNeuralOperatorBuilder`` has exactly the reference builder's fields, so any
YAML / checkpoint ``builder.config`` that builds the reference net builds this one.  It is
registered here under ``"B200SphericalFourierNeuralOperatorNet"``; ``install_into_fme()``
registers it in the real ``fme`` registry under that name and, with ``override=True``, also over
``"SphericalFourierNeuralOperatorNet"`` (``Registry.register`` overwrites silently,
registry.py:52) so released ACE2 checkpoints route to the B200 path unchanged.

The small ``Registry`` / ``ModuleSelector`` / ``Module`` classes below exist so the package is
usable (and testable) where ``fme`` itself is not importable; they follow the reference's
semantics: strict ``from_state`` (unknown keys raise, like dacite strict mode), config
normalised to include defaults, labels rejected for unconditional builders.
"""
import dataclasses
import os
from typing import Any, Callable, ClassVar, Literal, Mapping

import torch
from torch import nn

from .sfno import SphericalFourierNeuralOperatorNet


class Registry:
    def __init__(self):
        self._types = {}

    def register(self, type_name: str) -> Callable:
        def register_func(cls):
            self._types[type_name] = cls
            return cls

        return register_func

    def get(self, type_name: str, config: Mapping[str, Any]):
        return self._types[type_name].from_state(config)


@dataclasses.dataclass
class ModuleConfig:
    def build(self, n_in_channels: int, n_out_channels: int, dataset_info) -> nn.Module:
        raise NotImplementedError

    @classmethod
    def from_state(cls, state: Mapping[str, Any]):
        names = {f.name for f in dataclasses.fields(cls)}
        unknown = set(state) - names
        if unknown:
            raise ValueError(f"unexpected keys for {cls.__name__}: {sorted(unknown)}")
        return cls(**dict(state))


class Module:
    """fme/core/registry/module.py:69-118 for unconditional modules."""

    def __init__(self, module: nn.Module):
        self._module = module

    def __call__(self, input: torch.Tensor, labels=None) -> torch.Tensor:
        if labels is not None:
            raise TypeError("Labels are not allowed for unconditional models")
        return self._module(input)

    @property
    def torch_module(self) -> nn.Module:
        return self._module

    def get_state(self):
        return {**self._module.state_dict(), "label_encoding": None}

    def load_state(self, state):
        state = dict(state)
        state.pop("label_encoding", None)
        self._module.load_state_dict(state)

    def wrap_module(self, fn):
        return Module(fn(self._module))

    def to(self, device):
        return Module(self._module.to(device))


@dataclasses.dataclass
class ModuleSelector:
    type: str
    config: Mapping[str, Any]
    conditional: bool = False
    allow_missing_variables: bool = False
    registry: ClassVar[Registry] = Registry()

    def __post_init__(self):
        if self.conditional:
            raise ValueError(f"Conditional predictions require a conditional builder, got {self.type}")
        self._instance = self.registry.get(self.type, self.config)
        self.config = dataclasses.asdict(self._instance)

    @property
    def module_config(self):
        return self._instance

    @classmethod
    def register(cls, type_name: str):
        return cls.registry.register(type_name)

    def build(self, n_in_channels: int, n_out_channels: int, dataset_info) -> Module:
        return Module(self._instance.build(n_in_channels=n_in_channels, n_out_channels=n_out_channels, dataset_info=dataset_info))

    @classmethod
    def get_available_types(cls):
        return cls.registry._types.keys()


@dataclasses.dataclass
class DatasetInfo:
    """The two attributes of fme.core.dataset_info.DatasetInfo the SFNO builder reads."""

    img_shape: tuple
    all_labels: frozenset = frozenset()


B200_TYPE_NAME = "B200SphericalFourierNeuralOperatorNet"
REFERENCE_TYPE_NAME = "SphericalFourierNeuralOperatorNet"


@ModuleSelector.register(B200_TYPE_NAME)
@dataclasses.dataclass
class B200SphericalFourierNeuralOperatorBuilder(ModuleConfig):
    """Same fields and defaults as fme/ace/registry/sfno.py:21-42."""

    spectral_transform: str = "sht"
    filter_type: str = "linear"
    operator_type: str = "diagonal"
    scale_factor: int = 1
    residual_filter_factor: int = 1
    embed_dim: int = 256
    num_layers: int = 12
    hard_thresholding_fraction: float = 1.0
    normalization_layer: str = "instance_norm"
    use_mlp: bool = True
    activation_function: str = "gelu"
    encoder_layers: int = 1
    pos_embed: bool = True
    big_skip: bool = True
    rank: float = 1.0
    factorization: str | None = None
    separable: bool = False
    complex_network: bool = True
    complex_activation: str = "real"
    spectral_layers: int = 1
    checkpointing: int = 0
    data_grid: Literal["legendre-gauss", "equiangular"] = "legendre-gauss"

    def build(self, n_in_channels: int, n_out_channels: int, dataset_info):
        if len(dataset_info.all_labels) > 0:
            raise ValueError("SphericalFourierNeuralOperatorNet does not support labels")
        return SphericalFourierNeuralOperatorNet(
            params=self, in_chans=n_in_channels, out_chans=n_out_channels, img_shape=dataset_info.img_shape
        )


def disabled_by_env() -> bool:
    """``ACE_B200_DISABLE=1``: the A/B kill switch.  The install functions then leave the reference's builders / transforms / step in
    place and register the ``B200...`` type names as ALIASES of the reference classes, so the same config file runs on the reference
    path (for a parity or timing comparison) without editing it."""
    return os.environ.get("ACE_B200_DISABLE", "0") not in ("", "0")


def _fme_registered(selector, name):
    """The class the fme selector has registered under ``name`` (``Registry._types``, fme/core/registry/registry.py)."""
    reg = getattr(selector, "registry", None)
    types = getattr(reg, "_types", None)
    if types is None or name not in types:
        raise RuntimeError(f"ACE_B200_DISABLE: the reference type '{name}' is not registered; import its module before installing")
    return types[name]


def install_into_fme(override: bool = False):
    """Register the B200 builder with the real fme registry (needs ``fme`` importable)."""
    from fme.ace.registry.registry import ModuleConfig as FmeModuleConfig  # noqa: PLC0415
    from fme.ace.registry.registry import ModuleSelector as FmeModuleSelector  # noqa: PLC0415

    if disabled_by_env():
        for alias, ref in ((B200_TYPE_NAME, REFERENCE_TYPE_NAME), (B200_NOISE_TYPE_NAME, REFERENCE_NOISE_TYPE_NAME)):
            FmeModuleSelector.register(alias)(_fme_registered(FmeModuleSelector, ref))
        return None

    fields = [(f.name, f.type, f) for f in dataclasses.fields(B200SphericalFourierNeuralOperatorBuilder)]
    cls = dataclasses.make_dataclass(
        "B200SphericalFourierNeuralOperatorBuilder", fields, bases=(FmeModuleConfig,),
        namespace={"build": B200SphericalFourierNeuralOperatorBuilder.build},
    )
    FmeModuleSelector.register(B200_TYPE_NAME)(cls)
    if override:
        FmeModuleSelector.register(REFERENCE_TYPE_NAME)(cls)
    nfields = [(f.name, f.type, f) for f in dataclasses.fields(B200NoiseConditionedSFNOBuilder)]
    ncls = dataclasses.make_dataclass(
        "B200NoiseConditionedSFNOBuilder", nfields, bases=(FmeModuleConfig,),
        namespace={"build": B200NoiseConditionedSFNOBuilder.build, "__post_init__": B200NoiseConditionedSFNOBuilder.__post_init__},
    )
    FmeModuleSelector.register(B200_NOISE_TYPE_NAME)(ncls)
    if override:
        FmeModuleSelector.register(REFERENCE_NOISE_TYPE_NAME)(ncls)
    return cls


B200_NOISE_TYPE_NAME = "B200NoiseConditionedSFNO"
REFERENCE_NOISE_TYPE_NAME = "NoiseConditionedSFNO"


@ModuleSelector.register(B200_NOISE_TYPE_NAME)
@dataclasses.dataclass
class B200NoiseConditionedSFNOBuilder(ModuleConfig):
    """Same fields and defaults as ``NoiseConditionedSFNOBuilder`` (fme/ace/registry/stochastic_sfno.py:181-397), including the
    default ``noise_embed_dim = 256``: up to 64 channels of noise + positional context the ConditionalLayerNorm is one streaming
    kernel, beyond that a statistics pass + one tcgen05 GEMM whose epilogue normalises and modulates (csrc/cln.cu, GemmOp::cln)."""

    spectral_transform: str = "sht"
    filter_type: str = "linear"
    operator_type: str = "dhconv"
    residual_filter_factor: int = 1
    embed_dim: int = 256
    noise_embed_dim: int = 256
    context_pos_embed_dim: int = 0
    label_embed_dim: int = 0
    noise_type: Literal["isotropic", "gaussian"] = "gaussian"
    global_layer_norm: bool = False
    num_layers: int = 12
    use_mlp: bool = True
    mlp_ratio: float = 2.0
    activation_function: str = "gelu"
    encoder_layers: int = 1
    pos_embed: bool = True
    big_skip: bool = True
    rank: float = 1.0
    factorization: None = None
    separable: bool = False
    complex_network: bool = True
    complex_activation: str = "real"
    spectral_layers: int = 1
    checkpointing: int = 0
    data_grid: Literal["legendre-gauss", "equiangular"] = "legendre-gauss"
    filter_residual: bool = False
    filter_output: bool = False
    local_blocks: list | None = None
    normalize_big_skip: bool = False
    affine_norms: bool = False
    filter_num_groups: int = 1
    lora_rank: int = 0
    lora_alpha: float | None = None
    spectral_lora_rank: int = 0
    spectral_lora_alpha: float | None = None
    filter_preserves_global_mean: bool = False
    spectral_ratio: float = 1.0
    clip_latent_global_means: bool = False

    def __post_init__(self):
        # stochastic_sfno.py:300-318
        if self.context_pos_embed_dim > 0 and self.pos_embed:
            raise ValueError("context_pos_embed_dim and pos_embed should not both be set")
        if self.factorization is not None:
            raise ValueError("The 'factorization' parameter is no longer supported.")
        if self.separable:
            raise ValueError("The 'separable' parameter is no longer supported.")
        if self.operator_type != "dhconv":
            raise ValueError("Only 'dhconv' operator_type is supported for NoiseConditionedSFNO models.")

    def build(self, n_in_channels: int, n_out_channels: int, dataset_info):
        from .csfno import ContextConfig, NoiseConditionedModel, SFNONetConfig, get_lat_lon_sfnonet  # noqa: PLC0415

        n_labels = len(dataset_info.all_labels)
        effective_label_dim = self.label_embed_dim if self.label_embed_dim > 0 else n_labels
        names = {f.name for f in dataclasses.fields(SFNONetConfig)}
        cfg = SFNONetConfig(**{k: v for k, v in dataclasses.asdict(self).items() if k in names})
        net = get_lat_lon_sfnonet(
            params=cfg, in_chans=n_in_channels, out_chans=n_out_channels, img_shape=dataset_info.img_shape, data_grid=self.data_grid,
            context_config=ContextConfig(embed_dim_scalar=0, embed_dim_pos=self.context_pos_embed_dim, embed_dim_noise=self.noise_embed_dim,
                                         embed_dim_labels=effective_label_dim))
        isotropic = self.noise_type == "isotropic"
        return NoiseConditionedModel(
            net, img_shape=dataset_info.img_shape, embed_dim_noise=self.noise_embed_dim, embed_dim_pos=self.context_pos_embed_dim,
            n_labels=n_labels, label_embed_dim=self.label_embed_dim, inverse_sht=net.itrans_up if isotropic else None,
            lmax=net.itrans_up.lmax if isotropic else 0, mmax=net.itrans_up.mmax if isotropic else 0)
