"""A minimal ``fme`` package for boundary tests where the real one cannot be imported (no xarray / dacite / netCDF4 / ...).

Two kinds of modules are installed into ``sys.modules`` (and removed again by the context manager):

* REAL reference files, executed from ``/root/reference`` when that tree exists -- the registry machinery itself:
  ``fme/core/registry/registry.py``, ``fme/core/registry/module.py``, ``fme/ace/registry/registry.py``, ``fme/core/labels.py``,
  ``fme/core/typing_.py``, ``fme/core/step/args.py``, ``fme/core/stepper_state.py``, ``fme/core/corrector/state.py``,
  ``fme/core/random_state.py``, ``fme/core/device.py``.  So ``ModuleSelector`` / ``ModuleConfig`` / ``Registry`` / ``StepArgs`` /
  ``StepperState`` are the reference's own code;
* SMALL stand-ins written here for what those files import but this image lacks (``dacite``: strict ``from_dict`` for nested
  dataclasses; ``fme.core.dataset_info.DatasetInfo``; ``torch_harmonics``) and for the heavy step modules
  (``fme.core.step.step``: ``StepSelector`` / ``StepABC`` over the real ``Registry``; ``fme.core.step.single_module``:
  a ``SingleModuleStepConfig`` with the reference's field list, ``fme/core/step/single_module.py:89-105``, and a
  ``SingleModuleStep`` that runs the reference's normalise -> pack -> module -> unpack -> denormalise sequence in torch;
  ``fme.core.step.output.StepOutput``; ``fme.core.normalizer.StandardNormalizer``).

Nothing here is product code; it exists so that ``install_into_fme`` / ``install_step_into_fme`` / ``patch_torch_harmonics``
execute on every CI run instead of on first contact with a real installation.
"""
import abc
import contextlib
import dataclasses
import importlib.util
import os
import sys
import types
import typing
from typing import Any, Callable, ClassVar, Mapping

import torch
from torch import nn

REF = "/root/reference"
REAL_FILES = {
    "fme.core.typing_": "fme/core/typing_.py",
    "fme.core.device": "fme/core/device.py",
    "fme.core.labels": "fme/core/labels.py",
    "fme.core.random_state": "fme/core/random_state.py",
    "fme.core.corrector.state": "fme/core/corrector/state.py",
    "fme.core.stepper_state": "fme/core/stepper_state.py",
    "fme.core.registry.registry": "fme/core/registry/registry.py",
    "fme.core.registry.module": "fme/core/registry/module.py",
    "fme.ace.registry.registry": "fme/ace/registry/registry.py",
    "fme.core.step.args": "fme/core/step/args.py",
}
PACKAGES = ["fme", "fme.core", "fme.core.registry", "fme.core.corrector", "fme.core.step", "fme.ace", "fme.ace.registry"]


def available() -> bool:
    return all(os.path.exists(os.path.join(REF, p)) for p in REAL_FILES.values())


# ---------------------------------------------------------------------------------------------- dacite stand-in
def _from_dict(data_class, data, config=None):
    """dacite.from_dict(strict=True) for (nested) dataclasses with plain / Optional / list / dict fields."""
    if not dataclasses.is_dataclass(data_class):
        return data
    if dataclasses.is_dataclass(data) and not isinstance(data, type):
        return data
    hints = typing.get_type_hints(data_class)
    names = {f.name for f in dataclasses.fields(data_class) if f.init}
    unknown = set(data) - names
    if unknown and (config is None or getattr(config, "strict", False)):
        raise ValueError(f'can not match "{sorted(unknown)}" to any data class field of {data_class.__name__}')
    kwargs = {}
    for f in dataclasses.fields(data_class):
        if not f.init or f.name not in data:
            continue
        v, t = data[f.name], hints.get(f.name, Any)
        cands = [t] + [a for a in typing.get_args(t)]
        dc = [c for c in cands if isinstance(c, type) and dataclasses.is_dataclass(c)]
        if isinstance(v, Mapping) and dc:
            v = _from_dict(dc[0], v, config)
        kwargs[f.name] = v
    return data_class(**kwargs)


@dataclasses.dataclass
class _DaciteConfig:
    strict: bool = False


# ---------------------------------------------------------------------------------------------- stand-ins for heavy modules
@dataclasses.dataclass
class DatasetInfo:
    img_shape: tuple
    all_labels: frozenset = frozenset()
    timestep: Any = None
    vertical_coordinate: Any = None
    gridded_operations: Any = None


class StandardNormalizer:
    """fme/core/normalizer.py:122-243: per-name scalars, (x - mean) / std and x * std + mean."""

    def __init__(self, means, stds):
        self.means = {k: torch.as_tensor(v, dtype=torch.float) for k, v in means.items()}
        self.stds = {k: torch.as_tensor(v, dtype=torch.float) for k, v in stds.items()}

    def normalize(self, tensors):
        return {k: (v - self.means[k].to(v.device)) / self.stds[k].to(v.device) for k, v in tensors.items() if k in self.means}

    def denormalize(self, tensors):
        return {k: v * self.stds[k].to(v.device) + self.means[k].to(v.device) for k, v in tensors.items() if k in self.means}


@dataclasses.dataclass
class NetworkAndLossNormalizationConfig:
    means: Mapping[str, float]
    stds: Mapping[str, float]

    def get_network_normalizer(self, names):
        return StandardNormalizer({n: self.means[n] for n in names}, {n: self.stds[n] for n in names})

    def load(self):
        pass


@dataclasses.dataclass
class OceanConfig:
    surface_temperature_name: str
    ocean_fraction_name: str
    interpolate: bool = False
    slab: Any = None


@dataclasses.dataclass
class AtmosphereCorrectorConfig:
    conserve_dry_air: bool = False
    zero_global_mean_moisture_advection: bool = False
    moisture_budget_correction: Any = None
    force_positive_names: list = dataclasses.field(default_factory=list)
    total_energy_budget_correction: Any = None

    def get_corrector(self, dataset_info):
        return None


def _build_step_modules(mods):
    Registry = mods["fme.core.registry.registry"].Registry
    ModuleSelector = mods["fme.core.registry.module"].ModuleSelector
    StepArgs = mods["fme.core.step.args"].StepArgs
    StepperState = mods["fme.core.stepper_state"].StepperState

    # ---- fme.core.step.output
    out_mod = types.ModuleType("fme.core.step.output")

    @dataclasses.dataclass
    class StepOutput:
        output: dict
        stepper_state: Any = None

    out_mod.StepOutput = StepOutput

    # ---- fme.core.step.step
    step_mod = types.ModuleType("fme.core.step.step")

    @dataclasses.dataclass
    class StepConfigABC(abc.ABC):
        @abc.abstractmethod
        def get_step(self, dataset_info, init_weights):
            ...

    class StepABC(abc.ABC):
        def __init__(self):
            self._training = True

        def train(self, mode=True):
            self._training = mode
            for m in self.modules:
                m.train(mode)
            return self

        def eval(self):
            return self.train(False)

        @property
        def input_names(self):
            return self.config.in_names

        @property
        def output_names(self):
            return self.config.out_names

    @dataclasses.dataclass
    class StepSelector(StepConfigABC):
        type: str
        config: dict
        registry: ClassVar[Any] = Registry()

        def __post_init__(self):
            self._step_config_instance = self.registry.get(self.type, self.config)

        @classmethod
        def register(cls, name):
            return cls.registry.register(name)

        def get_step(self, dataset_info, init_weights=lambda x: None):
            return self._step_config_instance.get_step(dataset_info, init_weights)

    step_mod.StepConfigABC, step_mod.StepABC, step_mod.StepSelector = StepConfigABC, StepABC, StepSelector

    # ---- fme.core.step.single_module
    sm_mod = types.ModuleType("fme.core.step.single_module")

    @dataclasses.dataclass
    class SingleModuleStepConfig(StepConfigABC):
        builder: ModuleSelector
        in_names: list
        out_names: list
        normalization: NetworkAndLossNormalizationConfig
        secondary_decoder: Any = None
        ocean: typing.Optional[OceanConfig] = None
        corrector: AtmosphereCorrectorConfig = dataclasses.field(default_factory=AtmosphereCorrectorConfig)
        next_step_forcing_names: list = dataclasses.field(default_factory=list)
        prescribed_prognostic_names: list = dataclasses.field(default_factory=list)
        residual_prediction: bool = False
        include_channel_mask_inputs: bool = False
        global_mean_removal: Any = None
        input_dropout: Any = None

        @classmethod
        def from_state(cls, state):
            return _from_dict(cls, state, _DaciteConfig(strict=True))

        def get_step(self, dataset_info, init_weights):
            names = list(set(self.in_names) | set(self.out_names))
            return SingleModuleStep(self, dataset_info, self.corrector.get_corrector(dataset_info),
                                    self.normalization.get_network_normalizer(names), init_weights)

    class _DummyWrapper(nn.Module):  # fme/core/distributed/non_distributed.py:15-28
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, *a, **k):
            return self.module(*a, **k)

    class SingleModuleStep(StepABC):
        """fme/core/step/single_module.py:261-449,595-665 without corrector / ocean: the torch sequence the fused step replaces."""

        def __init__(self, config, dataset_info, corrector, normalizer, init_weights):
            super().__init__()
            self._config, self._normalizer = config, normalizer
            module = config.builder.build(n_in_channels=len(config.in_names), n_out_channels=len(config.out_names),
                                          dataset_info=dataset_info)
            init_weights([module.torch_module])
            self.module = module.wrap_module(_DummyWrapper)
            self.in_names, self.out_names = config.in_names, config.out_names

        @property
        def config(self):
            return self._config

        @property
        def normalizer(self):
            return self._normalizer

        @property
        def modules(self):
            return nn.ModuleList([self.module.torch_module])

        surface_temperature_name = None
        ocean_fraction_name = None

        def prescribe_sst(self, mask_data, gen_data, target_data):
            raise RuntimeError("no ocean")

        def get_regularizer_loss(self):
            return torch.tensor(0.0)

        def get_state(self):
            return {"module": self.module.get_state()}

        def load_state(self, state):
            self.module.load_state(state["module"])

        def step(self, args, wrapper=lambda x: x):
            norm = self._normalizer.normalize(args.input)
            x = torch.stack([norm[n] for n in self.in_names], dim=-3)
            y = self.module(x, labels=args.labels)
            out = {n: y.select(-3, i) for i, n in enumerate(self.out_names)}
            if self._config.residual_prediction:
                for n in set(self.in_names) & set(self.out_names):
                    out[n] = out[n] + norm[n]
            return StepOutput(output=self._normalizer.denormalize(out), stepper_state=args.stepper_state)

    sm_mod.SingleModuleStepConfig, sm_mod.SingleModuleStep = SingleModuleStepConfig, SingleModuleStep
    sm_mod.StepSelector = StepSelector
    StepSelector.register("single_module")(SingleModuleStepConfig)
    StepSelector.register("default")(SingleModuleStepConfig)
    del StepArgs, StepperState
    return {"fme.core.step.output": out_mod, "fme.core.step.step": step_mod, "fme.core.step.single_module": sm_mod}


@contextlib.contextmanager
def installed():
    """Install the fake package; yields the dict of modules.  Restores ``sys.modules`` afterwards."""
    saved = {k: v for k, v in sys.modules.items() if k == "dacite" or k == "fme" or k.startswith("fme.") or k.startswith("torch_harmonics")}
    for k in saved:
        del sys.modules[k]
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    mods = {}
    try:
        dacite = types.ModuleType("dacite")
        dacite.from_dict, dacite.Config = _from_dict, _DaciteConfig
        sys.modules["dacite"] = dacite
        for name in PACKAGES:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = mods[name] = m
        di = types.ModuleType("fme.core.dataset_info")
        di.DatasetInfo = DatasetInfo
        sys.modules["fme.core.dataset_info"] = mods["fme.core.dataset_info"] = di
        nm = types.ModuleType("fme.core.normalizer")
        nm.StandardNormalizer, nm.NetworkAndLossNormalizationConfig = StandardNormalizer, NetworkAndLossNormalizationConfig
        sys.modules["fme.core.normalizer"] = mods["fme.core.normalizer"] = nm
        for name, rel in REAL_FILES.items():
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = mods[name] = m
            spec.loader.exec_module(m)
        for name, m in _build_step_modules(mods).items():
            sys.modules[name] = mods[name] = m
        th = types.ModuleType("torch_harmonics")
        th.RealSHT = th.InverseRealSHT = None
        sys.modules["torch_harmonics"] = mods["torch_harmonics"] = th
        yield mods
    finally:
        sys.dont_write_bytecode = old_flag
        for k in [k for k in sys.modules if k == "dacite" or k == "fme" or k.startswith("fme.") or k.startswith("torch_harmonics")]:
            del sys.modules[k]
        sys.modules.update(saved)
