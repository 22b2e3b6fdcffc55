"""A ``StepSelector``-registrable fused step for ``fme``: ``SingleModuleStepConfig``'s fields, a step object that owns a
``FusedStepper``.

Through the ``ModuleSelector`` seam alone (``registry.install_into_fme``) only the network is accelerated: the reference's
``SingleModuleStep.step`` (``fme/core/step/single_module.py:396-449,595-733``) still spends ~330 small launches per step on
normalise / pack / unpack / denormalise / corrector / ocean.  ``install_step_into_fme()`` registers a step config under
``"b200_single_module"`` (and, with ``override=True``, over ``"single_module"`` / ``"default"``, the names released
checkpoints carry; ``Registry.register`` overwrites silently, ``fme/core/registry/registry.py:52``) so that a YAML

    stepper:
      step:
        type: b200_single_module
        config: {builder: {type: B200SphericalFourierNeuralOperatorNet, config: {...}}, in_names: [...], out_names: [...], ...}

routes the WHOLE step through ``ace_stepper_step`` (one CUDA-graph-capturable call per 6-hour step).

The config class subclasses the reference's ``SingleModuleStepConfig`` (same fields, same ``from_state`` / property
behaviour, ``fme/core/step/single_module.py:48-259``); its ``get_step`` first builds the reference ``SingleModuleStep`` --
which builds the module through the builder registry, the normaliser, the ocean and the corrector exactly as the reference
does -- and wraps it in ``B200FusedStep``.  Everything that is not the per-step device work (state dict, ``modules``,
normaliser, ``prescribe_sst``, ``train`` / ``eval``) is delegated to the wrapped reference object, so checkpoints load and
save unchanged.

What the fused step takes over, and what it refuses (loudly, at construction): see ``_fusible_or_raise``.
"""
import dataclasses
from typing import Any, Callable, Mapping, Optional

import torch

from .stepper import FusedStepper

STEP_TYPE_NAME = "b200_single_module"
REFERENCE_STEP_TYPE_NAMES = ("single_module", "default")  # fme/core/step/single_module.py:48-49


def _unwrap_module(step):
    """The ``nn.Module`` behind ``SingleModuleStep.module`` (a ``Module`` wrapper whose torch module is wrapped by
    ``dist.wrap_module``: ``DummyWrapper`` / DDP expose the inner net as ``.module``; fme/core/step/single_module.py:326,345)."""
    m = step.module
    m = getattr(m, "torch_module", m)
    while hasattr(m, "module") and isinstance(getattr(m, "module"), torch.nn.Module) and not hasattr(m, "native_handle") \
            and not hasattr(m, "conditional_model"):
        m = m.module
    return m


def _scalar(v) -> float:
    return float(v.detach().cpu().reshape(-1)[0]) if isinstance(v, torch.Tensor) else float(v)


def _fusible_or_raise(cfg, net):
    """The fused step covers the reference sequence network -> [residual] -> denormalise -> ForcePositive -> conservation
    correctors -> ocean -> prescribed overwrite.  Options outside it raise instead of silently computing something else."""
    if not (hasattr(net, "native_handle") or hasattr(net, "conditional_model")):
        raise TypeError(f"b200 fused step: the builder must produce an ace_b200 network (got {type(net).__name__}); use a "
                        "B200SphericalFourierNeuralOperatorNet / B200NoiseConditionedSFNO builder or install_into_fme(override=True)")
    for attr, what in (("secondary_decoder", "a secondary decoder"), ("global_mean_removal", "global-mean removal"),
                       ("input_dropout", "input dropout")):
        if getattr(cfg, attr, None) is not None:
            raise NotImplementedError(f"b200 fused step: {what} is not part of the fused step; use the reference 'single_module' step")
    if getattr(cfg, "include_channel_mask_inputs", False):
        raise NotImplementedError("b200 fused step: include_channel_mask_inputs is not part of the fused step")


def _corrector_kwargs(cfg, dataset_info) -> (Optional[dict], list):
    """``AtmosphereCorrectorConfig`` (fme/core/corrector/atmosphere.py:223-347) -> the ``corrector=`` dict of ``FusedStepper``."""
    c = getattr(cfg, "corrector", None)
    c = getattr(c, "_instance", None) or getattr(c, "config_instance", None) or c  # CorrectorSelector -> its config
    if c is None:
        return None, []
    fp = getattr(c, "force_positive_names", None) or []
    fp = [n for n in fp if n in cfg.out_names]
    moisture = getattr(c, "moisture_budget_correction", None)
    energy = getattr(c, "total_energy_budget_correction", None)
    on = bool(getattr(c, "conserve_dry_air", False)) or bool(getattr(c, "zero_global_mean_moisture_advection", False)) or \
        moisture is not None or energy is not None
    if not on:
        return None, fp
    vc = getattr(dataset_info, "vertical_coordinate", None)
    ops = getattr(dataset_info, "gridded_operations", None)
    ak, bk = getattr(vc, "ak", None), getattr(vc, "bk", None)
    area = getattr(ops, "area_weights", None)
    if area is None and ops is not None and hasattr(ops, "get_initialization_kwargs"):
        area = ops.get_initialization_kwargs().get("area_weights")
    if ak is None or bk is None or area is None:
        raise NotImplementedError("b200 fused step: the conservation correctors need dataset_info.vertical_coordinate (ak, bk) and "
                                  "lat-lon area weights")
    ts = getattr(dataset_info, "timestep", None)
    kw = dict(conserve_dry_air=bool(getattr(c, "conserve_dry_air", False)), ak=ak, bk=bk, area_weights=area,
              timestep_seconds=ts.total_seconds() if ts is not None else 21600.0)
    if moisture is not None:
        kw["moisture_budget_correction"] = moisture
    if getattr(c, "zero_global_mean_moisture_advection", False):
        kw["zero_global_mean_moisture_advection"] = True
    if energy is not None:
        kw["total_energy_budget_correction"] = dict(method=getattr(energy, "method", "constant_temperature"),
                                                    constant_unaccounted_heating=getattr(energy, "constant_unaccounted_heating", 0.0))
    return kw, fp


def _ocean_kwargs(cfg) -> Optional[dict]:
    o = getattr(cfg, "ocean", None)
    if o is None:
        return None
    kw = dict(surface_temperature_name=o.surface_temperature_name, ocean_fraction_name=o.ocean_fraction_name,
              interpolate=bool(getattr(o, "interpolate", False)))
    slab = getattr(o, "slab", None)
    if slab is not None:
        kw["slab"] = dict(mixed_layer_depth_name=slab.mixed_layer_depth_name, q_flux_name=slab.q_flux_name)
    return kw


def make_fused_step_class(StepABC, StepOutput, StepperState=None, CorrectorState=None):
    """``B200FusedStep`` bound to the installed ``fme``'s ABCs (``fme/core/step/step.py:246``, ``output.py:12``)."""

    class B200FusedStep(StepABC):
        def __init__(self, inner, dataset_info=None):
            super().__init__()
            self._inner = inner
            cfg = inner.config
            net = _unwrap_module(inner)
            _fusible_or_raise(cfg, net)
            nrm = inner.normalizer
            names = set(cfg.in_names) | set(cfg.out_names)
            means = {n: _scalar(nrm.means[n]) for n in names}
            stds = {n: _scalar(nrm.stds[n]) for n in names}
            corrector, force_positive = _corrector_kwargs(cfg, dataset_info)
            ocean = _ocean_kwargs(cfg)
            if ocean is not None and "slab" in ocean:
                ts = getattr(dataset_info, "timestep", None)
                ocean["slab"]["timestep_seconds"] = ts.total_seconds() if ts is not None else 21600.0
            self._fused = FusedStepper(
                net, cfg.in_names, cfg.out_names, means, stds, residual_prediction=bool(cfg.residual_prediction),
                force_positive_names=force_positive, ocean=ocean, corrector=corrector,
                next_step_forcing_names=list(cfg.next_step_forcing_names),
                prescribed_prognostic_names=list(cfg.prescribed_prognostic_names))

        # ---- everything but the device work: the wrapped reference step
        @property
        def config(self):
            return self._inner.config

        @property
        def modules(self):
            return self._inner.modules

        @property
        def normalizer(self):
            return self._inner.normalizer

        @property
        def surface_temperature_name(self):
            return self._inner.surface_temperature_name

        @property
        def ocean_fraction_name(self):
            return self._inner.ocean_fraction_name

        def prescribe_sst(self, mask_data, gen_data, target_data):
            return self._inner.prescribe_sst(mask_data, gen_data, target_data)

        def get_regularizer_loss(self):
            return self._inner.get_regularizer_loss()

        def get_state(self):
            return self._inner.get_state()

        def load_state(self, state):
            self._inner.load_state(state)  # FusedStepper re-uploads edited parameters before its next step / graph replay

        @property
        def fused_stepper(self) -> FusedStepper:
            """The ``FusedStepper`` (``rollout`` / ``rollout_host`` / ``predict*`` with CUDA-graph replay per step)."""
            return self._fused

        # ---- the per-step device work
        def step(self, args, wrapper: Callable = lambda x: x):
            if getattr(args, "labels", None) is not None:
                raise NotImplementedError("b200 fused step: batch labels are not supported in the fused step")
            if getattr(args, "data_mask", None) is not None:
                raise NotImplementedError("b200 fused step: variable masks are not supported in the fused step")
            if self._training and any(p.requires_grad for p in self._fused.module.parameters()) and torch.is_grad_enabled():
                raise RuntimeError("b200 fused step is inference-only: call under torch.no_grad() after .eval()")
            st_in = getattr(args, "stepper_state", None)
            cs = getattr(st_in, "corrector_state", None)
            gm = getattr(cs, "global_dry_air_mass", None)
            carried = {"corrector_state": {"global_dry_air_mass": gm}} if gm is not None else None
            # a random_state, when present, drives the noise of a NoiseConditionedModel the way fme.core.rand.use_generator does
            out = self._fused.step(args.input, args.next_step_input_data, stepper_state=carried)
            B = next(iter(out.values())).shape[0]
            new = self._fused.get_stepper_state(B)
            state_out = st_in
            if new is not None and StepperState is not None and CorrectorState is not None:
                dev = next(iter(out.values())).device
                cstate = CorrectorState(global_dry_air_mass=new["corrector_state"]["global_dry_air_mass"].to(dev))
                state_out = dataclasses.replace(st_in, corrector_state=cstate) if st_in is not None else StepperState(corrector_state=cstate)
            return StepOutput(output=dict(out), stepper_state=state_out)

    return B200FusedStep


def install_step_into_fme(override: bool = False, name: str = STEP_TYPE_NAME):
    """Register the fused step config with the real ``fme`` ``StepSelector`` (needs ``fme`` importable).  Returns the config class."""
    from fme.core.step.output import StepOutput  # noqa: PLC0415
    from fme.core.step.single_module import SingleModuleStepConfig  # noqa: PLC0415
    from fme.core.step.step import StepABC, StepSelector  # noqa: PLC0415

    from .registry import disabled_by_env  # noqa: PLC0415

    if disabled_by_env():  # ACE_B200_DISABLE=1: A/B kill switch -- `name` becomes an alias of the reference's single-module step
        StepSelector.register(name)(SingleModuleStepConfig)
        return SingleModuleStepConfig

    try:
        from fme.core.corrector.state import CorrectorState  # noqa: PLC0415
        from fme.core.stepper_state import StepperState  # noqa: PLC0415
    except ImportError:  # older fme without per-sample stepper state
        CorrectorState = StepperState = None
    step_cls = make_fused_step_class(StepABC, StepOutput, StepperState, CorrectorState)

    @dataclasses.dataclass
    class B200SingleModuleStepConfig(SingleModuleStepConfig):
        """``SingleModuleStepConfig`` (fme/core/step/single_module.py:48-259) whose step runs as one fused library call."""

        def get_step(self, dataset_info, init_weights: Callable[[list], None] = lambda x: None):
            inner = SingleModuleStepConfig.get_step(self, dataset_info, init_weights)
            return step_cls(inner, dataset_info)

    B200SingleModuleStepConfig.fused_step_class = step_cls
    StepSelector.register(name)(B200SingleModuleStepConfig)
    if override:
        for n in REFERENCE_STEP_TYPE_NAMES:
            StepSelector.register(n)(B200SingleModuleStepConfig)
    return B200SingleModuleStepConfig
