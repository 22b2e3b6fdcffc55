"""ace_b200: B200-native (sm_100a) implementation of ACE's SFNO inference hot path.

Public surface (mirrors the reference interfaces it replaces):
  RealSHT, InverseRealSHT                      <- fme.sht_fix.RealSHT / InverseRealSHT
  SphericalFourierNeuralOperatorNet            <- fme.ace.models.modulus.sfnonet.SphericalFourierNeuralOperatorNet
  B200SphericalFourierNeuralOperatorBuilder,
  ModuleSelector, DatasetInfo                  <- fme.ace.registry.sfno / fme.core.registry.module
  FusedStepper                                 <- device work of fme.core.step.single_module.step_with_adjustments
  install_step_into_fme                        <- a StepSelector-registered SingleModuleStepConfig whose step is the fused library call
  metrics.LatLonOperations, spherical_power_spectrum <- fme.core.gridded_ops.LatLonOperations / fme.core.metrics (device reductions)
  HealpixSHT, HealpixISHT                      <- fme.core.cuhpx.sht.SHT / iSHT (HEALPix ring-order transform)
  parallel                                     <- data-parallel surface of fme.core.distributed (ensemble sharding, one gather)
"""
from ._lib import AceError, get_option, launch_count, set_option  # noqa: F401
from .registry import (  # noqa: F401
    B200SphericalFourierNeuralOperatorBuilder,
    DatasetInfo,
    Module,
    ModuleSelector,
    install_into_fme,
)
from .sfno import SphericalFourierNeuralOperatorNet  # noqa: F401
from .sht import InverseRealSHT, RealSHT, patch_torch_harmonics  # noqa: F401
from .stepper import FusedStepper  # noqa: F401
from .fme_step import install_step_into_fme  # noqa: F401
from .corrector import AtmosphereCorrector  # noqa: F401
from . import parallel  # noqa: F401
from . import metrics  # noqa: F401
from . import timing  # noqa: F401
from .healpix import HealpixISHT, HealpixSHT  # noqa: F401

__version__ = "0.1.0"
from . import csfno  # noqa: F401
from .csfno import NoiseConditionedModel, NoiseConditionedSFNO  # noqa: F401
