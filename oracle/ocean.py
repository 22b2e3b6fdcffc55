"""Oracle restatement of the slab ocean and the ocean prescriber (torch CPU).  TEST INFRASTRUCTURE.

Follows /root/reference:
  fme/core/ocean.py:64-88       SlabOceanSurfaceTemperature.__call__
  fme/core/ocean.py:223-243     mixed_layer_temperature_tendency
  fme/core/metrics.py:299-334   net_surface_energy_flux (without frozen precipitation)
  fme/core/prescriber.py:94-108 + fme/core/spatial_masking.py:25-30   replace where round(mask) == 1, or the linear blend
Pinned against the reference's own functions in tests/test_oracle_corrector.py (build container).
"""
import torch

DENSITY_OF_WATER = 1000.0        # fme/core/constants.py
SPECIFIC_HEAT_OF_WATER = 4000.0


def net_surface_energy_flux_without_frozen_precip(dlw, ulw, dsw, usw, lhf, shf):
    return (dsw - usw + dlw - ulw) + (-lhf - shf) - 0.0


def mixed_layer_temperature_tendency(f_net, q_flux, depth):
    return (f_net + q_flux) / (DENSITY_OF_WATER * depth * SPECIFIC_HEAT_OF_WATER)


def slab_surface_temperature(t_in, f_net, q_flux, depth, timestep_seconds):
    return t_in + mixed_layer_temperature_tendency(f_net, q_flux, depth) * timestep_seconds


def prescribe(mask, gen, target, interpolate):
    if interpolate:
        return mask * target + (1 - mask) * gen
    return torch.where(torch.round(mask).to(int) == 1, target, gen)
