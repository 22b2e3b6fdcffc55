"""Oracle restatement of the ACE2 conservation correctors (torch CPU).  TEST INFRASTRUCTURE.

Follows /root/reference:
  fme/core/coordinates.py:201-205,241-284   HybridSigmaPressureCoordinate.get_ak/get_bk/interface_pressure/vertical_integral
  fme/core/atmosphere_data.py:180-197       surface_pressure_due_to_dry_air = ps - g * total_water_path
  fme/core/atmosphere_data.py:270-279       evaporation_rate = latent_heat_flux / Lv (and the inverse setter)
  fme/core/corrector/atmosphere.py:404-427  _seed_global_dry_air_mass
  fme/core/corrector/atmosphere.py:430-463  _adjust_gen_dry_air_to_target
  fme/core/corrector/atmosphere.py:518-608  _force_conserve_moisture
Fields are tensors [B, H, W]; water is [B, H, W, nz].  Pinned against the reference's own functions (extracted from the tree
by oracle/refload.py:load_corrector) in tests/test_oracle_corrector.py.
"""
import torch

GRAVITY = 9.80665  # fme/core/constants.py
LATENT_HEAT_OF_VAPORIZATION = 2.5e6


class VerticalCoordinate:
    def __init__(self, ak, bk):
        self.ak, self.bk = torch.as_tensor(ak), torch.as_tensor(bk)

    def get_ak(self):
        return self.ak

    def get_bk(self):
        return self.bk

    def interface_pressure(self, surface_pressure):
        return torch.stack([ak + bk * surface_pressure for ak, bk in zip(self.ak, self.bk)], dim=-1)

    def vertical_integral(self, integrand, surface_pressure):
        return (integrand * self.interface_pressure(surface_pressure).diff(dim=-1)).sum(dim=-1) / GRAVITY


def dry_air(ps, wat, vc):
    return ps - GRAVITY * vc.vertical_integral(wat, ps)


def seed_global_dry_air_mass(ps_in, wat_in, area_weighted_mean, vc, precision=torch.float64):
    return area_weighted_mean(dry_air(ps_in, wat_in, vc).to(precision), keepdim=True)


def adjust_dry_air_to_target(ps, wat, target, area_weighted_mean, vc, precision=torch.float64):
    gen_dry = dry_air(ps, wat, vc)
    error = area_weighted_mean(gen_dry.to(precision), keepdim=True) - target.to(precision)
    new_dry = gen_dry.to(precision) - error
    w = wat.to(precision)
    akd, bkd = vc.get_ak().diff().to(precision), vc.get_bk().diff().to(precision)
    return ((new_dry + (akd * w).sum(-1)) / (1 - (bkd * w).sum(-1))).to(ps.dtype)


def conserve_moisture(ps_in, wat_in, ps, wat, precip, lhf, area_weighted_mean, vc, timestep_seconds, terms_to_modify):
    """Returns (precipitation, latent_heat_flux, advective tendency or None)."""
    tend = (vc.vertical_integral(wat, ps) - vc.vertical_integral(wat_in, ps_in)) / timestep_seconds
    m_t = area_weighted_mean(tend, keepdim=True)
    evap = lhf / LATENT_HEAT_OF_VAPORIZATION
    m_e = area_weighted_mean(evap, keepdim=True)
    m_p = area_weighted_mean(precip, keepdim=True)
    if terms_to_modify.endswith("precipitation"):
        precip = precip * ((m_e - m_t) / m_p)
    elif terms_to_modify.endswith("evaporation"):
        lhf = evap * ((m_t + m_p) / m_e) * LATENT_HEAT_OF_VAPORIZATION  # set_evaporation_rate stores the flux ...
        evap = lhf / LATENT_HEAT_OF_VAPORIZATION  # ... and evaporation_rate is re-derived from it (atmosphere_data.py:270-279)
    adv = tend - (evap - precip) if terms_to_modify.startswith("advection") else None
    return precip, lhf, adv
