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


# ---- the remaining corrections of the sequence (atmosphere.py:349-398 builds it in this order: ForcePositive, dry air,
# zero-mean moisture advection, moisture budget (+ frozen-precipitation clip), total energy budget) -----------------------------
LATENT_HEAT_OF_FREEZING = 334000.0
SPECIFIC_HEAT_OF_DRY_AIR_CONST_PRESSURE = 1004.6
RVGAS = 461.5
RDGAS = 287.05
SPECIFIC_HEAT_OF_DRY_AIR_CONST_VOLUME = SPECIFIC_HEAT_OF_DRY_AIR_CONST_PRESSURE - RDGAS


def zero_global_mean_moisture_advection(adv, area_weighted_mean):
    """atmosphere.py:467-490: subtract the global mean from the advective tendency."""
    return adv - area_weighted_mean(adv)[..., None, None]


def clip_frozen_precipitation(frozen, precip):
    """atmosphere.py:493-515."""
    return torch.minimum(frozen, precip)


def layer_thickness(p_int, temp, wat):
    """atmosphere_data.py:376-393 (hydrostatic, virtual temperature from total water, TOA pressure clamped to 1 Pa)."""
    tv = temp * (1 + (RVGAS / RDGAS - 1.0) * wat)
    dlogp = torch.log(torch.clamp(p_int, min=1.0)).diff(dim=-1)
    return dlogp * RDGAS * tv / GRAVITY


def _rev_cumsum(x):
    return torch.cumsum(x.flip(dims=(-1,)), dim=-1).flip(dims=(-1,))


def total_energy_ace2_path(ps, temp, wat, hgt, vc):
    """atmosphere_data.py:310-326,341-365,396-418: cv T + Lv q + g z_mid, integrated over the column."""
    thick = layer_thickness(vc.interface_pressure(ps), temp, wat)
    hsfc = torch.where(hgt < 0.0, 0, hgt).reshape(*hgt.shape, 1)
    cum = _rev_cumsum(thick)
    h_int = torch.concat([cum + hsfc.broadcast_to(cum.shape), hsfc], dim=-1)
    h_mid = 0.5 * (h_int[..., :-1] + h_int[..., 1:])
    te = temp * SPECIFIC_HEAT_OF_DRY_AIR_CONST_VOLUME + wat * LATENT_HEAT_OF_VAPORIZATION + h_mid * GRAVITY
    return vc.vertical_integral(te, ps)


def net_energy_flux_into_atmosphere(f):
    """atmosphere_data.py:228-250 + metrics.py:299-352; ``f``: dict of the nine flux fields (frozen precipitation may be None)."""
    frozen = f["frozen"] * LATENT_HEAT_OF_FREEZING if f.get("frozen") is not None else 0.0
    rad = f["dsw_sfc"] - f["usw_sfc"] + f["dlw_sfc"] - f["ulw_sfc"]
    turb = -f["lhf"] - f["shf"]
    sfc = rad + turb - frozen
    toa = f["dsw_toa"] - f["usw_toa"] - f["ulw_toa"]
    return toa - sfc


def energy_correction_factor(ps, temp, wat, vc):
    """atmosphere.py:668-695."""
    q_times_dlogp = layer_thickness(vc.interface_pressure(ps), temp, wat) * GRAVITY / temp
    integrand = SPECIFIC_HEAT_OF_DRY_AIR_CONST_VOLUME - 0.5 * q_times_dlogp + _rev_cumsum(q_times_dlogp)
    return vc.vertical_integral(integrand, ps)


def conserve_total_energy(ps_in, temp_in, wat_in, hgt_in, ps, temp, wat, hgt_next, fluxes, area_weighted_mean, vc, timestep_seconds,
                          unaccounted_heating=0.0):
    """atmosphere.py:611-665 (method constant_temperature).  Returns the corrected air temperature [..., nz]."""
    e_gen = area_weighted_mean(total_energy_ace2_path(ps, temp, wat, hgt_next, vc), keepdim=True)
    e_in = area_weighted_mean(total_energy_ace2_path(ps_in, temp_in, wat_in, hgt_in, vc), keepdim=True)
    flux = area_weighted_mean(net_energy_flux_into_atmosphere(fluxes), keepdim=True)
    desired = e_in + (flux + unaccounted_heating) * timestep_seconds
    factor = area_weighted_mean(energy_correction_factor(ps, temp, wat, vc), True)
    return temp + ((desired - e_gen) / factor)[..., None]
