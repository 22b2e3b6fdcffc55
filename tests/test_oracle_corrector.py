"""CPU, build container only: oracle/corrector.py against the reference's own corrector functions."""
import pytest
import torch

from oracle import corrector as oc
from oracle import metrics as om
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present (GPU box)")

NZ = 8


def synthetic_state(seed, B=2, H=12, W=24):
    g = torch.Generator().manual_seed(seed)
    ak = torch.linspace(0.0, 5000.0, NZ + 1).flip(0) * torch.linspace(1.0, 0.0, NZ + 1)  # decreasing to 0 at the surface
    bk = torch.linspace(0.0, 1.0, NZ + 1)
    lat = torch.linspace(-80, 80, H)
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(H, W).contiguous()

    def fields(ps_mean):
        d = {"PRESsfc": ps_mean + 500.0 * torch.randn(B, H, W, generator=g)}
        for k in range(NZ):
            d[f"specific_total_water_{k}"] = 1e-3 * (k + 1) * torch.rand(B, H, W, generator=g)
        return d

    inp = fields(1.0e5)
    gen = fields(1.0005e5)
    gen["PRATEsfc"] = 3e-5 * torch.rand(B, H, W, generator=g)
    gen["LHTFLsfc"] = 80.0 + 40.0 * torch.rand(B, H, W, generator=g)
    gen["tendency_of_total_water_path_due_to_advection"] = 1e-5 * torch.randn(B, H, W, generator=g)
    return ak, bk, w, inp, gen


def _wat(d):
    return torch.stack([d[f"specific_total_water_{k}"] for k in range(NZ)], dim=-1)


@pytest.mark.parametrize("mode", ["precipitation", "advection_and_precipitation", "evaporation", "advection_and_evaporation"])
def test_oracle_corrector_equals_reference(mode):
    ref = refload.load_corrector()
    ak, bk, w, inp, gen = synthetic_state(0)
    vc = oc.VerticalCoordinate(ak, bk)

    def awm(data, keepdim=False, name=None):
        return om.weighted_mean(data, w.to(data.dtype), keepdim=keepdim)

    # dry air (atmosphere.py:404-463)
    st = ref.seed(inp, None, awm, vc, torch.float64)
    target = oc.seed_global_dry_air_mass(inp["PRESsfc"], _wat(inp), awm, vc)
    assert torch.equal(st.global_dry_air_mass, target)
    r = ref.adjust(dict(gen), target, awm, vc, torch.float64)
    new_ps = oc.adjust_dry_air_to_target(gen["PRESsfc"], _wat(gen), target, awm, vc)
    assert set(r) == {"PRESsfc"} and torch.equal(r["PRESsfc"], new_ps)
    # the pin holds: global dry air of the corrected state equals the target (to fp32 resolution of ps)
    achieved = awm(oc.dry_air(new_ps, _wat(gen), vc).double(), keepdim=True)
    assert float((achieved - target).abs().max()) < 0.02  # Pa
    # moisture budget (atmosphere.py:518-608) on the pressure-corrected state, as the corrector sequence applies it
    gen2 = {**gen, "PRESsfc": new_ps}
    r = ref.conserve_moisture(inp, dict(gen2), awm, vc, 21600.0, mode)
    precip, lhf, adv = oc.conserve_moisture(inp["PRESsfc"], _wat(inp), new_ps, _wat(gen2), gen2["PRATEsfc"], gen2["LHTFLsfc"], awm, vc,
                                            21600.0, mode)
    if mode.endswith("precipitation"):
        assert torch.equal(r["PRATEsfc"], precip) and "LHTFLsfc" not in r
    else:
        assert torch.equal(r["LHTFLsfc"], lhf) and "PRATEsfc" not in r
    if mode.startswith("advection"):
        assert torch.equal(r["tendency_of_total_water_path_due_to_advection"], adv)
    else:
        assert adv is None and "tendency_of_total_water_path_due_to_advection" not in r


def _temp(d):
    return torch.stack([d[f"air_temperature_{k}"] for k in range(NZ)], dim=-1)


def energy_state(seed, B=2, H=12, W=24):
    ak, bk, w, inp, gen = synthetic_state(seed, B, H, W)
    g = torch.Generator().manual_seed(seed + 100)
    for d in (inp, gen):
        for k in range(NZ):
            d[f"air_temperature_{k}"] = 210.0 + 10.0 * k + 3.0 * torch.randn(B, H, W, generator=g)
    inp["HGTsfc"] = 800.0 * torch.randn(B, H, W, generator=g)  # negative heights are clamped to 0 by the reference
    forcing = {"HGTsfc": inp["HGTsfc"], "DSWRFtoa": 340.0 + 100.0 * torch.rand(B, H, W, generator=g)}
    for n, m in [("DLWRFsfc", 340.0), ("ULWRFsfc", 390.0), ("DSWRFsfc", 190.0), ("USWRFsfc", 30.0), ("SHTFLsfc", 20.0),
                 ("USWRFtoa", 100.0), ("ULWRFtoa", 240.0)]:
        gen[n] = m + 20.0 * torch.randn(B, H, W, generator=g)
    return ak, bk, w, inp, gen, forcing


@pytest.mark.parametrize("frozen", [False, True])
def test_oracle_energy_and_small_corrections_equal_reference(frozen):
    ref = refload.load_corrector()
    ak, bk, w, inp, gen, forcing = energy_state(3)
    vc = oc.VerticalCoordinate(ak, bk)
    if frozen:
        gen["total_frozen_precipitation_rate"] = gen["PRATEsfc"] * 2.0 * torch.rand(gen["PRATEsfc"].shape, generator=torch.Generator().manual_seed(8))

    def awm(data, keepdim=False, name=None):
        return om.weighted_mean(data, w.to(data.dtype), keepdim=keepdim)

    # zero-mean advection (atmosphere.py:467-490) and frozen-precipitation clip (:493-515)
    r = ref.zero_mean_advection(dict(gen), awm)
    name = "tendency_of_total_water_path_due_to_advection"
    assert set(r) == {name} and torch.equal(r[name], oc.zero_global_mean_moisture_advection(gen[name], awm))
    r = ref.clip_frozen(dict(gen))
    if frozen:
        assert torch.equal(r["total_frozen_precipitation_rate"], oc.clip_frozen_precipitation(gen["total_frozen_precipitation_rate"], gen["PRATEsfc"]))
        assert not torch.equal(r["total_frozen_precipitation_rate"], gen["total_frozen_precipitation_rate"])
    else:
        assert r == {}
    # total energy budget (atmosphere.py:611-695)
    for heating in (0.0, 1.14):
        r = ref.conserve_energy(inp, dict(gen), forcing, awm, vc, 21600.0, "constant_temperature", heating)
        fl = dict(dsw_toa=forcing["DSWRFtoa"], usw_toa=gen["USWRFtoa"], ulw_toa=gen["ULWRFtoa"], dlw_sfc=gen["DLWRFsfc"], ulw_sfc=gen["ULWRFsfc"],
                  dsw_sfc=gen["DSWRFsfc"], usw_sfc=gen["USWRFsfc"], lhf=gen["LHTFLsfc"], shf=gen["SHTFLsfc"],
                  frozen=gen.get("total_frozen_precipitation_rate"))
        t = oc.conserve_total_energy(inp["PRESsfc"], _temp(inp), _wat(inp), inp["HGTsfc"], gen["PRESsfc"], _temp(gen), _wat(gen),
                                     forcing["HGTsfc"], fl, awm, vc, 21600.0, heating)
        assert set(r) == {f"air_temperature_{k}" for k in range(NZ)}
        for k in range(NZ):
            assert torch.equal(r[f"air_temperature_{k}"], t[..., k]), k
        assert float((t - _temp(gen)).abs().max()) > 1e-3  # the correction is not a no-op on this state


def test_oracle_slab_ocean_equals_reference():
    from oracle import ocean as oo

    ref = refload.load_corrector()
    ak, bk, w, inp, gen, forcing = energy_state(5)
    g = torch.Generator().manual_seed(1)
    q = 30.0 * torch.randn(gen["PRESsfc"].shape, generator=g)
    depth = 20.0 + 60.0 * torch.rand(gen["PRESsfc"].shape, generator=g)
    f_ref = ref.AtmosphereData(gen).net_surface_energy_flux_without_frozen_precip
    f = oo.net_surface_energy_flux_without_frozen_precip(gen["DLWRFsfc"], gen["ULWRFsfc"], gen["DSWRFsfc"], gen["USWRFsfc"], gen["LHTFLsfc"],
                                                         gen["SHTFLsfc"])
    assert torch.equal(f, f_ref)
    assert torch.equal(oo.mixed_layer_temperature_tendency(f, q, depth), ref.mixed_layer_temperature_tendency(f_ref, q, depth))
