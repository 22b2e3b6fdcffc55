"""GPU: conservation correctors (dry-air pin, moisture budget) vs oracle/corrector.py, alone and inside the fused step."""
import pytest
import torch

from tests.util import field_rel_err

pytestmark = pytest.mark.gpu

NZ = 8
MODES = [None, "precipitation", "advection_and_precipitation", "evaporation", "advection_and_evaporation"]
PROG = ["PRESsfc"] + [f"specific_total_water_{k}" for k in range(NZ)]
# deliberately shuffled: output order != prognostic order, water levels not contiguous
OUT = ["specific_total_water_3", "PRATEsfc", "PRESsfc", "specific_total_water_0", "LHTFLsfc", "specific_total_water_7",
       "tendency_of_total_water_path_due_to_advection", "specific_total_water_1", "specific_total_water_2", "specific_total_water_4",
       "specific_total_water_5", "specific_total_water_6"]
PROG_ORDER = ["specific_total_water_2", "PRESsfc"] + [f"specific_total_water_{k}" for k in (0, 1, 3, 4, 5, 6, 7)]


def _coords(H, W):
    ak = torch.linspace(0.0, 5000.0, NZ + 1).flip(0) * torch.linspace(1.0, 0.0, NZ + 1)
    bk = torch.linspace(0.0, 1.0, NZ + 1)
    lat = torch.linspace(-80, 80, H)
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(H, W).contiguous()
    return ak, bk, w


def _fields(g, B, H, W, ps_mean, diag):
    d = {"PRESsfc": ps_mean + 500.0 * torch.randn(B, H, W, generator=g)}
    for k in range(NZ):
        d[f"specific_total_water_{k}"] = 1e-3 * (k + 1) * torch.rand(B, H, W, generator=g)
    if diag:
        d["PRATEsfc"] = 3e-5 * torch.rand(B, H, W, generator=g)
        d["LHTFLsfc"] = 80.0 + 40.0 * torch.rand(B, H, W, generator=g)
        d["tendency_of_total_water_path_due_to_advection"] = 1e-5 * torch.randn(B, H, W, generator=g)
    return d


def _wat(d):
    return torch.stack([d[f"specific_total_water_{k}"] for k in range(NZ)], dim=-1)


def _oracle_correct(inp, gen, target, w, ak, bk, dry, mode):
    """AtmosphereCorrector.__call__ order (fme/core/corrector/atmosphere.py:349-398): dry air, then moisture."""
    from oracle import corrector as oc
    from oracle import metrics as om

    vc = oc.VerticalCoordinate(ak, bk)

    def awm(data, keepdim=False, name=None):
        return om.weighted_mean(data, w.to(data.dtype), keepdim=keepdim)

    gen = dict(gen)
    if target is None:
        target = oc.seed_global_dry_air_mass(inp["PRESsfc"], _wat(inp), awm, vc)
    if dry:
        gen["PRESsfc"] = oc.adjust_dry_air_to_target(gen["PRESsfc"], _wat(gen), target, awm, vc)
    if mode is not None:
        p, l, a = oc.conserve_moisture(inp["PRESsfc"], _wat(inp), gen["PRESsfc"], _wat(gen), gen["PRATEsfc"], gen["LHTFLsfc"], awm, vc,
                                       21600.0, mode)
        gen["PRATEsfc"], gen["LHTFLsfc"] = p, l
        if a is not None:
            gen["tendency_of_total_water_path_due_to_advection"] = a
    return gen, target


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dry", [True, False])
def test_corrector_apply_matches_oracle(mode, dry):
    from ace_b200.corrector import AtmosphereCorrector

    B, H, W = 3, 24, 48
    g = torch.Generator().manual_seed(11)
    ak, bk, w = _coords(H, W)
    ic = _fields(g, B, H, W, 1.0e5, False)       # initial condition (dry-air reference)
    inp = _fields(g, B, H, W, 1.0002e5, False)   # the step's input state (2nd step of a rollout: differs from the IC)
    gen = _fields(g, B, H, W, 1.0007e5, True)
    from oracle import corrector as oc
    from oracle import metrics as om

    vc = oc.VerticalCoordinate(ak, bk)
    target = oc.seed_global_dry_air_mass(ic["PRESsfc"], _wat(ic), lambda d, keepdim=False: om.weighted_mean(d, w.to(d.dtype), keepdim=keepdim), vc)
    ref, _ = _oracle_correct(inp, gen, target, w, ak, bk, dry, mode)
    dbl = lambda d: {k: v.double() for k, v in d.items()}  # noqa: E731
    ref64, _ = _oracle_correct(dbl(inp), dbl(gen), target, w.double(), ak.double(), bk.double(), dry, mode)  # same algorithm, fp64

    c = AtmosphereCorrector(OUT, PROG_ORDER, (H, W), ak, bk, w, 21600.0, conserve_dry_air=dry, moisture_budget_correction=mode)
    pack = lambda d, names: torch.stack([d[n] for n in names], dim=1).contiguous().cuda()  # noqa: E731
    out, prev, nxt = pack(gen, OUT), pack(inp, PROG_ORDER), pack(gen, PROG_ORDER)
    with pytest.raises(Exception):
        c.apply(prev, out.clone(), nxt.clone())      # not seeded
    c.seed(pack(ic, PROG_ORDER))
    assert c.seeded
    c.apply(prev, out, nxt)
    for i, n in enumerate(OUT):
        got, want = out[:, i].cpu(), ref[n]
        if torch.equal(want, gen[n]):
            assert torch.equal(got, want), f"{n}: field the reference leaves alone was modified"
        else:
            # The reference evaluates the column integrals and the budget's global means in fp32; the budget ratio
            # (<E> - <dTWP/dt>) / <P> amplifies that rounding (dTWP/dt is a difference of two ~20 kg/m2 columns) to ~1e-5
            # relative.  The device accumulates columns and means in fp64, so it must (1) agree with the fp32 reference to
            # that conditioning noise and (2) sit within fp32 output rounding of the same algorithm evaluated in fp64.
            e32 = field_rel_err(got[:, None], want[:, None])
            e64 = field_rel_err(got[:, None].double(), ref64[n][:, None])
            r64 = field_rel_err(want[:, None].double(), ref64[n][:, None])
            assert e32 < 3e-5, (n, mode, dry, e32)
            assert e64 < 2e-6 and e64 <= r64 + 2e-7, (n, mode, dry, e64, r64)
    for i, n in enumerate(PROG_ORDER):  # corrected prognostic fields are fed back identically
        assert torch.equal(nxt[:, i], out[:, OUT.index(n)]), n
    if dry:  # the pin holds on the device result: global dry-air mean == target to fp32 resolution of ps
        o = {n: out[:, i].cpu() for i, n in enumerate(OUT)}
        achieved = om.weighted_mean(oc.dry_air(o["PRESsfc"], _wat(o), vc).double(), w.double(), keepdim=True)
        assert float((achieved - target).abs().max()) < 0.02  # Pa
    if mode is not None:  # and the budget closes: <dTWP/dt> = <E> - <P> (+ <adv> = 0 when advection is the residual)
        o = {n: out[:, i].cpu().double() for i, n in enumerate(OUT)}
        twp1 = oc.VerticalCoordinate(ak.double(), bk.double()).vertical_integral(_wat(o), o["PRESsfc"])
        twp0 = oc.VerticalCoordinate(ak.double(), bk.double()).vertical_integral(_wat(inp).double(), inp["PRESsfc"].double())
        mean = lambda d: om.weighted_mean(d, w.double())  # noqa: E731
        resid = mean((twp1 - twp0) / 21600.0) - mean(o["LHTFLsfc"] / 2.5e6) + mean(o["PRATEsfc"])
        assert float(resid.abs().max()) < 1e-9, float(resid.abs().max())


def test_bad_options_raise():
    from ace_b200.corrector import AtmosphereCorrector

    ak, bk, w = _coords(8, 16)
    with pytest.raises(ValueError):  # no air temperature / flux fields among these outputs
        AtmosphereCorrector(OUT, PROG_ORDER, (8, 16), ak, bk, w, total_energy_budget_correction={"method": "constant_temperature"})
    with pytest.raises(NotImplementedError):
        AtmosphereCorrector(OUT, PROG_ORDER, (8, 16), ak, bk, w, total_energy_budget_correction={"method": "something_else"})
    with pytest.raises(ValueError):
        AtmosphereCorrector(OUT, PROG_ORDER, (8, 16), ak[:-1], bk[:-1], w)


def _stepper(mode, ocean):
    import ace_b200
    from oracle import sfno as osfno

    img = (32, 64)
    in_names = PROG_ORDER + ["forcing_a", "ocean_fraction"]
    out_names = OUT + ["surface_temperature_diag"]
    allnames = sorted(set(in_names + out_names))
    means = {n: 0.0 for n in allnames}
    stds = {n: 1.0 for n in allnames}
    means["PRESsfc"], stds["PRESsfc"] = 1.0e5, 800.0
    for k in range(NZ):
        means[f"specific_total_water_{k}"], stds[f"specific_total_water_{k}"] = 1e-3 * (k + 1), 4e-4 * (k + 1)
    means["PRATEsfc"], stds["PRATEsfc"] = 3e-5, 1.5e-5
    means["LHTFLsfc"], stds["LHTFLsfc"] = 90.0, 30.0
    stds["tendency_of_total_water_path_due_to_advection"] = 1e-5
    means["surface_temperature_diag"], stds["surface_temperature_diag"] = 285.0, 10.0
    torch.manual_seed(0)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, len(in_names), len(out_names), embed_dim=16, num_layers=2, operator_type="dhconv").eval()
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=16, num_layers=2, operator_type="dhconv"))
    net = sel.build(len(in_names), len(out_names), ace_b200.DatasetInfo(img_shape=img)).torch_module
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval().requires_grad_(False)
    ak, bk, w = _coords(*img)
    fp = [f"specific_total_water_{k}" for k in range(NZ)] + ["PRATEsfc"]
    st = ace_b200.FusedStepper(
        net, in_names, out_names, means, stds, residual_prediction=True, force_positive_names=fp,
        ocean=dict(surface_temperature_name="surface_temperature_diag", ocean_fraction_name="ocean_fraction", interpolate=False) if ocean else None,
        corrector=dict(conserve_dry_air=True, moisture_budget_correction=mode, ak=ak, bk=bk, area_weights=w, timestep_seconds=21600.0))
    return img, in_names, out_names, means, stds, onet, st, (ak, bk, w), fp


@pytest.mark.parametrize("mode", ["advection_and_precipitation", "evaporation"])
def test_fused_step_with_corrector_matches_oracle_sequence(mode):
    """Two chained steps through the reference's sequence (fme/core/step/single_module.py:648-709): network, denormalise,
    ForcePositive, dry air + moisture corrector (reference captured at the first step's input), ocean prescriber."""
    from tests.test_gpu_stepper import _oracle_step

    img, in_names, out_names, means, stds, onet, st, (ak, bk, w), fp = _stepper(mode, ocean=True)
    B = 2
    g = torch.Generator().manual_seed(5)
    state = {n: torch.randn(B, *img, generator=g) * stds[n] + means[n] for n in in_names}
    for k in range(NZ):
        state[f"specific_total_water_{k}"] = state[f"specific_total_water_{k}"].clamp(min=0)
    target = None
    dev_state = {n: v.cuda() for n, v in state.items()}
    st.reset_corrector_state()
    for step in range(2):
        mask = torch.rand(B, *img, generator=g)
        sst = 280.0 + torch.randn(B, *img, generator=g)
        ref = _oracle_step(onet, in_names, out_names, means, stds, True, state)
        for n in fp:
            ref[n] = torch.clamp(ref[n], min=0.0)
        ref, target = _oracle_correct(state, ref, target, w, ak, bk, True, mode)
        ref["surface_temperature_diag"] = torch.where(torch.round(mask).to(int) == 1, sst, ref["surface_temperature_diag"])
        out = st.step(dev_state, next_step_input_data={"ocean_fraction": mask.cuda(), "surface_temperature_diag": sst.cuda()})
        for n in out_names:
            a = ((out[n].cpu() - means[n]) / stds[n])[:, None]
            b = ((ref[n] - means[n]) / stds[n])[:, None]
            # the advective tendency is (TWP_out - TWP_in) / dt - (E - P): two ~20 kg/m2 columns that differ by ~1e-3 of
            # themselves here, so the 1e-4 agreement of the network outputs is a 1e-3 agreement of this residual
            tol = 1e-3 if n == "tendency_of_total_water_path_due_to_advection" else 1e-4
            assert field_rel_err(a, b) < tol, (n, step)
        # both sides advance on their own outputs (errors would compound, not hide)
        new_forcing = {n: torch.randn(B, *img, generator=g) for n in ("forcing_a", "ocean_fraction")}
        state = {**{n: ref[n] for n in PROG_ORDER}, **new_forcing}
        dev_state = {**{n: out[n] for n in PROG_ORDER}, **{n: v.cuda() for n, v in new_forcing.items()}}


def test_rollout_with_corrector_graph_equals_eager_and_reseeds():
    img, in_names, out_names, means, stds, onet, st, _, _ = _stepper("advection_and_precipitation", ocean=False)
    B, T = 2, 4
    g = torch.Generator().manual_seed(9)
    n_prog, n_f = len(st.prognostic_names), len(st.forcing_names)
    pm = torch.tensor([means[n] for n in st.prognostic_names])[None, :, None, None]
    ps = torch.tensor([stds[n] for n in st.prognostic_names])[None, :, None, None]
    prog0 = (torch.randn(B, n_prog, *img, generator=g) * ps + pm).cuda()
    prog1 = (torch.randn(B, n_prog, *img, generator=g) * ps + pm + 0.3 * ps).cuda()
    forcing = torch.randn(T, B, n_f, *img, generator=g).cuda()
    oe, fe = st.rollout(prog0, forcing, T, use_cuda_graph=False)
    og, fg = st.rollout(prog0, forcing, T, use_cuda_graph=True)
    torch.testing.assert_close(og, oe, rtol=0, atol=0)
    torch.testing.assert_close(fg, fe, rtol=0, atol=0)
    # a second rollout from another initial condition re-captures the dry-air reference (graph replays read it from the device)
    oe1, _ = st.rollout(prog1, forcing, T, use_cuda_graph=False)
    og1, _ = st.rollout(prog1, forcing, T, use_cuda_graph=True)
    torch.testing.assert_close(og1, oe1, rtol=0, atol=0)
    assert not torch.equal(og1, og)
    out_host = torch.empty(T, B, len(out_names), *img).pin_memory()
    st.rollout_host(prog0, forcing.cpu().pin_memory(), T, out_host)
    torch.cuda.synchronize()
    torch.testing.assert_close(out_host, og.cpu(), rtol=0, atol=0)
    # the invariant itself: the global dry-air mean of every step's output equals the initial condition's
    from oracle import corrector as oc
    from oracle import metrics as om

    ak, bk, w = _coords(*img)
    vc = oc.VerticalCoordinate(ak, bk)

    def dry_mean(d):
        return om.weighted_mean(oc.dry_air(d["PRESsfc"], _wat(d), vc).double(), w.double())

    ic = {n: prog0[:, i].cpu() for i, n in enumerate(st.prognostic_names)}
    for t in range(T):
        o = {n: og[t, :, i].cpu() for i, n in enumerate(out_names)}
        assert float((dry_mean(o) - dry_mean(ic)).abs().max()) < 0.05, t


def test_corrector_state_is_carried_across_windows_and_step_seeds_from_its_own_input():
    """(a) two chained windows with the carried ``stepper_state`` equal ONE rollout of both windows (the dry-air
    target stays pinned to the first initial condition, fme/ace/stepper/single_module.py:1160-1165), while a window started
    without the state re-pins it; (b) ``step()`` without a state seeds from each call's own input (fme/core/corrector/atmosphere.py:404-427)."""
    img, in_names, out_names, means, stds, onet, st, _, _ = _stepper("advection_and_precipitation", ocean=False)
    B, T = 2, 4
    g = torch.Generator().manual_seed(19)
    n_prog, n_f = len(st.prognostic_names), len(st.forcing_names)
    pm = torch.tensor([means[n] for n in st.prognostic_names])[None, :, None, None]
    ps = torch.tensor([stds[n] for n in st.prognostic_names])[None, :, None, None]
    prog0 = (torch.randn(B, n_prog, *img, generator=g) * ps + pm).cuda()
    forcing = torch.randn(2 * T, B, n_f, *img, generator=g).cuda()
    for graph in (False, True):
        o_all, f_all = st.rollout(prog0, forcing, 2 * T, use_cuda_graph=graph)
        o1, f1 = st.rollout(prog0, forcing[:T], T, use_cuda_graph=graph)
        state = st.get_stepper_state(B)
        gm = state["corrector_state"]["global_dry_air_mass"]
        assert gm.shape == (B, 1, 1) and gm.dtype == torch.float64
        o2, f2 = st.rollout(f1, forcing[T:], T, use_cuda_graph=graph, stepper_state=state)
        torch.testing.assert_close(torch.cat([o1, o2]), o_all, rtol=0, atol=0)
        torch.testing.assert_close(f2, f_all, rtol=0, atol=0)
        # without the carried state the second window re-pins the mass to ITS first state: the corrected pressure differs
        o2r, _ = st.rollout(f1, forcing[T:], T, use_cuda_graph=graph)
        gm2 = st.get_stepper_state(B)["corrector_state"]["global_dry_air_mass"]
        assert float((gm2 - gm).abs().max()) > 0  # the drifted state has a (slightly) different dry-air mean
    # (b) independent single steps on different inputs: each seeds from its own input
    s0 = {n: (torch.randn(B, *img, generator=g) * stds[n] + means[n]).cuda() for n in in_names}
    s1 = {n: v + 0.2 * stds[n] for n, v in s0.items()}
    for k in range(NZ):
        for d in (s0, s1):
            d[f"specific_total_water_{k}"] = d[f"specific_total_water_{k}"].clamp(min=0)
    a1 = st.step(s1)
    t1 = st.get_stepper_state(B)["corrector_state"]["global_dry_air_mass"].clone()
    st.step(s0)
    t0 = st.get_stepper_state(B)["corrector_state"]["global_dry_air_mass"].clone()
    assert float((t1 - t0).abs().min()) > 0
    b1 = st.step(s1)  # same result as the first call on s1: no memory of s0's target
    for n in out_names:
        torch.testing.assert_close(b1[n], a1[n], rtol=0, atol=0)
    # ... and with an explicit state the target is the given one
    c1 = st.step(s1, stepper_state={"corrector_state": {"global_dry_air_mass": t0}})
    assert not torch.equal(c1["PRESsfc"], a1["PRESsfc"])


# ---- the rest of the reference's sequence: zero-mean advection, frozen-precipitation clip, total energy budget ----------------
FLUXES = ["DLWRFsfc", "ULWRFsfc", "DSWRFsfc", "USWRFsfc", "SHTFLsfc", "USWRFtoa", "ULWRFtoa"]
TEMPS = [f"air_temperature_{k}" for k in range(NZ)]
PROG_E = PROG_ORDER[:3] + TEMPS[::-1] + PROG_ORDER[3:]                  # temperatures interleaved, reversed order
OUT_E = OUT[:5] + TEMPS[4:] + FLUXES[:3] + OUT[5:] + TEMPS[:4] + FLUXES[3:] + ["total_frozen_precipitation_rate"]
FORCING_E = ["land_fraction", "DSWRFtoa", "HGTsfc"]


def _energy_fields(g, B, H, W):
    ic = _fields(g, B, H, W, 1.0e5, False)
    inp = _fields(g, B, H, W, 1.0002e5, False)
    gen = _fields(g, B, H, W, 1.0007e5, True)
    for d in (ic, inp, gen):
        for k in range(NZ):
            d[f"air_temperature_{k}"] = 210.0 + 10.0 * k + 3.0 * torch.randn(B, H, W, generator=g)
    for n, m in zip(FLUXES, [340.0, 390.0, 190.0, 30.0, 20.0, 100.0, 240.0]):
        gen[n] = m + 20.0 * torch.randn(B, H, W, generator=g)
    gen["total_frozen_precipitation_rate"] = gen["PRATEsfc"] * 2.0 * torch.rand(B, H, W, generator=g)
    forcing = {"land_fraction": torch.rand(B, H, W, generator=g), "DSWRFtoa": 300.0 + 80.0 * torch.rand(B, H, W, generator=g),
               "HGTsfc": 800.0 * torch.randn(B, H, W, generator=g)}
    nxt = {"DSWRFtoa": 340.0 + 100.0 * torch.rand(B, H, W, generator=g), "HGTsfc": forcing["HGTsfc"] + 1.0}  # distinguishable
    return ic, inp, gen, forcing, nxt


def _temp(d):
    return torch.stack([d[f"air_temperature_{k}"] for k in range(NZ)], dim=-1)


def _oracle_full(inp, forcing, nxt, gen, target, w, ak, bk, mode, zero_adv, clip, heating):
    """The whole sequence of fme/core/corrector/atmosphere.py:349-398 after ForcePositive."""
    from oracle import corrector as oc
    from oracle import metrics as om

    vc = oc.VerticalCoordinate(ak, bk)

    def awm(data, keepdim=False, name=None):
        return om.weighted_mean(data, w.to(data.dtype), keepdim=keepdim)

    gen = dict(gen)
    adv = "tendency_of_total_water_path_due_to_advection"
    gen["PRESsfc"] = oc.adjust_dry_air_to_target(gen["PRESsfc"], _wat(gen), target, awm, vc)
    if zero_adv:
        gen[adv] = oc.zero_global_mean_moisture_advection(gen[adv], awm)
    if mode is not None:
        p, l, a = oc.conserve_moisture(inp["PRESsfc"], _wat(inp), gen["PRESsfc"], _wat(gen), gen["PRATEsfc"], gen["LHTFLsfc"], awm, vc, 21600.0, mode)
        gen["PRATEsfc"], gen["LHTFLsfc"] = p, l
        if a is not None:
            gen[adv] = a
        if clip:
            gen["total_frozen_precipitation_rate"] = oc.clip_frozen_precipitation(gen["total_frozen_precipitation_rate"], gen["PRATEsfc"])
    if heating is not None:
        fl = dict(dsw_toa=nxt["DSWRFtoa"], usw_toa=gen["USWRFtoa"], ulw_toa=gen["ULWRFtoa"], dlw_sfc=gen["DLWRFsfc"], ulw_sfc=gen["ULWRFsfc"],
                  dsw_sfc=gen["DSWRFsfc"], usw_sfc=gen["USWRFsfc"], lhf=gen["LHTFLsfc"], shf=gen["SHTFLsfc"], frozen=gen["total_frozen_precipitation_rate"])
        t = oc.conserve_total_energy(inp["PRESsfc"], _temp(inp), _wat(inp), forcing["HGTsfc"], gen["PRESsfc"], _temp(gen), _wat(gen), nxt["HGTsfc"],
                                     fl, awm, vc, 21600.0, heating)
        for k in range(NZ):
            gen[f"air_temperature_{k}"] = t[..., k]
    return gen


@pytest.mark.parametrize("mode,zero_adv,clip,heating", [
    ("advection_and_precipitation", False, True, 0.0),   # configs/baselines/era5/ace-train-config-1-step-pretrain.yaml:119-124
    ("advection_and_precipitation", False, False, 1.14),  # configs/baselines/shield-som/ace-train-config.yaml:140-145
    ("evaporation", True, True, 0.0),                     # corrected LHF and clipped frozen precipitation feed the energy flux
    ("precipitation", True, False, None),                 # zero-mean advection survives (the mode does not recompute advection)
    (None, True, False, 0.5),
])
def test_full_corrector_sequence_matches_oracle(mode, zero_adv, clip, heating):
    from ace_b200.corrector import AtmosphereCorrector
    from oracle import corrector as oc
    from oracle import metrics as om

    B, H, W = 2, 24, 48
    g = torch.Generator().manual_seed(21)
    ak, bk, w = _coords(H, W)
    ic, inp, gen, forcing, nxt = _energy_fields(g, B, H, W)
    vc = oc.VerticalCoordinate(ak, bk)
    target = oc.seed_global_dry_air_mass(ic["PRESsfc"], _wat(ic), lambda d, keepdim=False: om.weighted_mean(d, w.to(d.dtype), keepdim=keepdim), vc)
    ref = _oracle_full(inp, forcing, nxt, gen, target, w, ak, bk, mode, zero_adv, clip, heating)
    dbl = lambda d: {k: v.double() for k, v in d.items()}  # noqa: E731
    ref64 = _oracle_full(dbl(inp), dbl(forcing), dbl(nxt), dbl(gen), target, w.double(), ak.double(), bk.double(), mode, zero_adv, clip, heating)

    c = AtmosphereCorrector(OUT_E, PROG_E, (H, W), ak, bk, w, 21600.0, conserve_dry_air=True, moisture_budget_correction=mode,
                            zero_global_mean_moisture_advection=zero_adv, clip_frozen_precipitation=clip,
                            total_energy_budget_correction=None if heating is None else dict(method="constant_temperature", constant_unaccounted_heating=heating),
                            forcing_names=FORCING_E)
    pack = lambda d, names: torch.stack([d[n] for n in names], dim=1).contiguous().cuda()  # noqa: E731
    out, prev, nprog = pack(gen, OUT_E), pack(inp, PROG_E), pack(gen, PROG_E)
    c.seed(pack(ic, PROG_E))
    if heating is not None:
        with pytest.raises(ValueError):
            c.apply(prev, out.clone(), nprog.clone())
    c.apply(prev, out, nprog, prev_forcing=pack(forcing, FORCING_E), next_step=pack(nxt, ["DSWRFtoa", "HGTsfc"]))
    changed = set()
    for i, n in enumerate(OUT_E):
        got, want = out[:, i].cpu(), ref[n]
        if torch.equal(want, gen[n]):
            assert torch.equal(got, want), f"{n}: field the reference leaves alone was modified"
            continue
        changed.add(n)
        e32 = field_rel_err(got[:, None], want[:, None])
        e64 = field_rel_err(got[:, None].double(), ref64[n][:, None])
        r64 = field_rel_err(want[:, None].double(), ref64[n][:, None])
        assert e32 < 3e-5, (n, e32)
        assert e64 < 2e-6 and e64 <= r64 + 2e-7, (n, e64, r64)
    if heating is not None:
        assert set(TEMPS) <= changed
        # the temperature offset itself (a global-mean energy difference of ~1e9 J/m2 columns): compare the offsets, not T
        d_dev = (out[:, OUT_E.index("air_temperature_3")].cpu().double() - gen["air_temperature_3"].double()).mean(dim=(-1, -2))
        d_64 = (ref64["air_temperature_3"] - gen["air_temperature_3"].double()).mean(dim=(-1, -2))
        d_32 = (ref["air_temperature_3"].double() - gen["air_temperature_3"].double()).mean(dim=(-1, -2))
        assert float((d_dev - d_64).abs().max()) <= float((d_32 - d_64).abs().max()) + 2e-5, (d_dev, d_32, d_64)
        assert float(d_64.abs().min()) > 1e-3
    if clip and mode is not None:
        assert "total_frozen_precipitation_rate" in changed
    if zero_adv and mode in (None, "precipitation", "evaporation"):
        assert "tendency_of_total_water_path_due_to_advection" in changed
    for i, n in enumerate(PROG_E):
        assert torch.equal(nprog[:, i], out[:, OUT_E.index(n)]), n


def test_fused_rollout_with_full_corrector_graph_equals_eager_and_oracle_step():
    """FusedStepper with the ERA5 baseline's corrector block: one step against the oracle sequence, then graph == eager and
    rollout_host == rollout over 3 steps (the energy correction reads DSWRFtoa / HGTsfc of the next forcing time)."""
    import ace_b200
    from oracle import sfno as osfno
    from tests.test_gpu_stepper import _oracle_step

    img = (32, 64)
    in_names = PROG_E + FORCING_E
    out_names = OUT_E
    allnames = sorted(set(in_names + out_names))
    means, stds = {n: 0.0 for n in allnames}, {n: 1.0 for n in allnames}
    means["PRESsfc"], stds["PRESsfc"] = 1.0e5, 800.0
    for k in range(NZ):
        means[f"specific_total_water_{k}"], stds[f"specific_total_water_{k}"] = 1e-3 * (k + 1), 4e-4 * (k + 1)
        means[f"air_temperature_{k}"], stds[f"air_temperature_{k}"] = 210.0 + 10.0 * k, 4.0
    for n, m in zip(FLUXES, [340.0, 390.0, 190.0, 30.0, 20.0, 100.0, 240.0]):
        means[n], stds[n] = m, 20.0
    means["PRATEsfc"], stds["PRATEsfc"] = 3e-5, 1.5e-5
    means["total_frozen_precipitation_rate"], stds["total_frozen_precipitation_rate"] = 3e-5, 3e-5
    means["LHTFLsfc"], stds["LHTFLsfc"] = 90.0, 30.0
    stds["tendency_of_total_water_path_due_to_advection"] = 1e-5
    means["DSWRFtoa"], stds["DSWRFtoa"] = 340.0, 100.0
    means["HGTsfc"], stds["HGTsfc"] = 400.0, 800.0
    torch.manual_seed(0)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, len(in_names), len(out_names), embed_dim=16, num_layers=2, operator_type="dhconv").eval()
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=16, num_layers=2, operator_type="dhconv"))
    net = sel.build(len(in_names), len(out_names), ace_b200.DatasetInfo(img_shape=img)).torch_module
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval().requires_grad_(False)
    ak, bk, w = _coords(*img)
    fp = [f"specific_total_water_{k}" for k in range(NZ)] + ["PRATEsfc", "total_frozen_precipitation_rate"]
    st = ace_b200.FusedStepper(
        net, in_names, out_names, means, stds, residual_prediction=True, force_positive_names=fp,
        corrector=dict(conserve_dry_air=True, moisture_budget_correction="advection_and_precipitation", clip_frozen_precipitation=True,
                       total_energy_budget_correction=dict(method="constant_temperature", constant_unaccounted_heating=1.14),
                       ak=ak, bk=bk, area_weights=w, timestep_seconds=21600.0))
    assert st.forcing_names == FORCING_E and st.corrector_needs_next
    B, T = 2, 3
    g = torch.Generator().manual_seed(13)
    state = {n: torch.randn(B, *img, generator=g) * stds[n] + means[n] for n in in_names}
    for k in range(NZ):
        state[f"specific_total_water_{k}"] = state[f"specific_total_water_{k}"].clamp(min=0)
    nxt = {"DSWRFtoa": 340.0 + 100.0 * torch.rand(B, *img, generator=g), "HGTsfc": state["HGTsfc"]}
    # one step vs the oracle sequence
    from oracle import corrector as oc
    from oracle import metrics as om

    vc = oc.VerticalCoordinate(ak, bk)
    ref = _oracle_step(onet, in_names, out_names, means, stds, True, state)
    for n in fp:
        ref[n] = torch.clamp(ref[n], min=0.0)
    target = oc.seed_global_dry_air_mass(state["PRESsfc"], _wat(state), lambda d, keepdim=False: om.weighted_mean(d, w.to(d.dtype), keepdim=keepdim), vc)
    ref = _oracle_full(state, state, nxt, ref, target, w, ak, bk, "advection_and_precipitation", False, True, 1.14)
    st.reset_corrector_state()
    out = st.step({n: v.cuda() for n, v in state.items()}, next_step_input_data={n: v.cuda() for n, v in nxt.items()})
    for n in out_names:
        a = ((out[n].cpu() - means[n]) / stds[n])[:, None]
        b = ((ref[n] - means[n]) / stds[n])[:, None]
        tol = 1e-3 if n == "tendency_of_total_water_path_due_to_advection" else 1e-4  # see the moisture test above
        # both sides hold the DENORMALISED field in fp32: with |mean| = 17..70 std here (fluxes, temperatures) one ulp of the field
        # is up to 4e-6 std, which is not small against a random-init network's 1e-2-std outputs; allow two ulps of the field
        ulp = 2.0 * 2.0 ** -23 * (abs(means[n]) + float(b.abs().max()) * stds[n]) / stds[n]
        err = float((a - b).abs().amax())
        assert err < tol * float(b.abs().amax()) + ulp, (n, err, float(b.abs().amax()), ulp)
    # rollouts: forcing at T + 1 times
    prog0 = torch.stack([state[n] for n in st.prognostic_names], dim=1).cuda()
    fm = torch.tensor([means[n] for n in FORCING_E])[None, None, :, None, None]
    fs = torch.tensor([stds[n] for n in FORCING_E])[None, None, :, None, None]
    forcing = (torch.randn(T + 1, B, len(FORCING_E), *img, generator=g) * fs + fm).cuda()
    with pytest.raises(ValueError):
        st.rollout(prog0, forcing[:T], T, use_cuda_graph=False)
    oe, fe = st.rollout(prog0, forcing, T, use_cuda_graph=False)
    og, fg = st.rollout(prog0, forcing, T, use_cuda_graph=True)
    torch.testing.assert_close(og, oe, rtol=0, atol=0)
    torch.testing.assert_close(fg, fe, rtol=0, atol=0)
    out_host = torch.empty(T, B, len(out_names), *img).pin_memory()
    st.rollout_host(prog0, forcing.cpu().pin_memory(), T, out_host)
    torch.cuda.synchronize()
    torch.testing.assert_close(out_host, og.cpu(), rtol=0, atol=0)


@pytest.mark.parametrize("interpolate", [True, False])
def test_fused_step_with_slab_ocean_matches_oracle(interpolate):
    """The SHiELD-SOM baseline's post-step block (configs/baselines/shield-som/ace-train-config.yaml:133-146): full corrector
    sequence, then the slab ocean (fme/core/ocean.py:64-88) through the prescriber; graph rollout == eager."""
    import ace_b200
    from oracle import corrector as oc
    from oracle import metrics as om
    from oracle import ocean as oo
    from oracle import sfno as osfno
    from tests.test_gpu_stepper import _oracle_step

    img = (32, 64)
    sst = "surface_temperature"
    forcing_names = FORCING_E + ["ocean_fraction"]
    in_names = PROG_E + [sst] + forcing_names
    out_names = OUT_E + [sst]
    allnames = sorted(set(in_names + out_names))
    means, stds = {n: 0.0 for n in allnames}, {n: 1.0 for n in allnames}
    means["PRESsfc"], stds["PRESsfc"] = 1.0e5, 800.0
    for k in range(NZ):
        means[f"specific_total_water_{k}"], stds[f"specific_total_water_{k}"] = 1e-3 * (k + 1), 4e-4 * (k + 1)
        means[f"air_temperature_{k}"], stds[f"air_temperature_{k}"] = 210.0 + 10.0 * k, 4.0
    for n, m in zip(FLUXES, [340.0, 390.0, 190.0, 30.0, 20.0, 100.0, 240.0]):
        means[n], stds[n] = m, 20.0
    means["PRATEsfc"], stds["PRATEsfc"] = 3e-5, 1.5e-5
    means["total_frozen_precipitation_rate"], stds["total_frozen_precipitation_rate"] = 3e-5, 3e-5
    means["LHTFLsfc"], stds["LHTFLsfc"] = 90.0, 30.0
    stds["tendency_of_total_water_path_due_to_advection"] = 1e-5
    means["DSWRFtoa"], stds["DSWRFtoa"] = 340.0, 100.0
    means["HGTsfc"], stds["HGTsfc"] = 400.0, 800.0
    means[sst], stds[sst] = 288.0, 12.0
    torch.manual_seed(0)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, len(in_names), len(out_names), embed_dim=16, num_layers=2, operator_type="dhconv").eval()
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=16, num_layers=2, operator_type="dhconv"))
    net = sel.build(len(in_names), len(out_names), ace_b200.DatasetInfo(img_shape=img)).torch_module
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval().requires_grad_(False)
    ak, bk, w = _coords(*img)
    fp = [f"specific_total_water_{k}" for k in range(NZ)] + ["PRATEsfc", "total_frozen_precipitation_rate"]
    st = ace_b200.FusedStepper(
        net, in_names, out_names, means, stds, residual_prediction=True, force_positive_names=fp,
        ocean=dict(surface_temperature_name=sst, ocean_fraction_name="ocean_fraction", interpolate=interpolate,
                   slab=dict(mixed_layer_depth_name="prescribed_mixed_layer_depth", q_flux_name="prescribed_qflux", timestep_seconds=21600.0)),
        corrector=dict(conserve_dry_air=True, moisture_budget_correction="advection_and_precipitation",
                       total_energy_budget_correction=dict(method="constant_temperature", constant_unaccounted_heating=1.14),
                       ak=ak, bk=bk, area_weights=w, timestep_seconds=21600.0))
    assert st.n_ocean == 3
    B = 2
    g = torch.Generator().manual_seed(17)
    state = {n: torch.randn(B, *img, generator=g) * stds[n] + means[n] for n in in_names}
    for k in range(NZ):
        state[f"specific_total_water_{k}"] = state[f"specific_total_water_{k}"].clamp(min=0)
    nxt = {"DSWRFtoa": 340.0 + 100.0 * torch.rand(B, *img, generator=g), "HGTsfc": state["HGTsfc"],
           "ocean_fraction": torch.rand(B, *img, generator=g), "prescribed_qflux": 30.0 * torch.randn(B, *img, generator=g),
           "prescribed_mixed_layer_depth": 20.0 + 60.0 * torch.rand(B, *img, generator=g)}
    vc = oc.VerticalCoordinate(ak, bk)
    ref = _oracle_step(onet, in_names, out_names, means, stds, True, state)
    for n in fp:
        ref[n] = torch.clamp(ref[n], min=0.0)
    target = oc.seed_global_dry_air_mass(state["PRESsfc"], _wat(state), lambda d, keepdim=False: om.weighted_mean(d, w.to(d.dtype), keepdim=keepdim), vc)
    ref = _oracle_full(state, state, nxt, ref, target, w, ak, bk, "advection_and_precipitation", False, False, 1.14)
    f_net = oo.net_surface_energy_flux_without_frozen_precip(ref["DLWRFsfc"], ref["ULWRFsfc"], ref["DSWRFsfc"], ref["USWRFsfc"], ref["LHTFLsfc"],
                                                             ref["SHTFLsfc"])
    t_slab = oo.slab_surface_temperature(state[sst], f_net, nxt["prescribed_qflux"], nxt["prescribed_mixed_layer_depth"], 21600.0)
    ref[sst] = oo.prescribe(nxt["ocean_fraction"], ref[sst], t_slab, interpolate)
    st.reset_corrector_state()
    out = st.step({n: v.cuda() for n, v in state.items()}, next_step_input_data={n: v.cuda() for n, v in nxt.items()})
    for n in out_names:
        a = ((out[n].cpu() - means[n]) / stds[n])[:, None]
        b = ((ref[n] - means[n]) / stds[n])[:, None]
        tol = 1e-3 if n == "tendency_of_total_water_path_due_to_advection" else 1e-4
        ulp = 2.0 * 2.0 ** -23 * (abs(means[n]) + float(b.abs().max()) * stds[n]) / stds[n]
        err = float((a - b).abs().amax())
        assert err < tol * float(b.abs().amax()) + ulp, (n, err, float(b.abs().amax()), ulp)
    assert float((ref[sst] - _oracle_step(onet, in_names, out_names, means, stds, True, state)[sst]).abs().max()) > 0.1  # the ocean acted
    # rollouts: graph == eager, host-pipelined == device
    T = 3
    prog0 = torch.stack([state[n] for n in st.prognostic_names], dim=1).cuda()
    fm = torch.tensor([means[n] for n in st.forcing_names])[None, None, :, None, None]
    fs = torch.tensor([stds[n] for n in st.forcing_names])[None, None, :, None, None]
    forcing = (torch.randn(T + 1, B, len(st.forcing_names), *img, generator=g) * fs + fm).cuda()
    ocean = torch.stack([torch.rand(T, B, *img, generator=g), 30.0 * torch.randn(T, B, *img, generator=g),
                         20.0 + 60.0 * torch.rand(T, B, *img, generator=g)], dim=2).cuda()
    oe, fe = st.rollout(prog0, forcing, T, use_cuda_graph=False, ocean_seq=ocean)
    og, fg = st.rollout(prog0, forcing, T, use_cuda_graph=True, ocean_seq=ocean)
    torch.testing.assert_close(og, oe, rtol=0, atol=0)
    torch.testing.assert_close(fg, fe, rtol=0, atol=0)
    out_host = torch.empty(T, B, len(out_names), *img).pin_memory()
    st.rollout_host(prog0, forcing.cpu().pin_memory(), T, out_host, ocean_host=ocean.cpu().pin_memory())
    torch.cuda.synchronize()
    torch.testing.assert_close(out_host, og.cpu(), rtol=0, atol=0)
