"""GPU: fused step (normalise/pack/net/unpack/denormalise) and rollout vs an oracle restatement."""
import pytest
import torch

from tests.util import field_rel_err

pytestmark = pytest.mark.gpu


def _oracle_step(onet, in_names, out_names, means, stds, residual, state):
    """fme/core/step/single_module.py:648-665 with corrector/ocean off, on name dicts (CPU, fp32)."""
    prog = [n for n in out_names if n in in_names]
    norm = {n: (state[n] - means[n]) / stds[n] for n in in_names}
    x = torch.stack([norm[n] for n in in_names], dim=1)  # Packer.pack (fme/core/packer.py:45-52)
    with torch.no_grad():
        y = onet(x)
    out = {n: y[:, i] for i, n in enumerate(out_names)}
    if residual:
        for n in prog:
            out[n] = out[n] + norm[n]
    return {n: out[n] * stds[n] + means[n] for n in out_names}


def _setup(residual):
    import ace_b200
    from oracle import sfno as osfno

    img = (32, 64)
    in_names = ["a", "b", "f1", "c", "f2"]
    out_names = ["c", "d1", "a", "b", "d2", "d3"]
    means = {n: 0.1 * (i - 3) for i, n in enumerate(sorted(set(in_names + out_names)))}
    stds = {n: 0.5 + 0.25 * i for i, n in enumerate(sorted(set(in_names + out_names)))}
    torch.manual_seed(0)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, len(in_names), len(out_names), embed_dim=16, num_layers=2, operator_type="dhconv").eval()
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=16, num_layers=2, operator_type="dhconv"))
    net = sel.build(len(in_names), len(out_names), ace_b200.DatasetInfo(img_shape=img)).torch_module
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval().requires_grad_(False)
    st = ace_b200.FusedStepper(net, in_names, out_names, means, stds, residual_prediction=residual)
    return img, in_names, out_names, means, stds, onet, st


@pytest.mark.parametrize("residual", [False, True])
def test_step_matches_oracle(residual):
    img, in_names, out_names, means, stds, onet, st = _setup(residual)
    assert st.prognostic_names == ["c", "a", "b"] and st.forcing_names == ["f1", "f2"]
    torch.manual_seed(1)
    state = {n: torch.randn(2, *img) * stds[n] + means[n] for n in in_names}
    ref = _oracle_step(onet, in_names, out_names, means, stds, residual, state)
    out = st.step({n: v.cuda() for n, v in state.items()})
    assert list(out.keys()) == out_names
    for n in out_names:
        # tolerance in normalised units (the denormalisation offset would otherwise hide errors)
        a = ((out[n].cpu() - means[n]) / stds[n])[:, None]
        b = ((ref[n] - means[n]) / stds[n])[:, None]
        assert field_rel_err(a, b) < 1e-4, n


def test_rollout_graph_equals_eager_and_oracle():
    img, in_names, out_names, means, stds, onet, st = _setup(True)
    torch.manual_seed(2)
    T, B = 4, 2
    prog0 = torch.randn(B, 3, *img).cuda()
    forcing = torch.randn(T, B, 2, *img).cuda()
    outs_e, fin_e = st.rollout(prog0, forcing, T, use_cuda_graph=False)
    outs_g, fin_g = st.rollout(prog0, forcing, T, use_cuda_graph=True)
    torch.testing.assert_close(outs_g, outs_e, rtol=0, atol=0)
    torch.testing.assert_close(fin_g, fin_e, rtol=0, atol=0)
    # oracle loop (fme/ace/stepper/single_module.py:1135-1167): feed prognostic outputs back
    state = {n: prog0[:, i].cpu() for i, n in enumerate(st.prognostic_names)}
    for t in range(T):
        full = dict(state)
        for j, n in enumerate(st.forcing_names):
            full[n] = forcing[t, :, j].cpu()
        out = _oracle_step(onet, in_names, out_names, means, stds, True, full)
        ref_t = torch.stack([out[n] for n in out_names], dim=1)
        assert field_rel_err(outs_g[t].cpu(), ref_t) < 3e-4 * (t + 1), t
        state = {n: out[n] for n in st.prognostic_names}


def test_graph_replay_picks_up_weight_changes():
    """parameters edited / reloaded after the CUDA graph was captured must reach the replays (the graph is cached
    on (B, device) only and a replay never passes through the eager upload path)."""
    img, in_names, out_names, means, stds, onet, st = _setup(False)
    torch.manual_seed(3)
    T, B = 3, 2
    prog0 = torch.randn(B, 3, *img).cuda()
    forcing = torch.randn(T, B, 2, *img).cuda()
    og0, _ = st.rollout(prog0, forcing, T, use_cuda_graph=True)  # captures
    net = st.module
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd["decoder.2.weight"] = sd["decoder.2.weight"] * 1.5
    sd["blocks.0.filter.filter.weight"] = sd["blocks.0.filter.filter.weight"] * 0.5
    sd["pos_embed"] = sd["pos_embed"] + 0.01
    net.load_state_dict(sd)
    og1, fg1 = st.rollout(prog0, forcing, T, use_cuda_graph=True)   # replays of the cached graph
    oe1, fe1 = st.rollout(prog0, forcing, T, use_cuda_graph=False)  # eager: uploads on every step
    assert not torch.equal(og1, og0)
    torch.testing.assert_close(og1, oe1, rtol=0, atol=0)
    torch.testing.assert_close(fg1, fe1, rtol=0, atol=0)
    with torch.no_grad():
        net.encoder[0].bias.add_(0.05)  # in-place edit (version counter)
    out_host = torch.empty(T, B, len(out_names), *img).pin_memory()
    st.rollout_host(prog0, forcing.cpu().pin_memory(), T, out_host)
    oe2, _ = st.rollout(prog0, forcing, T, use_cuda_graph=False)
    torch.cuda.synchronize()
    torch.testing.assert_close(out_host, oe2.cpu(), rtol=0, atol=0)
    assert not torch.equal(oe2, oe1)


def test_step_packed_rejects_bad_buffers():
    img, in_names, out_names, means, stds, onet, st = _setup(False)
    B = 2
    prog = torch.randn(B, 3, *img).cuda()
    forcing = torch.randn(B, 2, *img).cuda()
    with pytest.raises(ValueError):
        st.step_packed(prog, forcing, out=torch.empty(B, len(out_names), *img, device="cuda", dtype=torch.float16))
    with pytest.raises(ValueError):
        st.step_packed(prog, forcing, out=torch.empty(B, len(out_names) + 1, *img, device="cuda")[:, 1:])
    with pytest.raises(ValueError):
        st.step_packed(prog, forcing, next_prog=torch.empty(B, 3, img[0], 2 * img[1], device="cuda")[..., ::2])
    with pytest.raises(ValueError):
        st.step_packed(prog, forcing, out=torch.empty(B, len(out_names), *img))  # CPU buffer


def test_rollout_host_pipelined_equals_device_rollout():
    """Host-resident forcing / outputs with copies on side streams == the device-resident graph rollout."""
    img, in_names, out_names, means, stds, onet, st = _setup(False)
    torch.manual_seed(3)
    T, B = 5, 2
    prog0 = torch.randn(B, 3, *img).cuda()
    forcing = torch.randn(T, B, 2, *img)
    outs_d, fin_d = st.rollout(prog0, forcing.cuda(), T, use_cuda_graph=True)
    out_host = torch.empty(T, B, len(out_names), *img).pin_memory()
    fin_h = st.rollout_host(prog0, forcing.pin_memory(), T, out_host)
    torch.cuda.synchronize()
    torch.testing.assert_close(out_host, outs_d.cpu(), rtol=0, atol=0)
    torch.testing.assert_close(fin_h, fin_d, rtol=0, atol=0)
    # forcing window shorter than the rollout is cycled; a second call reuses the staging buffers
    fin_h2 = st.rollout_host(prog0, forcing[:2].pin_memory(), 4, out_host[:4])
    torch.cuda.synchronize()
    outs_c, fin_c = st.rollout(prog0, forcing[[0, 1, 0, 1]].cuda(), 4, use_cuda_graph=True)
    torch.testing.assert_close(out_host[:4], outs_c.cpu(), rtol=0, atol=0)
    torch.testing.assert_close(fin_h2, fin_c, rtol=0, atol=0)


def test_force_positive_and_prescribed_ocean_match_reference_semantics():
    """Post-step adjustments in the reference's order (fme/core/step/single_module.py:670-709): ForcePositive clamp
    (fme/core/corrector/utils.py:26-43), then the ocean prescriber (fme/core/prescriber.py:94-108, replace where
    round(mask) == 1, fme/core/spatial_masking.py:25-30)."""
    import ace_b200

    img, in_names, out_names, means, stds, onet, st0 = _setup(False)
    for interpolate in (False, True):
        st = ace_b200.FusedStepper(st0.module, in_names, out_names, means, stds, residual_prediction=False,
                                   force_positive_names=["d1", "a"],
                                   ocean=dict(surface_temperature_name="c", ocean_fraction_name="f2", interpolate=interpolate))
        torch.manual_seed(4)
        state = {n: torch.randn(2, *img) * stds[n] + means[n] for n in in_names}
        mask = torch.rand(2, *img)
        target = torch.randn(2, *img) + 280.0
        ref = _oracle_step(onet, in_names, out_names, means, stds, False, state)
        for n in ("d1", "a"):
            ref[n] = torch.clamp(ref[n], min=0.0)
        if interpolate:
            ref["c"] = mask * target + (1 - mask) * ref["c"]
        else:
            ref["c"] = torch.where(torch.round(mask).to(int) == 1, target, ref["c"])
        out = st.step({n: v.cuda() for n, v in state.items()}, next_step_input_data={"f2": mask.cuda(), "c": target.cuda()})
        for n in out_names:
            a = ((out[n].cpu() - means[n]) / stds[n])[:, None]
            b = ((ref[n] - means[n]) / stds[n])[:, None]
            assert field_rel_err(a, b) < 1e-4, (n, interpolate)
        assert (out["d1"] >= 0).all() and (out["a"] >= 0).all()
        # rollout: graph == eager with ocean data; the prescribed field is fed back as next state
        T = 3
        prog0 = torch.randn(2, 3, *img).cuda()
        forcing = torch.randn(T, 2, 2, *img).cuda()
        ocean = torch.stack([torch.rand(T, 2, *img), torch.randn(T, 2, *img) + 280.0], dim=2).cuda()
        oe, fe = st.rollout(prog0, forcing, T, use_cuda_graph=False, ocean_seq=ocean)
        og, fg = st.rollout(prog0, forcing, T, use_cuda_graph=True, ocean_seq=ocean)
        torch.testing.assert_close(og, oe, rtol=0, atol=0)
        torch.testing.assert_close(fg, fe, rtol=0, atol=0)
        out_host = torch.empty(T, 2, len(out_names), *img).pin_memory()
        fh = st.rollout_host(prog0, forcing.cpu().pin_memory(), T, out_host, ocean_host=ocean.cpu().pin_memory())
        torch.cuda.synchronize()
        torch.testing.assert_close(out_host, og.cpu(), rtol=0, atol=0)
        with pytest.raises(ValueError):
            st.step_packed(prog0, forcing[0])  # ocean configured but no ocean data


def test_predict_with_next_step_forcing_and_prescribed_prognostics():
    """``predict`` / ``predict_generator`` (the reference Stepper's API on name -> tensor mappings) through the C ABI and the CUDA
    graph: forcing windows, ``next_step_forcing_names`` read from the output time, ``prescribed_prognostic_names`` overwritten
    after the step; against the reference loop (fme/ace/stepper/single_module.py:1136-1167) on the oracle."""
    import ace_b200

    img, in_names, out_names, means, stds, onet, st0 = _setup(True)
    st = ace_b200.FusedStepper(st0.module, in_names, out_names, means, stds, residual_prediction=True,
                               next_step_forcing_names=["f2"], prescribed_prognostic_names=["b", "d1"])
    T, B = 3, 2
    torch.manual_seed(4)
    ic = {n: torch.randn(B, 1, *img) * stds[n] + means[n] for n in st.prognostic_names}
    forcing = {n: torch.randn(B, T + 1, *img) * stds[n] + means[n] for n in st.next_step_input_names}
    assert st.next_step_input_names == ["f1", "f2", "b", "d1"]
    # reference loop
    state = {n: ic[n][:, 0] for n in st.prognostic_names}
    ref = []
    for t in range(T):
        full = dict(state)
        full["f1"], full["f2"] = forcing["f1"][:, t], forcing["f2"][:, t + 1]
        out = _oracle_step(onet, in_names, out_names, means, stds, True, full)
        out["b"], out["d1"] = forcing["b"][:, t + 1], forcing["d1"][:, t + 1]
        ref.append(out)
        state = {n: out[n] for n in st.prognostic_names}
    cu = lambda d: {k: v.cuda() for k, v in d.items()}  # noqa: E731
    data, new_ic = st.predict(cu(ic), cu(forcing))
    data_e, new_ic_e = st.predict(cu(ic), cu(forcing), use_cuda_graph=False)
    for n in out_names:
        assert data[n].shape == (B, T, *img)
        torch.testing.assert_close(data[n], data_e[n], rtol=0, atol=0)  # graph replay == eager launches
        for t in range(T):
            a = ((data[n][:, t].cpu() - means[n]) / stds[n])[:, None]
            b = ((ref[t][n] - means[n]) / stds[n])[:, None]
            assert field_rel_err(a, b) < 3e-4 * (t + 1), (n, t)
    for n in ("b", "d1"):
        assert torch.equal(data[n].cpu(), forcing[n][:, 1:])
    for n in st.prognostic_names:
        torch.testing.assert_close(new_ic[n], data[n][:, -1:], rtol=0, atol=0)
        torch.testing.assert_close(new_ic[n], new_ic_e[n], rtol=0, atol=0)
    gen = list(st.predict_generator(cu(ic), cu(forcing), T))
    for t in range(T):
        for n in out_names:
            torch.testing.assert_close(gen[t][n], data[n][:, t], rtol=0, atol=0)
    # host-pipelined rollout with the same data: prescribed fields staged alongside the forcing
    prog0 = torch.stack([ic[n][:, 0] for n in st.prognostic_names], dim=1).cuda()
    fh = torch.stack([forcing["f1"][:, :T], forcing["f2"][:, 1:T + 1]], dim=2).transpose(0, 1).contiguous().pin_memory()
    ph = torch.stack([forcing["b"][:, 1:], forcing["d1"][:, 1:]], dim=2).transpose(0, 1).contiguous().pin_memory()
    oh = torch.empty(T, B, len(out_names), *img).pin_memory()
    fin = st.rollout_host(prog0, fh, T, out_host=oh, prescribed_host=ph)
    torch.cuda.synchronize()
    for i, n in enumerate(out_names):
        torch.testing.assert_close(oh[:, :, i].transpose(0, 1), data[n].cpu(), rtol=0, atol=0)
    for i, n in enumerate(st.prognostic_names):
        torch.testing.assert_close(fin[:, i], new_ic[n][:, 0], rtol=0, atol=0)
