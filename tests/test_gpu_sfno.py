"""GPU: the SFNO forward through the module / C ABI vs committed reference vectors and the oracle."""
import pytest
import torch

from tests.util import NET_GOLDENS, field_rel_err, golden_net_fields, golden_state_dict, load_golden

pytestmark = pytest.mark.gpu

# Stated tolerance (BASELINE.json north_star: "rtol 1e-4 on prognostic fields after one step"):
# for every (sample, field):  max|y - y_ref| <= 1e-4 * max|y_ref|.
FIELD_RTOL = 1e-4


def _b200_net(img_shape, cin, cout, fields):
    import ace_b200

    builder = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields)
    return builder.build(cin, cout, ace_b200.DatasetInfo(img_shape=tuple(img_shape))).torch_module


@pytest.mark.parametrize("name", NET_GOLDENS)
def test_net_matches_reference_vectors(name):
    g = load_golden(name)
    fields = golden_net_fields(g)
    net = _b200_net(g["img_shape"], int(g["in_chans"]), int(g["out_chans"]), fields)
    net.load_state_dict(golden_state_dict(g))
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]).cuda())
    ref = torch.from_numpy(g["output"])
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert field_rel_err(y.cpu(), ref) < FIELD_RTOL


@pytest.mark.parametrize(
    "img,cin,cout,embed,layers,batch",
    [
        ((48, 96), 6, 7, 32, 3, 2),      # every GEMM eligible for the tcgen05 kernel
        ((180, 360), 8, 8, 16, 2, 1),    # BASELINE configs[0] grid, reduced width
        ((64, 128), 5, 3, 64, 2, 3),
        ((48, 96), 4, 4, 128, 2, 2),     # mlp hidden 256 -> fc1 runs on the CTA-pair (cta_group::2) variant
        ((45, 96), 5, 3, 32, 2, 1),      # odd nlat (as 721x1440): element-wise store variants, still no SIMT fallback
        ((45, 96), 5, 3, 128, 1, 2),     # odd nlat with the space-on-rows convolution variants (ragged last position tile)
        ((36, 72), 3, 5, 384, 2, 2),     # ACE2 width: 192-channel column tiles, per-sample folded weights of two members
    ],
)
def test_net_tcgen05_path_vs_oracle(img, cin, cout, embed, layers, batch):
    from ace_b200 import _lib
    from oracle import sfno as osfno

    fields = dict(embed_dim=embed, num_layers=layers, operator_type="dhconv")
    torch.manual_seed(5)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, cin, cout, **fields).eval()
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for k, p in onet.named_parameters():
            if k.endswith("bias") or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if k.endswith("filter.filter.weight"):
                p.mul_(p.shape[0])
    net = _b200_net(img, cin, cout, fields)
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval()
    x = torch.randn(batch, cin, *img, generator=g)
    u0, s0 = _lib.get_option("count_umma"), _lib.get_option("count_simt")
    with torch.no_grad():
        y = net(x.cuda())
        ref = onet(x)
    torch.cuda.synchronize()
    assert _lib.get_option("count_simt") == s0, "a GEMM fell back to the SIMT kernel"
    assert _lib.get_option("count_umma") - u0 == 4 + 8 * layers
    assert field_rel_err(y.cpu(), ref) < FIELD_RTOL
    # SIMT kernels on the same problem agree with the tensor-core path
    _lib.set_option("force_simt", 1)
    try:
        with torch.no_grad():
            y2 = net(x.cuda())
    finally:
        _lib.set_option("force_simt", 0)
    assert field_rel_err(y2.cpu(), ref) < FIELD_RTOL
    assert field_rel_err(y.cpu(), y2.cpu()) < FIELD_RTOL


@pytest.mark.parametrize("embed,img", [(32, (48, 96)), (136, (40, 80)), (384, (24, 48)), (32, (45, 96)), (64, (180, 360))])
def test_dhconv_orientations_agree(embed, img):
    """Both dhconv GEMM orientations (orders on the rows = cplx 2 / weights on the rows = cplx 1) against the oracle."""
    from ace_b200 import _lib
    from oracle import sfno as osfno

    fields = dict(embed_dim=embed, num_layers=2, operator_type="dhconv")
    torch.manual_seed(21)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, 5, 4, **fields).eval()
    with torch.no_grad():
        for k, p in onet.named_parameters():
            if k.endswith("filter.filter.weight"):
                p.mul_(p.shape[0])
    net = _b200_net(img, 5, 4, fields)
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval()
    x = torch.randn(2, 5, *img)
    with torch.no_grad():
        ref = onet(x)
    outs = []
    s0 = _lib.get_option("count_simt")
    base = dict(dhconv_t=0, inv2=1, tile_list=1)
    try:
        # defaults, then each second-generation piece switched back to its first-generation form
        for v in (dict(), dict(dhconv_t=1), dict(inv2=0), dict(tile_list=0), dict(dhconv_t=1, inv2=0, tile_list=0)):
            for k, val in {**base, **v}.items():
                _lib.set_option(k, val)
            with torch.no_grad():
                outs.append(net(x.cuda()).cpu())
    finally:
        for k, val in base.items():
            _lib.set_option(k, val)
    assert _lib.get_option("count_simt") == s0
    for y in outs:
        assert field_rel_err(y, ref) < FIELD_RTOL
        assert field_rel_err(outs[0], y) < FIELD_RTOL


def test_module_contract():
    import ace_b200

    fields = dict(embed_dim=16, num_layers=2, operator_type="dhconv")
    net = _b200_net((16, 32), 3, 4, fields)
    with pytest.raises(ace_b200.AceError):
        net(torch.zeros(1, 3, 16, 32))  # CPU input: no fallback
    net = net.cuda()
    x = torch.randn(2, 3, 16, 32, device="cuda")
    with pytest.raises(ace_b200.AceError):
        net(x)  # gradients enabled on trainable parameters
    with torch.no_grad():
        y1 = net(x)
        # batch size may change between calls (ensemble folding); results are per-sample independent
        y2 = net(x[:1])
        torch.testing.assert_close(y1[:1], y2, rtol=1e-5, atol=1e-6)
        # parameter edits are picked up (version counter) ...
        net.decoder[2].weight.mul_(2.0)
        y3 = net(x)
        torch.testing.assert_close(y3, 2 * y1, rtol=1e-4, atol=1e-6)
        # ... and so is load_state_dict
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        sd["decoder.2.weight"] = sd["decoder.2.weight"] / 2
        net.load_state_dict(sd)
        torch.testing.assert_close(net(x), y1, rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError):
        with torch.no_grad():
            net(torch.zeros(1, 5, 16, 32, device="cuda"))


def test_offload_parameters_single_weight_residency():
    """``offload_parameters()``: the fp32 torch parameters leave the GPU (the library's split planes are the only device copy),
    the forward is unchanged bit for bit, and edits of the host copies are re-uploaded."""
    import ace_b200

    net = _b200_net((32, 64), 3, 4, dict(embed_dim=64, num_layers=2, operator_type="dhconv")).cuda().eval().requires_grad_(False)
    with pytest.raises(ace_b200.AceError):
        net.offload_parameters()  # nothing uploaded yet
    x = torch.randn(2, 3, 32, 64, device="cuda")
    y1 = net(x)
    torch.cuda.synchronize()
    n_param_bytes = sum(p.numel() * 4 for p in net.parameters())
    m0 = torch.cuda.memory_allocated()
    net.offload_parameters()
    assert all(p.device.type == "cpu" for p in net.parameters())
    assert m0 - torch.cuda.memory_allocated() >= 0.95 * n_param_bytes
    assert torch.equal(net(x), y1)
    net.decoder[2].weight.mul_(2.0)  # host-side edit
    torch.testing.assert_close(net(x), 2 * y1, rtol=1e-4, atol=1e-6)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd["decoder.2.weight"] = sd["decoder.2.weight"] / 2
    net.load_state_dict(sd)
    assert torch.equal(net(x), y1)
    net.cuda()  # and back
    assert torch.equal(net(x), y1)


@pytest.mark.timeout(900)
def test_baseline_config_full_size_parity():
    """BASELINE.json configs[1] at full size: ACE2 1 degree (180x360, 44 in / 50 out, embed 384, 8 blocks, dhconv,
    instance_norm), B=1, one step; CUDA path through the C ABI vs the CPU oracle, rtol 1e-4 per output field."""
    from ace_b200 import _lib
    from oracle import sfno as osfno

    fields = dict(embed_dim=384, num_layers=8, operator_type="dhconv")
    torch.manual_seed(11)
    onet = osfno.SphericalFourierNeuralOperatorNet((180, 360), 44, 50, **fields).eval()
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for k, p in onet.named_parameters():
            if k.endswith("bias") or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if k.endswith("filter.filter.weight"):
                p.mul_(p.shape[0])  # O(1) spectral gain so the filter branch matters at random init
    net = _b200_net((180, 360), 44, 50, fields)
    net.load_state_dict(onet.state_dict())
    net = net.cuda().eval()
    x = torch.randn(1, 44, 180, 360, generator=g)
    s0 = _lib.get_option("count_simt")
    torch.set_num_threads(min(16, torch.get_num_threads()))
    with torch.no_grad():
        y = net(x.cuda())
        ref = onet(x)
    torch.cuda.synchronize()
    assert _lib.get_option("count_simt") == s0, "a GEMM fell back to the SIMT kernel at the benchmark shape"
    err = field_rel_err(y.cpu(), ref)
    assert err < FIELD_RTOL, err
