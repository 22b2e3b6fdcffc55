"""Generate tests/golden/*.npz from the reference (run in the build container only).

TEST INFRASTRUCTURE.  ``python -m oracle.make_golden`` must be run where
``/root/reference`` exists.  It

  1. re-derives the inputs of the reference's own stored goldens from their
     documented seeds and saves input + stored output
       fme/core/benchmark/testdata/sht-regression.pt           (fme/sht_fix.py:259-266)
       fme/core/benchmark/testdata/inverse_sht-regression.pt   (fme/sht_fix.py:307-316)
       fme/ace/models/modulus/testdata/test_sfnonet_output_is_unchanged.pt
                                     (fme/ace/models/modulus/test_sfnonet.py:13-36)
  2. runs the LIVE reference classes (``oracle/refload.py``) on seeded inputs for
     the cases no stored vector pins -- every grid, truncated lmax/mmax, the
     ``dhconv`` operator (the ACE2 setting) -- and saves input, parameters, output;
  3. asserts that the oracle restatement reproduces every one of them before
     anything is written.

Outputs are small numpy archives; the reference sources are never copied.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    t = t.detach().cpu()
    if t.is_complex():
        return t.numpy().astype(np.complex64)
    return t.numpy()


def _set_seed0():
    # fme.core.rand.set_seed(0)  (fme/core/rand.py:20-36): numpy seed+1, random seed+2, torch seed+3
    np.random.seed(1)
    random.seed(2)
    torch.manual_seed(3)


def _close(a, b, rtol, atol, what):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


def main():
    from . import refload
    from . import sfno as osfno
    from . import sht as osht

    if not refload.available():
        sys.exit("reference tree not present; goldens can only be regenerated in the build container")
    ref = refload.load()
    root = refload.REFERENCE_ROOT
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---------------------------------------------------------------- 1. stored goldens
    _set_seed0()
    x = torch.randn(1, 9, 18)
    g_sht = torch.load(os.path.join(root, "fme/core/benchmark/testdata/sht-regression.pt"))["output"]
    g_isht = torch.load(os.path.join(root, "fme/core/benchmark/testdata/inverse_sht-regression.pt"))["output"]
    y = osht.RealSHT(9, 18)(x)
    _close(y, g_sht, 1.3e-6, 1e-5, "oracle RealSHT vs sht-regression.pt")
    _close(osht.InverseRealSHT(9, 18)(y), g_isht, 1.3e-6, 1e-5, "oracle InverseRealSHT vs inverse_sht-regression.pt")
    np.savez(
        os.path.join(OUT, "ref_stored_sht_regression.npz"),
        x=_np(x), sht_output=_np(g_sht), isht_output=_np(g_isht),
        nlat=9, nlon=18, grid="lobatto",
    )

    g_net = torch.load(os.path.join(root, "fme/ace/models/modulus/testdata/test_sfnonet_output_is_unchanged.pt"))
    torch.manual_seed(0)
    net = osfno.SphericalFourierNeuralOperatorNet(
        (9, 18), 2, 3, embed_dim=16, num_layers=2, operator_type="diagonal", data_grid="equiangular"
    )
    xin = torch.randn(4, 2, 9, 18)
    _close(net(xin), g_net, 1.3e-6, 1e-5, "oracle net vs test_sfnonet_output_is_unchanged.pt")
    sd = {f"sd.{k}": _np(v) for k, v in net.state_dict().items()}
    np.savez(
        os.path.join(OUT, "ref_stored_sfnonet_output_is_unchanged.npz"),
        x=_np(xin), output=_np(g_net), img_shape=(9, 18), in_chans=2, out_chans=3, embed_dim=16, num_layers=2,
        operator_type="diagonal", data_grid="equiangular", **sd,
    )

    # ---------------------------------------------------------------- 2. live reference, SHT on every grid
    cases = [
        # nlat, nlon, lmax, mmax, grid
        (9, 18, None, None, "lobatto"),
        (9, 18, None, None, "legendre-gauss"),
        (9, 18, None, None, "equiangular"),
        (12, 24, 8, 9, "legendre-gauss"),      # truncated
        (12, 24, 12, 16, "equiangular"),       # zero-padded mmax > nlon//2+1
        (16, 32, 16, 17, "legendre-gauss"),
        (45, 90, 45, 46, "legendre-gauss"),
        (33, 64, 20, 25, "equiangular"),       # odd nlat, truncated
    ]
    arrs = {}
    for idx, (nlat, nlon, lmax, mmax, grid) in enumerate(cases):
        torch.manual_seed(100 + idx)
        x = torch.randn(2, 3, nlat, nlon)
        rs = ref.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        ri = ref.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        c = rs(x)
        # spectral input for the inverse that is NOT a forward image (exercises Im(m=0)/Nyquist zeroing)
        torch.manual_seed(200 + idx)
        cin = torch.view_as_complex(torch.randn(2, 3, rs.lmax, rs.mmax, 2))
        xr = ri(cin.clone())
        oc = osht.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        oi = osht.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        _close(oc.weights, rs.weights, 0, 0, f"forward table {grid} {nlat}x{nlon}")
        _close(oi.pct, ri.pct, 0, 0, f"inverse table {grid} {nlat}x{nlon}")
        _close(oc(x), c, 1e-6, 1e-6, f"oracle RealSHT vs live reference {grid} {nlat}x{nlon}")
        _close(oi(cin), xr, 1e-6, 1e-6, f"oracle InverseRealSHT vs live reference {grid} {nlat}x{nlon}")
        arrs.update({
            f"c{idx}.meta": np.array([nlat, nlon, rs.lmax, rs.mmax]), f"c{idx}.grid": grid,
            f"c{idx}.x": _np(x), f"c{idx}.sht": _np(c), f"c{idx}.spec_in": _np(cin), f"c{idx}.isht": _np(xr),
        })
        if nlat <= 16:
            arrs[f"c{idx}.fwd_table"] = rs.weights.numpy()
            arrs[f"c{idx}.inv_table"] = ri.pct.numpy()
    arrs["ncases"] = len(cases)
    np.savez(os.path.join(OUT, "ref_live_sht_cases.npz"), **arrs)

    # ---------------------------------------------------------------- 3. live reference, nets (dhconv = ACE2 operator)
    nets = [
        # name, img_shape, in, out, builder fields, batch
        ("dhconv_9x18", (9, 18), 3, 4, dict(embed_dim=16, num_layers=2, operator_type="dhconv"), 2),
        ("dhconv_16x32_eq", (16, 32), 5, 5, dict(embed_dim=24, num_layers=3, operator_type="dhconv", data_grid="equiangular"), 1),
        ("diag_12x24", (12, 24), 2, 2, dict(embed_dim=8, num_layers=2, operator_type="diagonal"), 3),
        ("dhconv_32x64_nonorm", (32, 64), 4, 6, dict(embed_dim=32, num_layers=2, operator_type="dhconv", normalization_layer="none", big_skip=False, pos_embed=False), 1),
        ("ace2like_48x96", (48, 96), 7, 9, dict(embed_dim=32, num_layers=4, operator_type="dhconv"), 2),
    ]
    for name, img, cin, cout, fields, batch in nets:
        torch.manual_seed(7)
        rnet = refload.build_reference_net(img, cin, cout, **fields).eval()
        # de-trivialise: reference init leaves norm affine = (1, 0), all biases 0, spectral bias 0
        g = torch.Generator().manual_seed(11)
        for k, p in rnet.named_parameters():
            if k.endswith("bias") or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if k.endswith("filter.filter.weight"):
                p.mul_(p.shape[0])  # scale 1/C^2 -> 1/C so the spectral path matters numerically
        torch.manual_seed(13)
        x = torch.randn(batch, cin, *img)
        y = rnet(x)
        onet = osfno.SphericalFourierNeuralOperatorNet(
            img, cin, cout,
            **{k: v for k, v in fields.items()},
        ).eval()
        assert list(onet.state_dict().keys()) == list(rnet.state_dict().keys()), name
        onet.load_state_dict(rnet.state_dict())
        _close(onet(x), y, 1e-5, 1e-5 * float(y.abs().max()), f"oracle net vs live reference {name}")
        sd = {f"sd.{k}": _np(v) for k, v in rnet.state_dict().items()}
        np.savez(
            os.path.join(OUT, f"ref_live_net_{name}.npz"),
            x=_np(x), output=_np(y), img_shape=img, in_chans=cin, out_chans=cout,
            fields=repr(fields), **sd,
        )
        print(f"net {name}: ok, |y|max={float(y.abs().max()):.3f}")
    print("goldens written to", OUT)


if __name__ == "__main__":
    main()
