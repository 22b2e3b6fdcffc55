"""Generate tests/golden/ref_live_healpix.npz from the LIVE reference cuHPX classes (build container only).

    python -m oracle.make_golden_healpix

Per nside: the per-ring quadrature weights the reference derives from its data files
(``fme.core.cuhpx.tools.apply_ring_weight``), random fields, their SHT (reference run field by field: its ring loops are
only valid for unbatched input), a random spectrum and its iSHT.
"""
import os

import numpy as np
import torch

from . import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref = refload.load_cuhpx()
    data = {"nsides": np.array([4, 8, 16])}
    for nside in (4, 8, 16):
        lmax = mmax = 2 * nside - 1
        w = ref.apply_ring_weight(nside)
        f, i = ref.SHT(nside, lmax=lmax, mmax=mmax, quad_weights="ring"), ref.iSHT(nside, lmax=lmax, mmax=mmax)
        torch.manual_seed(100 + nside)
        x = torch.randn(5, 12 * nside**2)
        c = torch.stack([f(x[k]) for k in range(x.shape[0])])
        spec = torch.complex(torch.randn(5, lmax, mmax), torch.randn(5, lmax, mmax))
        # only l >= m carries signal in a real field's spectrum; keep the full random array (the table is zero for l < m)
        y = torch.stack([i(spec[k].clone()) for k in range(spec.shape[0])])
        data.update({f"n{nside}.w": w, f"n{nside}.x": x.numpy(), f"n{nside}.sht": c.numpy(), f"n{nside}.spec": spec.numpy(),
                     f"n{nside}.isht": y.numpy()})
        print(f"nside {nside}: |c|max {float(c.abs().max()):.3f} |y|max {float(y.abs().max()):.3f}")
    np.savez(os.path.join(OUT, "ref_live_healpix.npz"), **data)


if __name__ == "__main__":
    main()
