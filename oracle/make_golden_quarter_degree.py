"""Fixture for the 721x1440 (BASELINE configs[3] grid) reduced-width parity test: the oracle's output for a seeded network and
input, SUBSAMPLED (every 8th latitude, every 16th longitude, plus the two polar rows) so that it fits the repository; the test
rebuilds the same weights / input from the same seeds (tests/test_gpu_parity_extra.py::_seeded).

    python -m oracle.make_golden_quarter_degree      (build container: ~2 minutes, 10 GB of host memory)
"""
import os

import numpy as np
import torch

from oracle import sfno as osfno

IMG, CIN, COUT, EMBED, LAYERS, SEED = (721, 1440), 3, 3, 8, 1, 61
LAT_IDX = sorted(set(range(0, 721, 8)) | {0, 1, 719, 720})
LON_STRIDE = 16


def seeded_weights_(module, seed):
    """The perturbation every seeded parity net of tests/test_gpu_parity_extra.py applies (same order for the oracle net and for
    the ace_b200 module: both create their parameters in the reference's order)."""
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in module.named_parameters():
            if k.endswith("bias") or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if k.endswith("filter.filter.weight"):
                p.mul_(p.shape[0])
    return g


def main():
    fields = dict(embed_dim=EMBED, num_layers=LAYERS, operator_type="dhconv", data_grid="legendre-gauss")
    torch.manual_seed(SEED)
    onet = osfno.SphericalFourierNeuralOperatorNet(IMG, CIN, COUT, **fields).eval()
    g = seeded_weights_(onet, SEED)
    x = torch.randn(1, CIN, *IMG, generator=g)
    with torch.no_grad():
        ref = onet(x)
    sub = ref[:, :, LAT_IDX][..., ::LON_STRIDE].numpy()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_quarter_degree_721x1440_embed8.npz")
    np.savez_compressed(out, ref_sub=sub, ref_absmax=ref.abs().amax(dim=(-2, -1)).numpy(), lat_idx=np.array(LAT_IDX), lon_stride=LON_STRIDE,
                        x_checksum=float(x.double().sum()), w_checksum=float(sum(p.double().sum() for p in onet.parameters())), seed=SEED)
    print(out, sub.shape, os.path.getsize(out))


if __name__ == "__main__":
    main()
