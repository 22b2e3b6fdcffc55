"""GPU: the split-bf16 GEMM kernels through the C ABI (ace_dev_gemm): SIMT vs fp64 truth, tcgen05 vs SIMT."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a, b, m, n, k, nb, layout, impl):
    from ace_b200 import _lib

    d = torch.empty(nb, m, n, device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().ace_dev_gemm(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                        ctypes.c_void_p(d.data_ptr()), m, n, k, nb, int(layout), impl, None))
    torch.cuda.synchronize()
    return d


def _case(m, n, k, nb, layout, seed=0):
    """layout bit 0: A stored [z][k][m]; bit 1: B stored [z][k][n]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(nb, m, k, generator=g)
    b = torch.randn(nb, n, k, generator=g)
    truth = torch.einsum("zmk,znk->zmn", a.double(), b.double())
    a_dev = (a.transpose(1, 2).contiguous() if layout & 1 else a).cuda()
    b_dev = (b.transpose(1, 2).contiguous() if layout & 2 else b).cuda()
    return a_dev, b_dev, truth


# 3-term split products: |err| <~ 3 * 2^-18 * sum|a||b|  -> relative to sqrt(k) scale ~1e-5
TOL = 3e-5


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(70, 50, 33, 2), (128, 64, 64, 1), (300, 130, 100, 3)])
def test_simt_gemm_matches_fp64(shape, layout):
    m, n, k, nb = shape
    a, b, truth = _case(m, n, k, nb, layout)
    d = _gemm(a, b, m, n, k, nb, layout, impl=0)
    err = (d.double().cpu() - truth).abs().max() / truth.abs().max()
    assert err < 1e-5, err  # SIMT path multiplies the re-joined hi+lo values in fp32


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize(
    "shape",
    [
        (128, 128, 64, 1),      # one tile, two K chunks
        (128, 128, 256, 1),     # pipeline wraps
        (256, 192, 128, 1),     # two M tiles, N tile 192
        (1000, 384, 200, 2),    # ragged M, K tail (zero filled by TMA), batch
        (136, 52, 40, 3),       # N < tile, K tail inside a chunk
        (64800 // 8, 384, 384, 1),  # conv-like
        (520, 768, 96, 1),      # N tile 256
        (384, 8104, 48, 2),     # conv orientation: wide N with a ragged last tile (8 valid columns)
    ],
)
def test_umma_gemm_matches_simt_and_fp64(shape, layout):
    m, n, k, nb = shape
    a, b, truth = _case(m, n, k, nb, layout, seed=1)
    d1 = _gemm(a, b, m, n, k, nb, layout, impl=1)
    d0 = _gemm(a, b, m, n, k, nb, layout, impl=0)
    scale = truth.abs().max()
    err_truth = (d1.double().cpu() - truth).abs().max() / scale
    err_simt = (d1.double() - d0.double()).abs().max().cpu() / scale
    assert err_truth < TOL and err_simt < TOL, (float(err_truth), float(err_simt))


@pytest.mark.parametrize("bn", [192, 256])
def test_umma_forced_n_tiles(bn):
    from ace_b200 import _lib

    m, n, k, nb = 384, 400, 192, 1
    a, b, truth = _case(m, n, k, nb, 1, seed=2)
    _lib.set_option("umma_bn", bn)
    try:
        d1 = _gemm(a, b, m, n, k, nb, 1, impl=1)
    finally:
        _lib.set_option("umma_bn", 0)
    err = (d1.double().cpu() - truth).abs().max() / truth.abs().max()
    assert err < TOL, float(err)


def test_umma_rejects_unaligned_shapes():
    from ace_b200 import AceError

    a, b, _ = _case(64, 50, 32, 1, 0)
    with pytest.raises(AceError):
        _gemm(a, b, 64, 50, 32, 1, 0, impl=1)  # n % 4 != 0: SIMT-only, the tcgen05 entry must say so


def test_umma_single_term_is_plain_bf16():
    from ace_b200 import _lib

    m, n, k, nb = 256, 128, 128, 1
    a, b, truth = _case(m, n, k, nb, 0, seed=3)
    _lib.set_option("split_terms", 1)
    try:
        d1 = _gemm(a, b, m, n, k, nb, 0, impl=1)
    finally:
        _lib.set_option("split_terms", 3)
    ref = torch.einsum("zmk,znk->zmn", a.bfloat16().double(), b.bfloat16().double()).cpu()
    err = (d1.double().cpu() - ref).abs().max() / ref.abs().max()
    assert err < 1e-5, float(err)  # exactly the hi*hi product
    assert (d1.double().cpu() - truth).abs().max() / truth.abs().max() > 1e-4  # and visibly worse than 3-term


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize(
    "shape",
    [
        (256, 256, 64, 1),      # exactly one pair tile
        (256, 256, 256, 1),     # pipeline wraps
        (384, 512, 96, 1),      # second pair tile half empty (rows 256..383 only on the leader)
        (1000, 384, 200, 2),    # ragged M, ragged N tile (128 valid columns), batch
        (384, 8104, 48, 2),     # conv orientation: wide N with a ragged last tile (168 valid columns)
        (520, 776, 96, 1),      # last N tile has 8 valid columns: 4 per CTA of the pair
    ],
)
def test_umma_pair_gemm_matches_fp64(shape, layout):
    """cta_group::2 variants: two CTAs of a cluster share one 256 x 256 tile."""
    from ace_b200 import _lib

    m, n, k, nb = shape
    a, b, truth = _case(m, n, k, nb, layout, seed=4)
    _lib.set_option("pair", 1)
    try:
        d1 = _gemm(a, b, m, n, k, nb, layout, impl=1)
    finally:
        _lib.set_option("pair", -1)
    err = (d1.double().cpu() - truth).abs().max() / truth.abs().max()
    assert err < TOL, float(err)


@pytest.mark.parametrize(
    "shape",
    [
        (128, 128, 256, 1),     # one position tile, one channel chunk group
        (384, 8104, 48, 2),     # 192-channel column tiles, ragged last position tile, batch
        (768, 2000, 96, 1),     # 256-channel column tiles
        (160, 4100, 320, 1),    # ragged channel chunk (160 = 5 x 32), position count not a multiple of 4 tiles
    ],
)
def test_umma_space_on_rows_variant(shape):
    """Cfg::SP: a 1x1-convolution-shaped product (activations MN-major) executed with the operand roles exchanged, CTA pairs."""
    from ace_b200 import _lib

    m, n, k, nb = shape
    a, b, truth = _case(m, n, k, nb, 2, seed=7)
    _lib.set_option("sp", 2)
    try:
        d1 = _gemm(a, b, m, n, k, nb, 2, impl=1)
    finally:
        _lib.set_option("sp", 1)
    err = (d1.double().cpu() - truth).abs().max() / truth.abs().max()
    assert err < TOL, float(err)
