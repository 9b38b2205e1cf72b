"""Unit tests of the raw CUDA kernels through the C ABI (B200 only)."""
import numpy as np
import pytest

from ped_b200 import api

pytestmark = pytest.mark.gpu


def _dev_u64(arr):
    import torch

    return torch.from_numpy(arr.view(np.int64).copy()).cuda()


def _host_u64(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("n,bits", [(0, 40), (1, 40), (33, 40), (6143, 40), (6144, 40), (6145, 40), (100_003, 40),
                                    (1_000_000, 62), (3_000_001, 64), (5_000_000, 8), (2_000_000, 17)])
def test_radix_sort_u64(ctx, n, bits):
    import torch

    rng = np.random.default_rng(n + bits)
    a = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if bits < 64:
        a &= np.uint64((1 << bits) - 1)
    d = _dev_u64(a) if n else torch.empty(0, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.test_sort_u64(d, n, bits)
    assert (_host_u64(d) == np.sort(a)).all()


def test_radix_sort_skewed_digits(ctx):
    """All keys in one bin / two bins: the look-back chain carries everything."""
    import torch

    n = 1_500_000
    for a in (np.full(n, 0x123456789A, dtype=np.uint64),
              np.where(np.arange(n) % 3 == 0, np.uint64(7), np.uint64(0xFF00000000)).astype(np.uint64)):
        d = _dev_u64(a)
        torch.cuda.synchronize()
        ctx.test_sort_u64(d, n, 40)
        assert (_host_u64(d) == np.sort(a)).all()


@pytest.mark.parametrize("n,bits", [(1, 40), (6145, 40), (2_000_003, 40), (1_000_000, 64)])
def test_radix_sort_pairs_is_stable(ctx, n, bits):
    import torch

    rng = np.random.default_rng(n)
    a = rng.integers(0, 1 << 20, size=n, dtype=np.uint64)  # many ties
    if bits == 64:
        a = a << np.uint64(44) | a
    v = np.arange(n, dtype=np.uint32)
    d = _dev_u64(a)
    dv = torch.from_numpy(v.view(np.int32).copy()).cuda()
    torch.cuda.synchronize()
    ctx.test_sort_pairs(d, dv, n, bits)
    order = np.argsort(a, kind="stable")
    assert (_host_u64(d) == a[order]).all()
    assert (dv.cpu().numpy().view(np.uint32) == v[order]).all()


@pytest.mark.parametrize("n,distinct", [(1, 1), (2048, 1), (2049, 3), (4097, 3), (1_000_000, 1000),
                                        (3_000_000, 2_500_000), (2_000_000, 7), (5_000_000, 1)])
def test_run_length_count(ctx, n, distinct):
    import torch

    rng = np.random.default_rng(n + distinct)
    a = np.sort(rng.integers(0, distinct, size=n, dtype=np.uint64) * np.uint64(0x10001))
    d = _dev_u64(a)
    uk = torch.empty(n, dtype=torch.int64, device="cuda")
    uc = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    D = ctx.test_rle(d, n, uk, uc)
    want_k, want_c = np.unique(a, return_counts=True)
    assert D == len(want_k)
    got_k = _host_u64(uk[:D])
    got_c = uc[:D].cpu().numpy().view(np.uint32)
    order = np.argsort(got_k, kind="stable")  # rows are appended in no particular order
    assert (got_k[order] == want_k).all()
    assert (got_c[order] == want_c.astype(np.uint32)).all()


def test_synth_device_equals_host(ctx):
    import torch

    g = api.synth_genome(77, 50_000)
    want = api.synth_fastq(g, 9, 20_011, read_len=150, dup_ppm=50_000, first=7)
    gd = torch.from_numpy(g).cuda()
    out = torch.empty(len(want), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    p = api.SynthParams(9, len(g), 150, 2000, 1000, 50_000)
    ctx.synth_fastq_device(p, gd, 7, 20_011, out)
    assert out.cpu().numpy().tobytes() == want
