"""Pins the tcgen05 shared-memory descriptor conventions and TMA box layouts the kernels rely on (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vexpress_b200 import _ffi, ops
    _ffi.require_sm100()
    return ops


def _bytes(t):
    return t.contiguous().view(torch.uint8).reshape(-1)


def test_tma_3d_noswizzle_chunk_layout(ops):
    """A [rows, ld] bf16 matrix viewed as (8, rows, ld/8) lands as [chunk][row][8] = no-swizzle core matrices."""
    rows, ld = 256, 320
    src = torch.arange(rows * ld, device="cuda", dtype=torch.float32).remainder(4093).bfloat16().reshape(rows, ld)
    R, hd, col0, row0 = 128, 40, 80, 64
    raw = ops.probe_tma(src, [8, rows, ld // 8], [ld * 2, 16], [8, R, hd // 8], 0, [0, row0, col0 // 8], R * hd * 2)
    got = raw.view(torch.bfloat16).reshape(hd // 8, R, 8)
    want = src[row0:row0 + R, col0:col0 + hd].reshape(R, hd // 8, 8).permute(1, 0, 2)
    assert torch.equal(got, want)


def _kmajor_noswz(mat):
    """[rows, K] -> [K/8][rows][8] (core matrices of 8 rows x 16 B, contiguous)."""
    r, k = mat.shape
    return mat.reshape(r, k // 8, 8).permute(1, 0, 2).contiguous()


def test_umma_kmajor_noswizzle(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    N, K = 128, 48
    a = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    # K-major, no swizzle: LBO = byte stride between the two 8-wide K chunks, SBO = stride between 8-row groups
    d = ops.probe_umma(_bytes(_kmajor_noswz(a)), _bytes(_kmajor_noswz(b)), 128 * 16, 128, 0, N * 16, 128, 0, 0, 0,
                       N, K // 16, 2 * 128 * 16, 2 * N * 16)
    torch.cuda.synchronize()
    err = (d - ref).abs().max().item()
    print("kmajor noswizzle max err", err)
    assert err < 1e-2


def test_umma_mnmajor_b_noswizzle(ops):
    """B = V [keys(K), hd(N)] row-major staged as [hd/8][keys][8]: MN-major B operand of the PV product."""
    g = torch.Generator(device="cuda").manual_seed(2)
    keys, hd = 64, 48
    p = torch.randn(128, keys, device="cuda", generator=g).bfloat16()          # A: K-major [128, keys]
    v = torch.randn(keys, hd, device="cuda", generator=g).bfloat16()
    ref = p.float() @ v.float()
    v_img = v.reshape(keys, hd // 8, 8).permute(1, 0, 2).contiguous()          # [hd/8][keys][8]
    results = {}
    for name, (lbo, sbo) in {"lbo=128,sbo=keys*16": (128, keys * 16), "lbo=keys*16,sbo=128": (keys * 16, 128)}.items():
        d = ops.probe_umma(_bytes(_kmajor_noswz(p)), _bytes(v_img), 128 * 16, 128, 0, lbo, sbo, 0, 0, 1, hd, keys // 16,
                           2 * 128 * 16, 16 * 16)
        torch.cuda.synchronize()
        results[name] = (d - ref).abs().max().item()
    print("mn-major B hypotheses:", results)
    assert results["lbo=128,sbo=keys*16"] < 1e-2, results


def test_umma_a_from_tmem(ops):
    """TS-MMA: A = P [128 queries, keys] in tensor memory (lane = row, 32-bit column = two K-adjacent bf16, low half
    = even k), B = V [keys, hd] staged [hd/8][keys][8] (MN-major, no swizzle): the P.V product with P kept in TMEM."""
    g = torch.Generator(device="cuda").manual_seed(4)
    keys, hd = 128, 48
    p = torch.randn(128, keys, device="cuda", generator=g).bfloat16()
    v = torch.randn(keys, hd, device="cuda", generator=g).bfloat16()
    ref = p.float() @ v.float()
    a_packed = p.contiguous().view(torch.int32).contiguous()              # [128, keys/2]: (k even | k odd << 16)
    v_img = v.reshape(keys, hd // 8, 8).permute(1, 0, 2).contiguous()
    d = ops.probe_umma_ts(a_packed, keys, _bytes(v_img), 128, keys * 16, 0, 1, hd, 16 * 16)
    torch.cuda.synchronize()
    err = (d - ref).abs().max().item()
    print("TS-MMA (A in TMEM) max err", err)
    assert err < 2e-2
