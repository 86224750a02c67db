"""Numerics of the tcgen05 contraction kernels of the PPO update (include/dnppo.h, csrc/dn_umma.cuh) against plain
PyTorch references of the same op (FP64 on the same device), through the C ABI."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from drl_dronenavigation_b200 import _lib as L
    return L


def split_planes(x):
    """FP32 [R, C] -> BF16 planes [2, R, C]: hi = bf16(x), lo = bf16(x - hi)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def join_planes(p, passes):
    return p[0].double() + (p[1].double() if passes == 3 else 0.0)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run_gemm(kind, passes, M, N, K, slices, a, b, bias=None, act=1, h=None, out=None, partial=None):
    L = _lib()
    L.check(L.lib().dn_mlp_gemm(kind, passes, M, N, K, slices, a.data_ptr(), b.data_ptr(), bias.data_ptr() if bias is not None else None,
                                act, h.data_ptr() if h is not None else None, out.data_ptr() if out is not None else None,
                                partial.data_ptr() if partial is not None else None, stream_ptr()), "dn_mlp_gemm")
    torch.cuda.synchronize()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


TOL = {3: 3e-5, 1: 2e-2}


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 64), (384, 256, 128), (1024, 512, 512), (4096, 512, 64), (2048, 256, 512)])
def test_forward_matches_fp64(passes, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    a, b = split_planes(x), split_planes(w)
    out = torch.zeros(2, M, N, dtype=torch.bfloat16, device="cuda")
    for act in (0, 1):
        run_gemm(0, passes, M, N, K, 1, a, b, bias=bias, act=act, out=out)
        z = join_planes(a, passes) @ join_planes(b, passes).t() + bias.double()
        ref = torch.tanh(z) if act else z
        got = join_planes(out, passes)
        assert rel(got, ref) < TOL[passes], (act, rel(got, ref))
        if passes == 3:      # against the un-split FP32 operands: this is the FP32-faithfulness claim
            z32 = x.double() @ w.double().t() + bias.double()
            ref32 = torch.tanh(z32) if act else z32
            assert rel(got, ref32) < 3e-5, (act, rel(got, ref32))


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 128), (1024, 512, 512), (2048, 512, 256), (512, 256, 64)])
def test_dgrad_matches_fp64(passes, M, N, K):
    """out[M,N] = (dY[M,K] W[K,N]) * (1 - H^2): W is read MN-major (no transposed copy)."""
    g = torch.Generator(device="cuda").manual_seed(7 * M + N + K)
    dy = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    hval = torch.tanh(torch.randn(M, N, device="cuda", generator=g))
    a, b, h = split_planes(dy), split_planes(w), split_planes(hval)
    out = torch.zeros(2, M, N, dtype=torch.bfloat16, device="cuda")
    run_gemm(1, passes, M, N, K, 1, a, b, h=h, out=out)
    hj = join_planes(h, passes)
    ref = (join_planes(a, passes) @ join_planes(b, passes)) * (1 - hj * hj)
    got = join_planes(out, passes)
    assert rel(got, ref) < TOL[passes], rel(got, ref)


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("Mo,No,rows,slices", [(128, 64, 64, 1), (128, 64, 512, 2), (256, 128, 1024, 4), (512, 512, 4096, 8), (512, 64, 2048, 4),
                                                (256, 512, 8192, 32)])
def test_wgrad_matches_fp64(passes, Mo, No, rows, slices):
    """partial[s] = dY[rows_s, Mo]^T X[rows_s, No]: both operands read MN-major, split over the batch rows."""
    g = torch.Generator(device="cuda").manual_seed(Mo + No + rows)
    dy = torch.randn(rows, Mo, device="cuda", generator=g)
    x = torch.randn(rows, No, device="cuda", generator=g)
    a, b = split_planes(dy), split_planes(x)
    partial = torch.zeros(slices, Mo, No, device="cuda")
    run_gemm(2, passes, Mo, No, rows, slices, a, b, partial=partial)
    aj, bj = join_planes(a, passes), join_planes(b, passes)
    rs = rows // slices
    for s in range(slices):
        ref = aj[s * rs:(s + 1) * rs].t() @ bj[s * rs:(s + 1) * rs]
        assert rel(partial[s].double(), ref) < TOL[passes], (s, rel(partial[s].double(), ref))
    assert rel(partial.double().sum(0), aj.t() @ bj) < TOL[passes]


def test_gemm_rejects_bad_shapes():
    L = _lib()
    a = torch.zeros(2, 128, 64, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(2, 128, 64, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(64, device="cuda")
    rc = L.lib().dn_mlp_gemm(0, 3, 100, 64, 64, 1, a.data_ptr(), a.data_ptr(), bias.data_ptr(), 1, None, out.data_ptr(), None, stream_ptr())
    assert rc < 0 and b"128" in L.lib().dn_last_error()
    rc = L.lib().dn_mlp_gemm(0, 2, 128, 64, 64, 1, a.data_ptr(), a.data_ptr(), bias.data_ptr(), 1, None, out.data_ptr(), None, stream_ptr())
    assert rc < 0
