"""tcgen05 GEMM core vs torch fp32 on the same bf16 inputs (all operand modes)."""
import pytest
import torch
import torch.nn.functional as F

from from_ddpm_to_stable_diffusion_b200 import _lib

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _close(got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"max abs err {err} vs scale {scale}"


def pack_conv(w):  # OIHW fp32 -> [co][tap][ci] bf16
    co, ci = w.shape[:2]
    return _bf(w.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous())


@pytest.mark.parametrize("M,N,c0,c1", [(256, 128, 128, 0), (320, 256, 128, 64), (64, 128, 64, 0), (4096, 384, 128, 0)])
def test_gemm_fwd(cuda, M, N, c0, c1):
    g = torch.Generator(device="cuda").manual_seed(1)
    a0 = _bf(torch.randn(M, c0, device=cuda, generator=g))
    a1 = _bf(torch.randn(M, c1, device=cuda, generator=g)) if c1 else None
    w = _bf(torch.randn(N, c0 + c1, device=cuda, generator=g) * 0.1)
    bias = torch.randn(N, device=cuda, generator=g)
    rps = 64
    rb = torch.randn((M + rps - 1) // rps, N, device=cuda, generator=g)
    res = _bf(torch.randn(M, N, device=cuda, generator=g))
    d = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    _lib.call("tsd_gemm_fwd", a0, a1, c0, c1, M, w, N, bias, rb, rps, res, 0, d)
    a = a0.float() if a1 is None else torch.cat([a0, a1], 1).float()
    ref = a @ w.float().t() + bias + rb.repeat_interleave(rps, 0)[:M] + res.float()
    torch.cuda.synchronize()
    _close(d, ref)
    # no epilogue extras
    _lib.call("tsd_gemm_fwd", a0, a1, c0, c1, M, w, N, None, None, 1, None, 0, d)
    _close(d, a @ w.float().t())


def test_gemm_geglu(cuda):
    M, C = 512, 128
    g = torch.Generator(device="cuda").manual_seed(2)
    a = _bf(torch.randn(M, C, device=cuda, generator=g))
    w = torch.randn(8 * C, C, device=cuda, generator=g) * 0.1
    b = torch.randn(8 * C, device=cuda, generator=g) * 0.1
    # pack: tile t (128 rows) = 64 value rows then the 64 matching gate rows
    H = 4 * C
    idx = torch.arange(8 * C, device=cuda).view(-1, 128)
    t = torch.arange(idx.shape[0], device=cuda)[:, None]
    j = torch.arange(64, device=cuda)[None, :]
    perm = torch.cat([t * 64 + j, H + t * 64 + j], 1).reshape(-1)
    wp = _bf(w[perm].contiguous())
    bp = b[perm].contiguous()
    d = torch.empty(M, H, device=cuda, dtype=torch.bfloat16)
    _lib.call("tsd_gemm_fwd", a, None, C, 0, M, wp, 8 * C, bp, None, 1, None, 1, d)
    h = a.float() @ _bf(w).float().t() + b
    ref = h[:, :H] * F.gelu(h[:, H:])
    torch.cuda.synchronize()
    _close(d, ref)


@pytest.mark.parametrize("n,H,c0,c1,cout", [(3, 16, 128, 0, 128), (2, 8, 128, 128, 256), (5, 32, 64, 0, 128)])
def test_groupnorm_statistics_from_the_producing_epilogue(cuda, n, H, c0, c1, cout):
    """conv3x3 / GEMM with gn=True leave per-channel partial sums of their output behind; GroupNorm statistics from those
    partials (no pass over the tensor) equal the stand-alone statistics kernel, also for a channel concat of two such
    tensors (decoder skip connections)."""
    from from_ddpm_to_stable_diffusion_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n * H + cout)
    hw = H * H
    x0 = _bf(torch.randn(n * hw, c0, device=cuda, generator=g))
    x1 = _bf(torch.randn(n * hw, c1, device=cuda, generator=g)) if c1 else None
    w = pack_conv(torch.randn(cout, c0 + c1, 3, 3, device=cuda, generator=g) * 0.05)
    bias = torch.randn(cout, device=cuda, generator=g)
    res = _bf(torch.randn(n * hw, cout, device=cuda, generator=g))
    y = ops.conv3x3(x0, n, H, H, w, cout, x1=x1, bias=bias, residual=res, gn=True)
    assert torch.equal(y, ops.conv3x3(x0, n, H, H, w, cout, x1=x1, bias=bias, residual=res))  # the by-product changes nothing
    wl = _bf(torch.randn(cout, cout, device=cuda, generator=g) * 0.1)
    z = ops.gemm(y, wl, cout, bias=bias, residual=y, gn=True)
    scratch = torch.zeros(ops.gn_scratch_floats(n), device=cuda)
    for (a0, a1) in ((y, None), (z, None), (y, z)):
        got = ops.gn_stats(a0, n, hw, 1e-5, scratch, x1=a1)          # from the partials
        b0 = a0.clone()
        b1 = a1.clone() if a1 is not None else None                  # clones carry no partials: the stand-alone kernel
        ref = ops.gn_stats(b0, n, hw, 1e-5, scratch, x1=b1)
        assert hasattr(a0, "_gn_part") and not hasattr(b0, "_gn_part")
        assert (got[..., 0] - ref[..., 0]).abs().max().item() < 1e-5
        assert ((got[..., 1] - ref[..., 1]).abs() / ref[..., 1]).max().item() < 1e-5
        full = a0.float() if a1 is None else torch.cat([a0, a1], 1).float()
        m_ref = full.view(n, hw, 32, -1).permute(0, 2, 1, 3).reshape(n, 32, -1).mean(-1)
        assert (got[..., 0] - m_ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,C", [(1000, 128), (4096, 256), (77, 128)])
def test_gemm_geglu_bwd_recompute(cuda, M, C):
    """Backward of value * gelu(gate) with the pre-activations recomputed inside the GEMM (no stored 8C-wide tensor)
    and the bias gradient as a by-product, against torch autograd of the erf-form GEGLU on the same bf16 inputs."""
    from from_ddpm_to_stable_diffusion_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + C)
    H = 4 * C
    a = _bf(torch.randn(M, C, device=cuda, generator=g))
    w = torch.randn(8 * C, C, device=cuda, generator=g) * 0.1
    b = torch.randn(8 * C, device=cuda, generator=g) * 0.1
    dgg = _bf(torch.randn(M, H, device=cuda, generator=g))
    wp = ops.pack_linear(w.contiguous(), geglu=True)
    bp = ops.pack_geglu_bias(b.contiguous())
    dbias = torch.ones(8 * C, device=cuda)
    dh8 = ops.gemm_geglu_bwd(a, wp, bp, dgg, dbias=dbias)
    h = (a.float() @ _bf(w).float().t() + b).requires_grad_(True)
    (h[:, :H] * F.gelu(h[:, H:])).backward(dgg.float())
    torch.cuda.synchronize()
    assert dh8.shape == (M, 8 * C)
    _close(dh8, h.grad, tol=1.5e-2)
    rel = ((dh8.float() - h.grad).norm() / h.grad.norm()).item()
    assert rel < 1e-2, rel
    cs_ref = h.grad.sum(0)
    assert ((dbias - 1.0) - cs_ref).abs().max().item() < 2e-2 * cs_ref.abs().max().item() + 0.05 * M ** 0.5 * 0.05
    assert torch.equal(ops.gemm_geglu_bwd(a, wp, bp, dgg), dh8)  # the by-product does not change the result


@pytest.mark.parametrize("n,H,W,c0,c1,cout,stride", [
    (2, 16, 16, 64, 0, 128, 1), (3, 8, 8, 128, 128, 256, 1), (1, 64, 64, 64, 64, 128, 1),
    (2, 32, 32, 128, 0, 128, 2), (2, 16, 16, 64, 0, 128, 2), (4, 4, 4, 64, 0, 128, 1), (8, 2, 2, 64, 0, 128, 1),
    (2, 32, 32, 128, 128, 256, 1), (3, 32, 64, 64, 0, 128, 1), (150, 32, 32, 64, 0, 128, 1), (3, 16, 32, 64, 64, 128, 1), (2, 16, 24, 64, 0, 128, 1),  # halo-mode patches
])
def test_conv3x3_fwd(cuda, n, H, W, c0, c1, cout, stride):
    g = torch.Generator(device="cuda").manual_seed(3)
    cin = c0 + c1
    x = torch.randn(n, cin, H, W, device=cuda, generator=g)
    w = torch.randn(cout, cin, 3, 3, device=cuda, generator=g) * 0.05
    bias = torch.randn(cout, device=cuda, generator=g)
    rb = torch.randn(n, cout, device=cuda, generator=g)
    xh = _bf(x.permute(0, 2, 3, 1).contiguous())
    x0 = xh[..., :c0].contiguous()
    x1 = xh[..., c0:].contiguous() if c1 else None
    Ho, Wo = H // stride, W // stride
    res = _bf(torch.randn(n, Ho, Wo, cout, device=cuda, generator=g))
    d = torch.empty(n, Ho, Wo, cout, device=cuda, dtype=torch.bfloat16)
    _lib.call("tsd_conv3x3_fwd", x0, x1, c0, c1, n, H, W, stride, pack_conv(w), cout, bias, rb, 0, res, d)
    ref = F.conv2d(xh.float().permute(0, 3, 1, 2), _bf(w).float(), bias, stride=stride, padding=1)
    ref = ref.permute(0, 2, 3, 1) + rb[:, None, None, :] + res.float()
    torch.cuda.synchronize()
    _close(d, ref)


def test_gemm_dgrad(cuda):
    M, N, K = 384, 256, 128
    g = torch.Generator(device="cuda").manual_seed(4)
    dy = _bf(torch.randn(M, N, device=cuda, generator=g))
    w = _bf(torch.randn(N, K, device=cuda, generator=g) * 0.1)
    res = _bf(torch.randn(M, K, device=cuda, generator=g))
    dx = torch.empty(M, K, device=cuda, dtype=torch.bfloat16)
    _lib.call("tsd_gemm_dgrad", dy, M, N, w, K, res, dx)
    torch.cuda.synchronize()
    _close(dx, dy.float() @ w.float() + res.float())


@pytest.mark.parametrize("n,H,W,cin,cout", [(2, 16, 16, 128, 64), (2, 8, 8, 256, 128), (1, 64, 64, 128, 128),
                                            (3, 32, 32, 256, 128), (2, 64, 32, 128, 64)])
def test_conv3x3_dgrad(cuda, n, H, W, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = _bf(torch.randn(n, H, W, cout, device=cuda, generator=g))
    w = torch.randn(cout, cin, 3, 3, device=cuda, generator=g) * 0.05
    dx = torch.empty(n, H, W, cin, device=cuda, dtype=torch.bfloat16)
    _lib.call("tsd_conv3x3_dgrad", dy, n, H, W, cout, pack_conv(w), cin, None, dx)
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), _bf(w).float(), stride=1, padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    _close(dx, ref)


def test_gemm_wgrad(cuda):
    M, N, c0, c1 = 1024, 256, 128, 128
    g = torch.Generator(device="cuda").manual_seed(6)
    dy = _bf(torch.randn(M, N, device=cuda, generator=g))
    x0 = _bf(torch.randn(M, c0, device=cuda, generator=g))
    x1 = _bf(torch.randn(M, c1, device=cuda, generator=g))
    dw = torch.ones(N, c0 + c1, device=cuda)
    _lib.call("tsd_gemm_wgrad", dy, x0, x1, c0, c1, M, N, dw)
    ref = 1.0 + dy.float().t() @ torch.cat([x0, x1], 1).float()
    torch.cuda.synchronize()
    _close(dw, ref, tol=1e-3)


@pytest.mark.parametrize("n,H,W,c0,c1,cout,stride", [
    (2, 16, 16, 128, 0, 128, 1), (4, 8, 8, 128, 128, 256, 1), (1, 64, 64, 128, 0, 128, 1), (2, 32, 32, 128, 0, 128, 2),
    (3, 32, 16, 128, 128, 256, 1), (40, 32, 32, 256, 0, 128, 1), (2, 8, 8, 128, 0, 64, 1),
])
def test_conv3x3_wgrad(cuda, n, H, W, c0, c1, cout, stride):
    g = torch.Generator(device="cuda").manual_seed(7)
    cin = c0 + c1
    Ho, Wo = H // stride, W // stride
    x = _bf(torch.randn(n, H, W, cin, device=cuda, generator=g))
    dy = _bf(torch.randn(n, Ho, Wo, cout, device=cuda, generator=g))
    x0 = x[..., :c0].contiguous()
    x1 = x[..., c0:].contiguous() if c1 else None
    dw = torch.zeros(cout, 9 * cin, device=cuda)
    _lib.call("tsd_conv3x3_wgrad", dy, x0, x1, c0, c1, n, H, W, stride, cout, dw)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.zeros(cout, cin, 3, 3, device=cuda, requires_grad=True)
    out = F.conv2d(xr, wref, stride=stride, padding=1)
    out.backward(dy.float().permute(0, 3, 1, 2))
    ref = wref.grad.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    torch.cuda.synchronize()
    _close(dw, ref, tol=1e-3)
