"""Kernel-level checks on the GPU, through the C ABI (dalle_pytorch_b200.ops): every kernel against a plain
torch fp32 evaluation of the same op on the same seeded inputs.  fp32 mode: rtol 1e-3 / atol 1e-5
(north_star).  bf16 mode: inputs are rounded to bf16 first, tolerance 2e-2 relative to the tensor scale."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from util import report
from dalle_oracle import allowed_mask, token_shift, rotary_angle_table, apply_rotary

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-5


def dev():
    return torch.device('cuda:0')


def ops():
    from dalle_pytorch_b200 import ops as o
    return o


def bf16_tol(want):
    return dict(rtol=2e-2, atol=2e-2 * float(want.abs().max()) + 1e-6)


# ---- LayerNorm + token shift ------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [24, 19, 9, 5])
@pytest.mark.parametrize('do_shift', [True, False])
def test_ln_shift_fwd_bwd(n, do_shift):
    o = ops()
    torch.manual_seed(0)
    b, d, T, fm = 2, 64, 9, 4
    shift = do_shift and n >= T
    x = torch.randn(b, n, d, device=dev()) * 2 + 0.5
    w = torch.randn(d, device=dev()) * 0.2 + 1
    bias = torch.randn(d, device=dev()) * 0.2
    out, mean, rstd = o.ln_shift_fwd(x, w, bias, torch.float32, T, fm, do_ln=True, do_shift=shift)
    xr = x.detach().cpu().requires_grad_()
    wr, br = w.cpu().requires_grad_(), bias.cpu().requires_grad_()
    y = F.layer_norm(xr, (d,), wr, br, 1e-5)
    if shift:
        y = token_shift(y, T, fm)
    report('ln_shift_fwd', out.view(b, n, d), y, RTOL, ATOL)
    gy = torch.randn(b, n, d)
    dres = torch.randn(b, n, d)
    y.backward(gy)
    dg = torch.zeros(d, device=dev()); db = torch.zeros(d, device=dev())
    dx = o.ln_shift_bwd(gy.to(dev()).view(b * n, d).contiguous(), x, mean, rstd, w, dres.to(dev()), T, fm, do_ln=True, do_shift=shift,
                        dgamma=dg, dbeta=db)
    report('ln_shift_bwd dx', dx, xr.grad + dres, RTOL, 1e-5)
    report('ln_shift_bwd dgamma', dg, wr.grad, RTOL, 1e-4)
    report('ln_shift_bwd dbeta', db, br.grad, RTOL, 1e-4)


def test_ln_shift_full_size_property():
    """C2 geometry (n=1280, d=1024): un-shifted output rows are plain LayerNorm; shifted halves equal neighbours."""
    o = ops()
    torch.manual_seed(1)
    b, n, d, T, fm = 2, 1280, 1024, 257, 32
    x = torch.randn(b, n, d, device=dev())
    w = torch.ones(d, device=dev()); bias = torch.zeros(d, device=dev())
    out, _, _ = o.ln_shift_fwd(x, w, bias, torch.float32, T, fm)
    out = out.view(b, n, d)
    ln = F.layer_norm(x, (d,))
    report('pass-through half', out[..., d // 2:], ln[..., d // 2:], RTOL, ATOL)
    report('text shift', out[:, 1:T, :d // 2], ln[:, :T - 1, :d // 2], RTOL, ATOL)
    report('image top', out[:, T + fm:, :d // 4], ln[:, T:n - fm, :d // 4], RTOL, ATOL)
    assert out[:, 0, :d // 2].abs().max() == 0 and out[:, T:T + fm, :d // 4].abs().max() == 0
    assert out[:, T::fm, d // 4:d // 2].abs().max() == 0


# ---- GEMM + epilogues ---------------------------------------------------------------------------------------
def _mk(shape, dtype, scale=1.0):
    return (torch.randn(*shape, device=dev()) * scale).to(dtype)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize('M,N,K', [(96, 64, 72), (257, 192, 128), (130, 320, 264)])
def test_gemm_store_all_majors(dtype, a_mn, b_mn, M, N, K):
    o = ops()
    torch.manual_seed(2)
    A = _mk((K, M) if a_mn else (M, K), dtype)
    B = _mk((K, N) if b_mn else (N, K), dtype)
    bias = torch.randn(N, device=dev())
    want = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t()) + bias
    got = o.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, bias=bias)
    tol = dict(rtol=RTOL, atol=1e-4) if dtype == torch.float32 else dict(rtol=1e-3, atol=1e-3 * math.sqrt(K))
    report(f'gemm_store {dtype} a_mn={a_mn} b_mn={b_mn}', got, want, **tol)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_qkv_epilogue(dtype):
    o = ops()
    torch.manual_seed(3)
    b, n, d, h, dh = 2, 24, 64, 2, 64
    A = _mk((b * n, d), dtype)
    W = _mk((3 * h * dh, d), dtype, d ** -0.5)
    ang = rotary_angle_table(9, 4, dh)
    from dalle_pytorch_b200.attention import rotary_tables
    cos_t, sin_t = rotary_tables(ang.to(dev()), dh)
    q, k, v = o.gemm_qkv(A, W, b, n, h, dh, cos_t, sin_t, dh ** -0.5)
    qkv = (A.float() @ W.float().t()).cpu().view(b, n, 3, h, dh).permute(2, 0, 3, 1, 4)
    wq, wk, wv = (apply_rotary(ang[:n], qkv[i]) for i in range(3))
    tol = dict(rtol=RTOL, atol=ATOL) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    report('q', q, wq * dh ** -0.5, **tol)
    report('k', k, wk, **tol)
    report('v', v, wv, **tol)
    # same math as plain GEMM + streaming head-split/rotary pass (the bf16 default)
    q2, k2, v2 = o.gemm_qkv_auto(A, W, b, n, h, dh, cos_t, sin_t, dh ** -0.5)
    report('q (streaming)', q2, wq * dh ** -0.5, **tol)
    report('k (streaming)', k2, wk, **tol)
    report('v (streaming)', v2, wv, **tol)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_resid_geglu_epilogues(dtype):
    o = ops()
    torch.manual_seed(4)
    M, d, H = 200, 64, 256
    A = _mk((M, d), dtype)
    W1 = _mk((2 * H, d), dtype, d ** -0.5)
    b1 = torch.randn(2 * H, device=dev()) * 0.1
    h, u = o.gemm_geglu(A, W1, b1)
    uw = A.float() @ W1.float().t() + b1
    hw = uw[:, :H] * F.gelu(uw[:, H:])
    tol = dict(rtol=RTOL, atol=ATOL) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    report('geglu u', u, uw, **tol)
    report('geglu h', h, hw, **tol)
    W2 = _mk((d, H), dtype, H ** -0.5)
    b2 = torch.randn(d, device=dev()) * 0.1
    resid = torch.randn(M, d, device=dev())
    scale = torch.rand(d, device=dev()) + 0.5
    out, y = o.gemm_resid(h, W2, b2, resid, scale, sign=-1.0, keep_y=True)
    yw = h.float() @ W2.float().t() + b2
    report('resid y', y, yw, **tol)
    report('resid out', out, resid - scale * yw, **tol)
    # GEGLU backward epilogue
    dy = _mk((M, d), dtype)
    du = o.gemm_geglu_bwd(dy, W2, u)
    dh = dy.float() @ W2.float()
    uf = u.float().requires_grad_()
    (uf[:, :H] * F.gelu(uf[:, H:])).backward(dh)
    report('geglu_bwd du', du, uf.grad, **tol)
    # the same adjoint as a streaming pass (plain dgrad GEMM + dalle_b200_geglu_bwd), with the bias gradient
    dh_t = o.gemm_store(dy, W2, a_mn=False, b_mn=True)
    du2, db1 = o.geglu_bwd(dh_t, u)
    report('geglu_bwd (streaming) du', du2, uf.grad, **tol)
    report('geglu_bwd (streaming) db1', db1, uf.grad.sum(0), rtol=2e-2, atol=2e-2 * float(uf.grad.sum(0).abs().max()) + 1e-4)


# ---- attention ----------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, allow, key_mask=None):
    s = q @ k.transpose(-1, -2)
    s = s.masked_fill(~allow, float('-inf'))
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float('-inf'))
    p = s.softmax(-1)
    return p @ v


PATTERNS = [('full', 0), ('axial_row', 1), ('axial_col', 2), ('conv_like', 3)]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('kind,code', PATTERNS)
@pytest.mark.parametrize('n,T,fm', [(24, 9, 4), (100, 37, 8), (191, 65, 12)])
def test_attention_fwd_bwd_patterns(dtype, kind, code, n, T, fm):
    o = ops()
    torch.manual_seed(5)
    b, h, dh = 2, 2, 64
    q = _mk((b, h, n, dh), dtype, dh ** -0.5)
    k = _mk((b, h, n, dh), dtype)
    v = _mk((b, h, n, dh), dtype)
    spec = o.AttnSpec(code, causal=True, text_len=T, fmap=fm, kernel_size=5 if kind == 'conv_like' else 0, dilation=1)
    out, lse = o.attn_fwd(spec, q, k, v)
    allow = allowed_mask(kind, n, n, T, fm).to(dev())
    qr, kr, vr = (t.float().detach().requires_grad_() for t in (q, k, v))
    want = _attn_ref(qr, kr, vr, allow)
    tol = dict(rtol=RTOL, atol=ATOL) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    report(f'attn_fwd {kind}', out.view(b, n, h, dh).permute(0, 2, 1, 3), want, **tol)
    g = _mk((b, n, h * dh), dtype)
    want.backward(g.float().view(b, n, h, dh).permute(0, 2, 1, 3))
    dqkv = o.attn_bwd(spec, q, k, v, out, lse, g, None, None, 1.0)
    dq, dk, dv = (dqkv.view(b, n, 3, h, dh)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    tolb = dict(rtol=RTOL, atol=2e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    report(f'attn_bwd dq {kind}', dq, qr.grad, **tolb)
    report(f'attn_bwd dk {kind}', dk, kr.grad, **tolb)
    report(f'attn_bwd dv {kind}', dv, vr.grad, **tolb)


def test_attention_static_mask_noncausal_keymask_and_cache():
    o = ops()
    torch.manual_seed(6)
    b, h, n, dh = 2, 2, 50, 64
    q, k, v = (_mk((b, h, n, dh), torch.float32, 0.3) for _ in range(3))
    # non-causal with key mask (CLIP text encoder use, dalle_pytorch.py:324)
    km = torch.rand(b, n, device=dev()) > 0.3
    km[:, 0] = True
    out, _ = o.attn_fwd(o.AttnSpec(0, causal=False), q, k, v, km.to(torch.uint8))
    want = _attn_ref(q, k, v, torch.ones(n, n, dtype=torch.bool, device=dev()), km)
    report('noncausal+keymask', out.view(b, n, h, dh).permute(0, 2, 1, 3), want, RTOL, ATOL)
    # static mask == axial predicate (transformer.py:333-350)
    T, fm = 15, 6
    n2 = T + fm * fm - 1
    q2, k2, v2 = (_mk((b, h, n2, dh), torch.float32, 0.3) for _ in range(3))
    import dalle_pytorch_b200 as D
    t = D.Transformer(dim=64, depth=1, seq_len=n2, heads=2, image_fmap_size=fm)
    sm = t._get_attention_mask('axial_col').to(dev()).to(torch.uint8).contiguous()
    a, _ = o.attn_fwd(o.AttnSpec(4, causal=True, static_mask=sm), q2, k2, v2)
    bb, _ = o.attn_fwd(o.AttnSpec(2, causal=True, text_len=T, fmap=fm), q2, k2, v2)
    report('static == axial_col', a, bb, 1e-6, 1e-6)
    # cached decoding: the last 3 queries against all keys == the tail of the full causal result
    full, _ = o.attn_fwd(o.AttnSpec(0, causal=True), q, k, v)
    tail, _ = o.attn_fwd(o.AttnSpec(0, causal=True), q[:, :, -3:].contiguous(), k, v)
    report('n_q < n_k', tail, full[:, -3:], 1e-6, 1e-6)


def test_attention_full_size_properties():
    """C2 geometry (n=1280): rows are convex combinations (constant V -> constant out); axial == static-mask dense."""
    o = ops()
    torch.manual_seed(7)
    b, h, n, dh, T, fm = 1, 2, 1280, 64, 257, 32
    q, k = (_mk((b, h, n, dh), torch.float32, 0.2) for _ in range(2))
    v = torch.ones(b, h, n, dh, device=dev()) * 0.75
    for code in (0, 1, 2, 3):
        out, lse = o.attn_fwd(o.AttnSpec(code, causal=True, text_len=T, fmap=fm, kernel_size=5, dilation=1), q, k, v)
        assert (out - 0.75).abs().max() < 1e-5 and torch.isfinite(lse).all()
    import dalle_pytorch_b200 as D
    t = D.Transformer(dim=64, depth=1, seq_len=n, heads=2, image_fmap_size=fm)
    v = _mk((b, h, n, dh), torch.float32)
    for kind, code in (('axial_row', 1), ('axial_col', 2)):
        sm = t._get_attention_mask(kind).to(dev()).to(torch.uint8).contiguous()
        a, _ = o.attn_fwd(o.AttnSpec(4, causal=True, static_mask=sm), q, k, v)
        bb, _ = o.attn_fwd(o.AttnSpec(code, causal=True, text_len=T, fmap=fm), q, k, v)
        report(f'static == {kind} @1280', a, bb, 1e-5, 1e-6)


# ---- glue ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_scale_bwd_colsum_cast_axpby(dtype):
    o = ops()
    torch.manual_seed(8)
    M, d = 777, 64
    g = torch.randn(M, d, device=dev())
    y = _mk((M, d), dtype)
    s = torch.rand(d, device=dev()) + 0.5
    dy, ds, dbias = o.scale_bwd(g, y, s, -1.0, dtype)
    tol = dict(rtol=RTOL, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    report('scale_bwd dy', dy, -g * s, **tol)
    report('scale_bwd dscale', ds, (-g * y.float()).sum(0), rtol=1e-3, atol=1e-3)
    report('scale_bwd dbias', dbias, (-g * s).sum(0), rtol=1e-3, atol=1e-3)
    report('colsum', o.colsum(y), y.float().sum(0), rtol=1e-3, atol=1e-3)
    a, bb = torch.randn(1000, device=dev()), torch.randn(1000, device=dev())
    report('axpby', o.axpby(a, bb, -0.5), a - 0.5 * bb, 1e-6, 1e-6)
    x = torch.randn(1003, device=dev())
    assert torch.equal(o.cast_bf16(x), x.to(torch.bfloat16))


# ---- attention backends (CUDA-core fp32 / mma.sync / tcgen05) must agree ----------------------------------------
@pytest.mark.parametrize('backend', ['simt', 'mma', 'tc', 'tc_smem'])
@pytest.mark.parametrize('kind,code', PATTERNS)
@pytest.mark.parametrize('n,T,fm', [(191, 65, 12), (300, 45, 16), (1280, 257, 32)])
def test_attention_backends_bf16(backend, kind, code, n, T, fm, monkeypatch):
    o = ops()
    torch.manual_seed(9)
    monkeypatch.setenv('DALLE_B200_ATTN', 'tc' if backend.startswith('tc') else backend)
    if backend == 'tc_smem':
        monkeypatch.setenv('DALLE_B200_ATTN_P', 'smem')
    else:
        monkeypatch.delenv('DALLE_B200_ATTN_P', raising=False)
    b, h, dh = 2, 2, 64
    q = _mk((b, h, n, dh), torch.bfloat16, dh ** -0.5)
    k = _mk((b, h, n, dh), torch.bfloat16)
    v = _mk((b, h, n, dh), torch.bfloat16)
    spec = o.AttnSpec(code, causal=True, text_len=T, fmap=fm, kernel_size=5 if kind == 'conv_like' else 0, dilation=1)
    out, lse = o.attn_fwd(spec, q, k, v)
    allow = allowed_mask(kind, n, n, T, fm).to(dev())
    qr, kr, vr = (t.float().detach().requires_grad_() for t in (q, k, v))
    want = _attn_ref(qr, kr, vr, allow)
    report(f'attn_fwd[{backend}] {kind} n={n}', out.view(b, n, h, dh).permute(0, 2, 1, 3), want, 2e-2, 2e-2)
    s = (qr @ kr.transpose(-1, -2)).masked_fill(~allow, float('-inf'))
    report(f'lse[{backend}] {kind} n={n}', lse, torch.logsumexp(s, -1), 1e-2, 1e-2)
    g = _mk((b, n, h * dh), torch.bfloat16)
    want.backward(g.float().view(b, n, h, dh).permute(0, 2, 1, 3))
    dqkv = o.attn_bwd(spec, q, k, v, out, lse, g, None, None, 1.0)
    dq, dk, dv = (dqkv.view(b, n, 3, h, dh)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    for name, got, ref in (('dq', dq, qr.grad), ('dk', dk, kr.grad), ('dv', dv, vr.grad)):
        report(f'attn_bwd[{backend}] {name} {kind} n={n}', got, ref, 3e-2, 3e-2 * float(ref.abs().max()))


@pytest.mark.gpu
def test_token_embedding_gather_and_scatter():
    """dalle_b200_embed_fwd / _bwd against nn.Embedding + cat (dalle_pytorch.py:616-630): bit-exact forward, fp32-atomic
    backward (same sums, different order) within 1e-5."""
    from dalle_pytorch_b200.functional import EmbedTokensFn
    torch.manual_seed(0)
    B, Lt, Li, d = 3, 9, 16, 64
    wt = torch.randn(40, d, device='cuda', requires_grad=True)
    wi = torch.randn(32, d, device='cuda', requires_grad=True)
    text = torch.randint(0, 40, (B, Lt), device='cuda')
    image = torch.randint(0, 32, (B, Li), device='cuda')
    image[0, :4] = image[0, 0]                                   # repeated ids: the scatter must accumulate
    out = EmbedTokensFn.apply(text, image, wt, wi)
    ref = torch.cat((torch.nn.functional.embedding(text, wt), torch.nn.functional.embedding(image, wi)), dim=1)
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    out.backward(g)
    got = (wt.grad.clone(), wi.grad.clone())
    wt.grad = wi.grad = None
    ref.backward(g)
    report('embed dW text', got[0], wt.grad, 1e-5, 1e-5)
    report('embed dW image', got[1], wi.grad, 1e-5, 1e-5)
    # text only (generation prefix)
    out2 = EmbedTokensFn.apply(text, None, wt, wi)
    assert torch.equal(out2, torch.nn.functional.embedding(text, wt))


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam_with_clipping():
    """dalle_b200_sumsq + dalle_b200_adam over flat buffers vs clip_grad_norm_ + torch.optim.Adam (train_dalle.py:617-619),
    three steps, with and without weight decay / clipping; sizes that are not multiples of 4 exercise the tails."""
    from dalle_pytorch_b200 import ops
    for n, max_norm, wd in ((1003, 0.5, 0.0), (4096 * 33 + 1, 0.0, 0.01), (7, 2.0, 0.0)):
        torch.manual_seed(n)
        p0 = torch.randn(n, device='cuda')
        ref = torch.nn.Parameter(p0.clone())
        opt = torch.optim.Adam([ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        p, m, v = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
        gn = torch.zeros(1, device='cuda')
        for step in range(1, 4):
            g = torch.randn(n, device='cuda') * (3.0 if step == 2 else 0.01)      # one step clips hard, the others do not
            ref.grad = g.clone()
            if max_norm > 0:
                torch.nn.utils.clip_grad_norm_([ref], max_norm)
            opt.step()
            gn.zero_()
            ops.sumsq_(g, gn)
            report(f'sumsq n={n} step {step}', gn, (g.double() ** 2).sum().float().reshape(1), 1e-5, 1e-6)
            ops.adam_(p, g, m, v, step, 3e-4, 0.9, 0.999, 1e-8, wd, max_norm, gn if max_norm > 0 else None)
            report(f'adam n={n} step {step}', p, ref.detach(), 1e-5, 1e-6)


@pytest.mark.gpu
def test_fused_adam_trains_the_model_like_torch_adam():
    """FusedAdam (flat parameters + flat gradients, weight-gradient GEMMs writing straight into the flat buffer) against
    clip_grad_norm_ + torch.optim.Adam on the same tiny DALLE for three steps."""
    import copy
    import dalle_pytorch_b200 as D
    torch.manual_seed(5)
    kw = dict(dim=64, num_text_tokens=50, text_seq_len=8, depth=2, heads=2, dim_head=64, attn_types=('full', 'axial_row'))
    a = D.DALLE(vae=D.TokenVAE(image_size=32, num_layers=3, num_tokens=32), **kw).cuda().train()
    b = copy.deepcopy(a)
    text = torch.randint(1, 50, (2, 8)).cuda()
    image = torch.randint(0, 32, (2, 16)).cuda()
    opt_a = torch.optim.Adam(a.parameters(), lr=3e-4)
    opt_b = D.FusedAdam(b.parameters(), lr=3e-4, max_grad_norm=0.5)
    for _ in range(3):
        opt_a.zero_grad(set_to_none=True)
        la = a(text, image, return_loss=True)
        la.backward()
        torch.nn.utils.clip_grad_norm_(a.parameters(), 0.5)
        opt_a.step()
        lb = b(text, image, return_loss=True)
        lb.backward()
        opt_b.step()
        report('loss', lb.detach(), la.detach(), 1e-4, 1e-5)
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        report(f'param {k}', pb.detach(), pa.detach(), 1e-3, 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_ln_shift_bwd_fused_upstream_scale_adjoint(dtype):
    """ln_shift_bwd with the upstream LayerScale adjoint fused (db200_ln_shift_bwd_params::up_*) against the two separate
    launches it replaces (ln_shift_bwd then scale_bwd on its dx); rows not a multiple of the 4-row stage."""
    from dalle_pytorch_b200 import ops
    torch.manual_seed(11)
    b, n, d, T, fm = 2, 37, 1024, 8, 6
    x = torch.randn(b, n, d, device='cuda')
    da = torch.randn(b, n, d, device='cuda').to(dtype)
    dres = torch.randn(b, n, d, device='cuda')
    gamma = torch.randn(d, device='cuda')
    mean, rstd = x.mean(-1).reshape(-1).contiguous(), (x.var(-1, unbiased=False) + 1e-5).rsqrt().reshape(-1).contiguous()
    y_up = torch.randn(b * n, d, device='cuda').to(dtype)
    sc_up = torch.randn(d, device='cuda') * 0.1
    dg0, db0 = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
    dx0 = ops.ln_shift_bwd(da, x, mean, rstd, gamma, dres, T, fm, do_ln=True, do_shift=True, dgamma=dg0, dbeta=db0)
    dy0, dsc0, dbi0 = ops.scale_bwd(dx0.view(b * n, d), y_up, sc_up, -1.0, dtype)
    dg1, db1 = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
    dx1, dy1, dsc1, dbi1 = ops.ln_shift_bwd(da, x, mean, rstd, gamma, dres, T, fm, do_ln=True, do_shift=True, dgamma=dg1, dbeta=db1,
                                            up=(y_up, sc_up, -1.0))
    assert torch.equal(dx1, dx0)
    assert torch.equal(dy1, dy0)
    tol = 1e-4 if dtype == torch.float32 else 2e-3
    report('fused dscale', dsc1, dsc0, tol, tol * float(dsc0.abs().max()))
    report('fused dbias', dbi1, dbi0, tol, tol * float(dbi0.abs().max()))
    report('dgamma', dg1, dg0, 1e-4, 1e-4 * float(dg0.abs().max()))


# ---- gathered axial attention (segment tiling + strided TMA boxes) ------------------------------------------------
def _gather_stat_index(T, fm, n, col):
    """token position -> index into the [b, h, n_stat] statistics arrays of the gathered kernels (attn_common.cuh)."""
    t_pad = (T + 63) // 64 * 64
    idx = torch.arange(n)
    img = idx - T
    r, c = img // fm, img % fm
    virt = torch.where(idx < T, idx, t_pad + (c * fm + r if col else img))
    return virt


@pytest.mark.parametrize('kind,code', [('axial_row', 1), ('axial_col', 2)])
@pytest.mark.parametrize('T,fm,missing', [(21, 16, 1), (64, 16, 1), (65, 16, 0), (257, 32, 1), (130, 64, 1)])
def test_axial_gather_kernels(kind, code, T, fm, missing):
    """Gathered axial kernels (db200_attn_fwd_params::gather) against a torch fp32 evaluation, rotary adjoint and q scale
    included, and bit-for-bit layout checks of what the library promises about padding rows."""
    o = ops()
    torch.manual_seed(21)
    b, h, dh = 2, 3, 64
    n = T + fm * fm - missing
    spec = o.AttnSpec(code, causal=True, text_len=T, fmap=fm)
    lay = o.gather_layout(spec, torch.bfloat16, n)
    assert lay is not None and lay.n_alloc == T + fm * fm and lay.col == (code == 2)
    inner = h * dh
    ang = torch.randn(n, 60) * 2.0
    ang[:, 1::2] = ang[:, 0::2]                                  # pair-repeated angles (transformer.py:304-328)
    from dalle_pytorch_b200.attention import rotary_tables
    cos_t, sin_t = rotary_tables(ang.to(dev()), dh)
    scale = dh ** -0.5
    raw = _mk((b * n, 3 * inner), torch.bfloat16, 1.0)           # the to_qkv output
    qkv = torch.full((3, b, h, lay.n_alloc, dh), float('nan'), device=dev(), dtype=torch.bfloat16)
    from dalle_pytorch_b200 import _lib
    import ctypes
    _lib.check(_lib.lib().dalle_b200_qkv_rotary(o._p(raw), o._p(qkv[0]), o._p(qkv[1]), o._p(qkv[2]), o._p(cos_t), o._p(sin_t), 1, b * n, n, h, dh, 0,
                                                scale, lay.n_alloc, o._stream()), 'qkv_rotary')
    q, k, v = qkv[0], qkv[1], qkv[2]
    assert torch.isfinite(qkv.float()).all() and (qkv[:, :, :, n:] == 0).all(), 'rows behind the sequence must be written as zeros'
    # reference: fp32 evaluation from the same bf16 operands
    x = raw.float().view(b, n, 3, h, dh).permute(2, 0, 3, 1, 4).detach().requires_grad_()     # [3, b, h, n, dh]
    rot = apply_rotary(ang.to(dev()), x)
    qr, kr, vr = rot[0] * scale, rot[1], rot[2]
    report('q layout', q[:, :, :n], qr, 2e-2, 2e-2)
    allow = allowed_mask(kind, n, n, T, fm).to(dev())
    want = _attn_ref(q[:, :, :n].float(), k[:, :, :n].float(), v[:, :, :n].float(), allow)
    out, lse = o.attn_fwd(spec, q, k, v, lay=lay)
    assert out.shape == (b, n, inner) and lse.shape == (b, h, lay.n_stat)
    report(f'gather fwd {kind}', out.view(b, n, h, dh).permute(0, 2, 1, 3), want, 2e-2, 2e-2)
    s = (q[:, :, :n].float() @ k[:, :, :n].float().transpose(-1, -2)).masked_fill(~allow, float('-inf'))
    sidx = _gather_stat_index(T, fm, n, lay.col).to(dev())
    report(f'gather lse {kind}', lse[:, :, sidx], torch.logsumexp(s, -1), 1e-2, 1e-2)
    assert torch.isfinite(lse).all(), 'padding entries of lse must be finite (they are bulk-copied by the backward kernels)'
    # the dense-tile predicate kernels on the same operands must agree
    o2, lse2 = o.attn_fwd(spec, q[:, :, :n].contiguous(), k[:, :, :n].contiguous(), v[:, :, :n].contiguous())
    report(f'gather == predicate fwd {kind}', out, o2, 1e-2, 1e-2)
    # backward
    g = _mk((b, n, inner), torch.bfloat16)
    gbuf = o.attn_dout_buffer(lay, b * n, inner, dev(), torch.bfloat16)
    gbuf.copy_(g.view(b * n, inner))
    dqkv = o.attn_bwd(spec, q, k, v, out, lse, gbuf.view(b, n, inner), cos_t, sin_t, scale, lay=lay)
    # reference gradient w.r.t. the to_qkv output through rotary, scale and attention
    qa, ka, va = rot[0] * scale, rot[1], rot[2]
    wa = _attn_ref(qa, ka, va, allow)
    wa.backward(g.float().view(b, n, h, dh).permute(0, 2, 1, 3))
    ref = x.grad.permute(1, 3, 0, 2, 4).reshape(b * n, 3 * inner)                     # [b*n, 3*inner]
    scale_ref = float(ref.abs().max())
    report(f'gather bwd {kind}', dqkv, ref, 3e-2, 3e-2 * scale_ref)
    d2 = o.attn_bwd(spec, q[:, :, :n].contiguous(), k[:, :, :n].contiguous(), v[:, :, :n].contiguous(), o2, lse2, g, cos_t, sin_t, scale)
    report(f'gather == predicate bwd {kind}', dqkv, d2, 1e-2, 1e-2 * scale_ref)


# ---- fp32 GEMMs on the tensor cores (bf16x6) ----------------------------------------------------------------------
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize('M,N,K', [(130, 320, 264), (1280, 3072, 1024), (1024, 1024, 2560)])
def test_gemm_bf16x6_matches_fp64(a_mn, b_mn, M, N, K):
    """The parity mode's GEMM (three-way bf16 split of both fp32 operands, six partial products accumulated in fp32 by ONE
    launch of gemm_tcgen05_kernel over K' = 6K) against a float64 product.  Measured on B200: max error / max|C| = 3e-6 (K = 264),
    1.3e-5 (K = 1024), 2.5e-5 (K = 2560) -- it grows linearly with K because the tensor core's fp32 accumulator truncates at
    every K = 16 step, where the CUDA-core FFMA kernel (round-to-nearest) stays at 4e-7 .. 2e-6.  Two to three orders of
    magnitude below a bf16 (4e-3) or tf32 (5e-4) product, and inside the rtol 1e-3 / atol 1e-5 contract on every golden."""
    import dalle_pytorch_b200 as D
    o = ops()
    torch.manual_seed(31)
    A = torch.randn(K, M, device=dev()) if a_mn else torch.randn(M, K, device=dev())
    B = torch.randn(K, N, device=dev()) if b_mn else torch.randn(N, K, device=dev())
    bias = torch.randn(N, device=dev()) if not a_mn else None
    Am = (A.t() if a_mn else A).double()
    Bm = (B.t() if b_mn else B).double()
    want = Am @ Bm.t() + (bias.double() if bias is not None else 0)
    o.gemm_timing(True)
    with D.fp32_gemm_ctx('bf16x6'):
        got6 = o.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias)
    st = o.gemm_timing(False)
    if a_mn and M % 8:          # M-major A needs M % 8 == 0 on the tcgen05 kernel: the parity mode falls back to the FFMA kernel
        assert st['simt']['launches'] == 1
        return
    assert st['tcgen05']['launches'] == 1 and st['simt']['launches'] == 0, st
    with D.fp32_gemm_ctx('simt'):
        got1 = o.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias)
    scale = float(want.abs().max())
    e6 = float((got6.double() - want).abs().max()) / scale
    e1 = float((got1.double() - want).abs().max()) / scale
    assert got6.dtype == torch.float32 and e6 < 4e-6 + 1.5e-8 * K, (e6, e1)
    assert e1 < 5e-6, (e6, e1)


# ---- dropout (Philox, counter-based) --------------------------------------------------------------------------------
def test_dropout_kernel_mask_properties():
    o = ops()
    n, p = 1 << 20, 0.3
    ones = torch.ones(n, device=dev())
    a = o.dropout_(ones.clone(), p, 1234, 10)
    kept = a != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 5e-3
    assert torch.allclose(a[kept], torch.full_like(a[kept], 1 / (1 - p)))
    assert torch.equal(a, o.dropout_(ones.clone(), p, 1234, 10))                       # a pure function of (seed, offset, i)
    assert not torch.equal(a, o.dropout_(ones.clone(), p, 1235, 10))
    b = o.dropout_(ones.clone(), p, 1234, 10 + 7)                                      # counter-based: offset k = shift by 4k elements
    assert torch.equal(a[28:], b[:-28])
    c = o.dropout_(torch.ones(n, device=dev(), dtype=torch.bfloat16), p, 1234, 10)     # same mask in every dtype
    assert torch.equal(c != 0, kept)
    assert torch.equal(o.dropout_(ones.clone(), 0.0, 1, 0), ones) and not o.dropout_(ones.clone(), 1.0, 1, 0).any()
    x = torch.randn(1003, device=dev())                                                # ragged tail
    y = o.dropout_(x.clone(), 0.5, 9, 0)
    assert torch.equal((y != 0), (o.dropout_(torch.ones_like(x), 0.5, 9, 0) != 0)) and torch.allclose(y[y != 0], 2 * x[y != 0])


# ---- decoding: top-k + Gumbel-max sampling kernel, in-place KV cache ------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('V,thres', [(18448, 0.5), (18448, 0.9), (1000, 0.0), (90, 0.99)])
def test_sample_topk_gumbel_matches_torch(dtype, V, thres):
    """dalle_b200_sample_topk_gumbel against the reference's top_k + gumbel_sample (dalle_pytorch.py:43-58) fed with the same
    Gumbel noise; half of the row is masked to -fp32max like the logits of DALLE.forward (:648-652)."""
    o = ops()
    torch.manual_seed(41)
    B = 16
    logits = (torch.randn(B, V, device=dev()) * 3).to(dtype)
    logits[:, : V // 2] = torch.finfo(torch.float32).min if dtype == torch.float32 else torch.finfo(torch.bfloat16).min
    noise = -torch.log(-torch.log(torch.rand(B, V, device=dev()).clamp_min(1e-20)))
    k = max(int((1 - thres) * V), 1)
    val, ind = torch.topk(logits.float(), k)
    filt = torch.full_like(logits.float(), float('-inf')).scatter_(1, ind, val)
    for temp in (1.0, 0.7):
        want = (filt / temp + noise).argmax(-1)
        got = o.sample_topk_gumbel(logits, thres, temp, gumbel=noise)
        assert torch.equal(got, want), (got, want)
    # internal Philox noise: deterministic in (seed, offset), inside the top-k set, and distributed like softmax over it
    a = o.sample_topk_gumbel(logits, thres, 1.0, seed=5, offset=0)
    assert torch.equal(a, o.sample_topk_gumbel(logits, thres, 1.0, seed=5, offset=0))
    assert (filt.gather(1, a[:, None]) > float('-inf')).all()
    if V == 90:
        row = logits[:1].float().repeat(4096, 1).contiguous().to(dtype)
        s = o.sample_topk_gumbel(row, 0.9, 1.0, seed=7, offset=123)
        kk = max(int(0.1 * V), 1)
        tv, ti = torch.topk(row[0].float(), kk)
        p = torch.softmax(tv, 0)
        freq = torch.stack([(s == i).float().mean() for i in ti])
        assert (freq - p).abs().max() < 0.04, (freq, p)


def test_inplace_kv_cache_attention():
    """Forward attention over an in-place KV cache (kv_rows > n_k) == attention over the compact tensors, all backends."""
    o = ops()
    torch.manual_seed(43)
    b, h, dh, cap = 2, 3, 64, 300
    for dtype in (torch.float32, torch.bfloat16):
        for n_k, n_q in ((1, 1), (130, 1), (257, 3)):
            kbuf = _mk((b, h, cap, dh), dtype)
            vbuf = _mk((b, h, cap, dh), dtype)
            kbuf[:, :, n_k:] = 3.0e4                             # rows behind the valid ones must not influence the result (they must be
            vbuf[:, :, n_k:] = -3.0e4                            # FINITE: the tensor-core kernels multiply them by exact zeros)
            q = _mk((b, h, n_q, dh), dtype, dh ** -0.5)
            spec = o.AttnSpec(0, causal=True)
            a, _ = o.attn_fwd(spec, q, kbuf, vbuf, n_k=n_k)
            w, _ = o.attn_fwd(spec, q, kbuf[:, :, :n_k].contiguous(), vbuf[:, :, :n_k].contiguous())
            assert torch.isfinite(a.float()).all()
            report(f'in-place cache n_k={n_k} {dtype}', a, w, 1e-6, 1e-6)


# ---- decoding kernels (csrc/decode.cu): position read from device memory ----------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_decode_shift_and_kv_append_kernels(dtype):
    """dalle_b200_decode_shift == the cache branch of PreShiftToken (transformer.py:155-170) written with torch indexing, over a run of
    consecutive positions (ring wrap-around, first-of-row zeroing); dalle_b200_decode_kv_append == a row assignment."""
    o = ops()
    torch.manual_seed(61)
    for b, d, fm, text_len in ((3, 128, 4, 9), (2, 1024, 32, 257), (1, 64, 1, 5), (5, 40, 3, 2)):
        q, half = d // 4, d // 2
        top = torch.randn(fm, b, q, device='cuda')
        left = torch.randn(fm, b, half - q, device='cuda')
        top_w, left_w = top.clone(), left.clone()
        pos_t = torch.tensor([text_len + 2], device='cuda', dtype=torch.int64)
        for step in range(2 * fm + 3):
            h = torch.randn(b, d, device='cuda')
            pos = int(pos_t)
            slot = (pos - text_len) % fm
            prev = (slot + fm - 1) % fm
            want = h.clone()
            want[:, :q] = top_w[slot]
            want[:, q:half] = left_w[prev] if slot != 0 else 0.
            top_w[slot] = h[:, :q]
            left_w[slot] = h[:, q:half]
            got = o.decode_shift(h, top, left, pos_t, text_len, fm, dtype)
            assert got.dtype == dtype and torch.equal(got, want.to(dtype)), (b, d, fm, step)
            assert torch.equal(top, top_w) and torch.equal(left, left_w), (b, d, fm, step)
            pos_t.add_(1)
    bsz, h_, dh, cap = 2, 3, 64, 37
    kc, vc = _mk((bsz, h_, cap, dh), dtype), _mk((bsz, h_, cap, dh), dtype)
    kw, vw = kc.clone(), vc.clone()
    for pos in (0, 5, 36):
        kn, vn = _mk((bsz, h_, 1, dh), dtype), _mk((bsz, h_, 1, dh), dtype)
        o.decode_kv_append(kn, vn, kc, vc, torch.tensor([pos], device='cuda', dtype=torch.int64))
        kw[:, :, pos], vw[:, :, pos] = kn[:, :, 0], vn[:, :, 0]
        assert torch.equal(kc, kw) and torch.equal(vc, vw)


@pytest.mark.parametrize('n_k,cap', [(1, 1), (7, 40), (64, 64), (513, 600), (1281, 1281), (1500, 1536)])
def test_single_query_attention_kernel(n_k, cap, monkeypatch):
    """attn_decode_kernel (one query per head, bf16): against the fp32 softmax of the same operands for the causal, key-masked,
    static-mask and axial patterns, a fully masked row (uniform attention, as the reference's -fp32max fill gives), and against the
    general tensor-core kernel (DALLE_B200_DECODE_ATTN=0)."""
    o = ops()
    torch.manual_seed(62 + n_k)
    b, h, dh = 2, 3, 64
    q = _mk((b, h, 1, dh), torch.bfloat16, dh ** -0.5 * 3)
    kbuf, vbuf = _mk((b, h, cap, dh), torch.bfloat16), _mk((b, h, cap, dh), torch.bfloat16)
    kbuf[:, :, n_k:] = 3.0e4
    vbuf[:, :, n_k:] = -3.0e4
    s = torch.einsum('bhqd,bhkd->bhqk', q.float(), kbuf[:, :, :n_k].float())[:, :, 0]            # [b, h, n_k]
    neg = -torch.finfo(torch.float32).max

    def want(allowed):                                                                      # allowed: bool [b, n_k]
        p = s.masked_fill(~allowed[:, None, :], neg).softmax(dim=-1)
        return torch.einsum('bhk,bhkd->bhd', p, vbuf[:, :, :n_k].float()).reshape(b, 1, h * dh)

    all_keys = torch.ones(b, n_k, dtype=torch.bool, device='cuda')
    cases = [('causal', o.AttnSpec(0, causal=True), None, all_keys)]
    km = torch.rand(b, n_k, device='cuda') > 0.4
    km[:, 0] = True
    cases.append(('key mask, non-causal', o.AttnSpec(0, causal=False), km.to(torch.uint8).contiguous(), km))
    km0 = km.clone()
    km0[1] = False                                                                          # batch 1: no key allowed -> uniform
    cases.append(('fully masked row', o.AttnSpec(0, causal=False), km0.to(torch.uint8).contiguous(), km0))
    sm = torch.rand(n_k, n_k, device='cuda') > 0.5
    sm[:, 0] = True
    cases.append(('static mask', o.AttnSpec(4, causal=True, static_mask=sm.to(torch.uint8).contiguous()), None, sm[n_k - 1][None].expand(b, -1)))
    T, fm = 5, 6
    if n_k > T + 1:
        j = torch.arange(n_k, device='cuda')
        qi = n_k - 1 - T
        row_ok = (j < T) | (((j - T) // fm == qi // fm) & (j <= n_k - 1))
        cases.append(('axial row predicate', o.AttnSpec(1, causal=True, text_len=T, fmap=fm), None, row_ok[None].expand(b, -1)))
    for name, spec, mask, allowed in cases:
        monkeypatch.setenv('DALLE_B200_DECODE_ATTN', '1')
        got, lse = o.attn_fwd(spec, q, kbuf, vbuf, mask, n_k=n_k)
        assert got.shape == (b, 1, h * dh) and torch.isfinite(got.float()).all()
        w = want(allowed)
        if name == 'fully masked row':
            w[1] = vbuf[1, :, :n_k].float().mean(dim=1).reshape(1, h * dh)
        report(f'single-query attention n_k={n_k} {name}', got, w, 2e-2, 4e-3)
        w_lse = s.masked_fill(~allowed[:, None, :], neg).logsumexp(dim=-1)
        if name != 'fully masked row':
            report(f'single-query lse n_k={n_k} {name}', lse[:, :, 0], w_lse, 1e-3, 1e-3)
        if name in ('causal', 'key mask, non-causal', 'static mask'):                       # the combinations the decoding paths use
            monkeypatch.setenv('DALLE_B200_DECODE_ATTN', '0')
            ref, _ = o.attn_fwd(spec, q, kbuf, vbuf, mask, n_k=n_k)
            report(f'single-query vs general kernel n_k={n_k} {name}', got, ref, 2e-2, 6e-3)


@pytest.mark.parametrize('M', [1, 3, 16])
@pytest.mark.parametrize('N,K', [(1024, 1024), (3072, 1024), (1024, 4096), (64, 256), (48, 768)])
def test_small_m_gemm_kernel(M, N, K):
    """gemm_smallm_kernel (mma.sync weight streaming for M <= 16, csrc/gemm_smallm.cu) through the same ops entry points the decoding
    step uses: STORE (+bias, bf16 / fp32 result), RESID (bias, residual, LayerScale vector, sign, kept y), GEGLU (kept u), against
    torch's fp32 product of the same bf16 operands -- and the launches really went to the small-M backend."""
    o = ops()
    torch.manual_seed(71 + M + N)
    A, W = _mk((M, K), torch.bfloat16), _mk((N, K), torch.bfloat16, K ** -0.5)
    bias = _mk((N,), torch.float32)
    acc = A.float() @ W.float().t()
    o.gemm_timing(True)
    got = o.gemm_store(A, W)
    got_b = o.gemm_store(A, W, bias=bias, out_dtype=torch.float32)
    resid, scale = _mk((M, N), torch.float32), _mk((N,), torch.float32)
    out, y = o.gemm_resid(A, W, bias, resid, scale, sign=-1.0, keep_y=True)
    out2, y2 = o.gemm_resid(A, W, bias, None, None, 1.0)
    H = N // 2
    h, u = o.gemm_geglu(A, W, bias, keep_u=True)
    h2, u2 = o.gemm_geglu(A, W, bias, keep_u=False)
    st = o.gemm_timing(False)
    assert st['smallm']['launches'] == 6 and st['tcgen05']['launches'] == 0 and st['simt']['launches'] == 0, st
    report(f'small-M store {M}x{N}x{K}', got, acc, 1e-2, 1e-2)
    report(f'small-M store+bias fp32 {M}x{N}x{K}', got_b, acc + bias, 1e-4, 1e-4)
    report(f'small-M resid {M}x{N}x{K}', out, resid - scale * (acc + bias), 1e-4, 1e-4)
    report(f'small-M resid y {M}x{N}x{K}', y, acc + bias, 1e-2, 1e-2)
    report(f'small-M plain projection {M}x{N}x{K}', out2, acc + bias, 1e-4, 1e-4)
    assert y2 is None and u2 is None
    ub = acc + bias
    report(f'small-M geglu u {M}x{N}x{K}', u, ub, 1e-2, 1e-2)
    want_h = ub[:, :H] * torch.nn.functional.gelu(ub[:, H:])
    report(f'small-M geglu h {M}x{N}x{K}', h, want_h, 1e-2, 1e-2)
    assert torch.equal(h, h2)


# ---- tcgen05 GEMM at benchmark scale (persistent multi-wave grid, dynamic tile scheduler, split-K, all operand majors) -----------
@pytest.mark.parametrize('name,M,N,K,a_mn,b_mn,out_f32', [
    ('fwd FF1-like (store)', 20480, 4096, 1024, False, False, False),
    ('dgrad FF1 (B read N-major)', 20480, 1024, 8192, False, True, False),
    ('wgrad FF2 (both MN-major, fp32 out)', 1024, 4096, 20480, True, True, True),
    ('wgrad out-proj (split-K + fp32 atomics)', 1024, 1024, 20480, True, True, True),
    ('ragged M / N tails', 20000, 3000, 1032, False, False, False),
])
def test_gemm_tcgen05_at_benchmark_scale(name, M, N, K, a_mn, b_mn, out_f32):
    """The shapes of a C2 step (M = 16 x 1280 rows) through gemm_tcgen05_kernel against torch's fp32 matmul of the same bf16
    operands: thousands of tiles per launch (multi-wave persistent grid + atomic tile scheduler), the MN-major operand paths and
    the split-K reduction, which the small-shape tests above do not reach."""
    o = ops()
    torch.manual_seed(51)
    A = _mk((K, M) if a_mn else (M, K), torch.bfloat16)
    B = _mk((K, N) if b_mn else (N, K), torch.bfloat16)
    o.gemm_timing(True)
    got = o.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32 if out_f32 else None)
    st = o.gemm_timing(False)
    assert st['tcgen05']['launches'] == 1 and st['simt']['launches'] == 0, st
    want = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    scale = float(want.abs().max())
    # fp32 result: the tensor core's truncating fp32 accumulator over K / 16 steps (measured 3e-5 at K = 20480); bf16 result: one rounding
    tol = (2e-5 + 2e-9 * K) * scale if out_f32 else 8e-3 * scale
    report(f'gemm {name}', got, want, 0.0, tol)
