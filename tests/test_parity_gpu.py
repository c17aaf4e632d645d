"""Parity of the CUDA path (through the module API -> C ABI) with the reference, on the B200.

 * golden fixtures generated from the UNMODIFIED reference (tests/golden, oracle/make_golden.py): logits, loss and
   every parameter gradient, fp32 mode, tolerance rtol 1e-3 / atol 1e-5 (BASELINE.json north_star); masked logits
   must be exactly -fp32max;
 * the CPU oracle on fresh seeded inputs, sub-layer by sub-layer (localises a failure);
 * bf16 speed mode against the same goldens with a stated loose tolerance (it is the same computation).
"""
import pytest
import torch

from conftest import load_golden, golden_names
from util import report
from dalle_oracle import (OracleConfig, make_state_dict, make_inputs, dalle_forward, layer_params, attn_sublayer, ff_sublayer,
                          rotary_angle_table, allowed_mask, transformer_forward)

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-5


def _cfg(rec):
    c = dict(rec['cfg'])
    c['attn_types'] = tuple(c['attn_types'])
    return OracleConfig(**c)


def build(cfg, sd, device='cuda:0'):
    import dalle_pytorch_b200 as D
    vae = D.TokenVAE(image_size=8 * cfg.fmap, num_layers=3, num_tokens=cfg.num_image_tokens)
    m = D.DALLE(dim=cfg.dim, vae=vae, num_text_tokens=cfg.num_text_tokens, text_seq_len=cfg.text_seq_len, depth=cfg.depth,
                heads=cfg.heads, dim_head=cfg.dim_head, reversible=cfg.reversible, attn_types=cfg.attn_types, stable=cfg.stable,
                sandwich_norm=cfg.sandwich_norm, shift_tokens=cfg.shift_tokens, loss_img_weight=cfg.loss_img_weight)
    m.load_state_dict(sd)
    return m.to(device)


TIED = {'tiny_shared', 'tiny_tied_emb'}      # fixtures that carry the reference's own (tied) state dict


def run_model(rec, dtype):
    import dalle_pytorch_b200 as D
    cfg = _cfg(rec)
    if 'state' in rec:
        vae = D.TokenVAE(image_size=8 * cfg.fmap, num_layers=3, num_tokens=cfg.num_image_tokens)
        m = D.DALLE(dim=cfg.dim, vae=vae, num_text_tokens=cfg.num_text_tokens, text_seq_len=cfg.text_seq_len, depth=cfg.depth,
                    heads=cfg.heads, dim_head=cfg.dim_head, reversible=cfg.reversible, attn_types=cfg.attn_types, stable=cfg.stable,
                    sandwich_norm=cfg.sandwich_norm, shift_tokens=cfg.shift_tokens, loss_img_weight=cfg.loss_img_weight,
                    shared_attn_ids=cfg.shared_attn_ids, shared_ff_ids=cfg.shared_ff_ids, **rec.get('extra', {}))
        m.load_state_dict(rec['state'])            # strict: the tied reference state dict must fit key for key
        m = m.cuda()
    else:
        sd = make_state_dict(cfg, seed=rec['seed'])
        m = build(cfg, sd)
    m.train()
    text, image = rec['text'].cuda(), rec['image'].cuda()
    with D.compute_dtype_ctx(dtype):
        loss = m(text, image, return_loss=True)
        loss.backward()
        with torch.no_grad():
            logits = m(text, image)
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    return loss.detach(), logits, grads


@pytest.fixture(params=['bf16x6', 'simt'])
def fp32_gemm(request):
    """Both evaluations of the fp32 parity mode's GEMMs: 'bf16x6' = the tcgen05 kernel of the speed path fed with three-way
    bf16 splits of the fp32 operands (the default), 'simt' = the CUDA-core FFMA kernel (cross-check)."""
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import ops
    with D.fp32_gemm_ctx(request.param):
        ops.gemm_timing(True)
        yield request.param
        st = ops.gemm_timing(False)
    if request.param == 'bf16x6':
        assert st['tcgen05']['launches'] > 0, 'parity mode did not reach the tcgen05 kernel'


@pytest.mark.parametrize('name', golden_names('tiny_'))
def test_tiny_goldens_fp32(name, fp32_gemm):
    """Includes `tiny_shared` (shared_attn_ids / shared_ff_ids, transformer.py:261-292), `tiny_tied_emb`
    (share_input_output_emb, dalle_pytorch.py:432-443) and `tiny_axial_stable` (stable softmax on the axial classes)."""
    rec = load_golden(name)
    loss, logits, grads = run_model(rec, torch.float32)
    report('loss', loss, rec['loss'], RTOL, ATOL)
    lg = logits.cpu()
    masked = rec['logits'] < -1e30
    assert torch.equal(lg[masked], rec['logits'][masked]), 'masked logits must be exactly -fp32max'
    report('logits', lg[~masked], rec['logits'][~masked], RTOL, ATOL)
    assert set(grads) == set(rec['grads']), set(grads) ^ set(rec['grads'])
    for k, g in rec['grads'].items():
        report(f'grad {k}', grads[k], g, RTOL, ATOL)


@pytest.mark.parametrize('name', golden_names('c1_') + ['c3_geom', 'c4_geom_rev'])
def test_c1_goldens_fp32(name, fp32_gemm):
    """c1_*: BASELINE.json configs[0] (depth 2, dim 256, heads 4, text 64, image 8x8, batch 2).  c3_geom / c4_geom_rev: the
    geometry of configs[2..4] (dim 1024, heads 16, text 256, image 32x32, full vocabulary; axial row+column, sequential and
    reversible executors) at depth 2, batch 1 -- sampled logits, log-sum-exp of every row, sampled gradients and gradient norms
    of every parameter from the unmodified reference."""
    rec = load_golden(name)
    loss, logits, grads = run_model(rec, torch.float32)
    report('loss', loss, rec['loss'], RTOL, ATOL)
    lg = logits.cpu()
    samp = lg[..., ::rec['logits_stride']]
    masked = rec['logits_sample'] < -1e30
    assert torch.equal(samp[masked], rec['logits_sample'][masked])
    report('logits sample', samp[~masked], rec['logits_sample'][~masked], RTOL, ATOL)
    report('logits lse', torch.logsumexp(lg.double(), -1).float(), rec['logits_lse'], RTOL, 1e-4)
    for k, (vals, step) in rec['grad_samples'].items():
        mine = grads[k].reshape(-1)[::step][:vals.numel()]
        report(f'grad {k}', mine, vals, RTOL, ATOL)
        gn = float(grads[k].double().norm())
        assert abs(gn - rec['grad_norms'][k]) <= 1e-3 * rec['grad_norms'][k] + 1e-7, (k, gn, rec['grad_norms'][k])


@pytest.mark.parametrize('name', ['tiny_full', 'tiny_axial', 'tiny_axial_rev', 'tiny_cycle4', 'tiny_shared', 'c1_full', 'c1_axial_rev',
                                  'c3_geom', 'c4_geom_rev'])
def test_goldens_bf16_mode(name):
    """Speed mode (bf16 storage, tensor-core GEMMs) computes the same function: loss within 2e-2, logits within
    6e-2 absolute (bf16 has 8 mantissa bits; SURVEY.md App. C measured 1.5e-2 for a bf16 evaluation of the reference)."""
    rec = load_golden(name)
    loss, logits, grads = run_model(rec, torch.bfloat16)
    report('loss', loss, rec['loss'], 2e-2, 2e-2)
    lg = logits.cpu()
    if 'logits' in rec:
        want, got = rec['logits'], lg
    else:
        want, got = rec['logits_sample'], lg[..., ::rec['logits_stride']]
    masked = want < -1e30
    assert torch.equal(got[masked], want[masked])
    report('logits', got[~masked], want[~masked], 0.0, 6e-2)
    for k, g in grads.items():
        assert torch.isfinite(g).all(), k
    if 'grads' in rec:
        for k, g in rec['grads'].items():
            scale = float(g.abs().max())
            report(f'grad {k}', grads[k], g, 0.0, 0.06 * scale + 1e-6)


@pytest.mark.parametrize('kind', ['full', 'axial_row', 'axial_col', 'conv_like'])
@pytest.mark.parametrize('reslike', ['sequential', 'separate_resid'])
def test_sublayers_against_oracle(kind, reslike):
    """Fused sub-layer kernels (functional.py) vs the oracle's torch-CPU sub-layer incl. all gradients (fp32)."""
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import ops
    from dalle_pytorch_b200.functional import SublayerGeom, AttnSublayerFn, FFSublayerFn
    from dalle_pytorch_b200.attention import rotary_tables
    torch.manual_seed(11)
    cfg = OracleConfig(dim=128, depth=1, heads=2, text_seq_len=20, fmap=6, num_text_tokens=50, num_image_tokens=32, attn_types=(kind,))
    sd = make_state_dict(cfg, seed=3)
    P = layer_params(sd, cfg, 0)
    b, n, d = 2, cfg.seq_len, cfg.dim
    x = torch.randn(b, n, d)
    r = torch.randn(b, n, d)
    ang = rotary_angle_table(cfg.text_len, cfg.fmap, 64)
    allow = allowed_mask(kind, n, n, cfg.text_len, cfg.fmap)
    Pr = {k: v.clone().requires_grad_() for k, v in P.items()}
    xr, rr = x.clone().requires_grad_(), r.clone().requires_grad_()
    if reslike == 'sequential':
        want = xr + attn_sublayer(xr, Pr, cfg, kind, ang, allow)
    else:
        want = rr - attn_sublayer(xr, Pr, cfg, kind, ang, allow)
    gout = torch.randn(b, n, d)
    want.backward(gout)

    code = {'full': 0, 'axial_row': 1, 'axial_col': 2, 'conv_like': 3}[kind]
    spec = ops.AttnSpec(code, causal=True, text_len=cfg.text_len, fmap=cfg.fmap, kernel_size=5, dilation=1)
    g = SublayerGeom(dtype=torch.float32, text_len=cfg.text_len, fmap=cfg.fmap, do_ln=True, do_shift=True, heads=2, dim_head=64, attn_spec=spec)
    cos_t, sin_t = rotary_tables(ang.cuda(), 64)
    Pg = {k: v.clone().cuda().requires_grad_() for k, v in P.items()}
    xg, rg = x.clone().cuda().requires_grad_(), r.clone().cuda().requires_grad_()
    if reslike == 'sequential':
        got = AttnSublayerFn.apply(g, True, 1.0, cos_t, sin_t, None, xg, None, Pg['a_ln_w'], Pg['a_ln_b'], Pg['w_qkv'], Pg['w_out'], Pg['b_out'], Pg['a_scale'])
    else:
        got = AttnSublayerFn.apply(g, False, -1.0, cos_t, sin_t, None, xg, rg, Pg['a_ln_w'], Pg['a_ln_b'], Pg['w_qkv'], Pg['w_out'], Pg['b_out'], Pg['a_scale'])
    got.backward(gout.cuda())
    report('attn sublayer out', got, want, RTOL, ATOL)
    report('attn dx', xg.grad, xr.grad, RTOL, ATOL)
    if reslike != 'sequential':
        report('attn dresid', rg.grad, rr.grad, RTOL, ATOL)
    for k in ('a_ln_w', 'a_ln_b', 'w_qkv', 'w_out', 'b_out', 'a_scale'):
        report(f'attn d{k}', Pg[k].grad, Pr[k].grad, RTOL, 2e-5)

    # feed-forward
    Pr = {k: v.clone().requires_grad_() for k, v in P.items()}
    xr = x.clone().requires_grad_()
    want = xr + ff_sublayer(xr, Pr, cfg)
    want.backward(gout)
    gf = SublayerGeom(dtype=torch.float32, text_len=cfg.text_len, fmap=cfg.fmap, do_ln=True, do_shift=True)
    Pg = {k: v.clone().cuda().requires_grad_() for k, v in P.items()}
    xg = x.clone().cuda().requires_grad_()
    got = FFSublayerFn.apply(gf, True, 1.0, xg, None, Pg['f_ln_w'], Pg['f_ln_b'], Pg['w1'], Pg['b1'], Pg['w2'], Pg['b2'], Pg['f_scale'])
    got.backward(gout.cuda())
    report('ff sublayer out', got, want, RTOL, ATOL)
    report('ff dx', xg.grad, xr.grad, RTOL, ATOL)
    for k in ('f_ln_w', 'f_ln_b', 'w1', 'b1', 'w2', 'b2', 'f_scale'):
        report(f'ff d{k}', Pg[k].grad, Pr[k].grad, RTOL, 2e-5)


def test_module_api_standalone_and_cache():
    """Attention / SparseAxialCausalAttention / FeedForward used stand-alone (reference signatures), shorter
    sequences (n < seq_len), and KV-cache decoding == full-prefix evaluation (transformer.py:251-260)."""
    import dalle_pytorch_b200 as D
    from dalle_oracle import attention_core, feed_forward
    torch.manual_seed(12)
    dim, heads, fm, tsl = 128, 2, 4, 8
    seq_len = tsl + fm * fm
    T = tsl + 1
    ang = rotary_angle_table(T, fm, 64)
    for n in (seq_len, 19):
        x = torch.randn(2, n, dim)
        for axis, kind in ((0, 'axial_row'), (1, 'axial_col')):
            m = D.SparseAxialCausalAttention(dim, seq_len, image_size=fm, axis=axis, heads=heads).cuda()
            got = m(x.cuda(), rotary_pos_emb=ang[None].cuda())
            want = attention_core(x, m.to_qkv.weight.cpu(), m.to_out[0].weight.cpu(), m.to_out[0].bias.cpu(), heads, ang,
                                  allowed_mask(kind, n, n, T, fm), False)
            report(f'SparseAxialCausalAttention axis={axis} n={n}', got, want, RTOL, ATOL)
    m = D.Attention(dim, seq_len, causal=True, heads=heads, stable=True).cuda()
    x = torch.randn(2, seq_len, dim)
    got = m(x.cuda(), rotary_pos_emb=ang[None].cuda())
    want = attention_core(x, m.to_qkv.weight.cpu(), m.to_out[0].weight.cpu(), m.to_out[0].bias.cpu(), heads, ang,
                          allowed_mask('full', seq_len, seq_len, T, fm), True)
    report('Attention(stable)', got, want, RTOL, ATOL)
    ff = D.FeedForward(dim, mult=4).cuda()
    report('FeedForward', ff(x.cuda()), feed_forward(x, ff.net[0].weight.cpu(), ff.net[0].bias.cpu(), ff.net[3].weight.cpu(), ff.net[3].bias.cpu()),
           RTOL, ATOL)
    # cached decoding of the optimize_for_inference model == uncached logits at every step
    cfg = OracleConfig(dim=64, depth=2, heads=2, text_seq_len=8, fmap=4, num_text_tokens=50, num_image_tokens=32,
                       attn_types=('axial_row', 'axial_col'))
    sd = make_state_dict(cfg, seed=5)
    vae = D.TokenVAE(image_size=32, num_layers=3, num_tokens=32)
    kw = dict(dim=64, vae=vae, num_text_tokens=50, text_seq_len=8, depth=2, heads=2, attn_types=('axial_row', 'axial_col'))
    fast = D.DALLE(optimize_for_inference=True, **kw)
    fast.load_state_dict(sd)
    fast = fast.cuda().eval()
    text, image = make_inputs(cfg, 2, seed=9)
    text, image = text.cuda(), image.cuda()
    with torch.no_grad():
        full = fast(text, image)                                      # [2, 24, V]
        want_cpu = dalle_forward(text.cpu(), image.cpu(), sd, cfg)
        report('optimize_for_inference logits', full.cpu()[want_cpu > -1e30], want_cpu[want_cpu > -1e30], RTOL, ATOL)
        cache = {}
        for cur in range(8, 8 + 6):
            lg = fast(text, image[:, :cur - 8], cache=cache)
            ref = full[:, cur] if cur < 24 else None
            live = ref > -1e30
            report(f'cached step {cur}', lg[:, -1][live], ref[live], RTOL, 2e-5)


def test_c2_geometry_single_layer_fp32_vs_oracle():
    """BASELINE.json configs[1] geometry (dim 1024, heads 16, text 256, image 32x32, n = 1280) at depth 1 / batch 1, fp32
    mode against the CPU oracle: loss, sampled logits and a few gradients at rtol 1e-3 / atol 1e-5."""
    import dalle_pytorch_b200 as D
    cfg = OracleConfig(dim=1024, depth=1, heads=16, text_seq_len=256, fmap=32, num_text_tokens=10000, num_image_tokens=8192,
                       attn_types=('full',))
    sd = make_state_dict(cfg, seed=2, fast=True)
    text, image = make_inputs(cfg, 1, seed=3)
    params = {k: v.clone().requires_grad_(k != 'transformer.pos_emb') for k, v in sd.items()}
    want = dalle_forward(text, image, params, cfg, return_loss=True)
    want.backward()
    with torch.no_grad():
        want_logits = dalle_forward(text, image, sd, cfg)
    m = build(cfg, sd).train()
    with D.compute_dtype_ctx(torch.float32):
        loss = m(text.cuda(), image.cuda(), return_loss=True)
        loss.backward()
        with torch.no_grad():
            logits = m(text.cuda(), image.cuda()).cpu()
    report('loss', loss, want, RTOL, ATOL)
    live = want_logits > -1e30
    assert torch.equal(logits[~live], want_logits[~live])
    report('logits', logits[live][::97], want_logits[live][::97], RTOL, ATOL)
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    for k in ('transformer.layers.layers.0.0.fn.fn.fn.fn.fn.to_qkv.weight', 'transformer.layers.layers.0.1.fn.fn.fn.fn.net.3.weight',
              'transformer.layers.layers.0.0.scale', 'transformer.layers.layers.0.1.fn.norm.weight', 'to_logits.1.bias'):
        report(f'grad {k}', grads[k], params[k].grad, RTOL, ATOL)


def test_c2_geometry_bf16_consistent_with_fp32_mode():
    """Same model, both precision modes of the library at the benchmark geometry (depth 2, batch 2, axial + full)."""
    import dalle_pytorch_b200 as D
    cfg = OracleConfig(dim=1024, depth=2, heads=16, text_seq_len=256, fmap=32, num_text_tokens=10000, num_image_tokens=8192,
                       attn_types=('full', 'axial_col'))
    sd = make_state_dict(cfg, seed=4, fast=True)
    text, image = make_inputs(cfg, 2, seed=5)
    out = {}
    for dt in (torch.float32, torch.bfloat16):
        m = build(cfg, sd).train()
        with D.compute_dtype_ctx(dt):
            loss = m(text.cuda(), image.cuda(), return_loss=True)
            loss.backward()
        out[dt] = (loss.detach(), m.transformer.layers.layers[0][1].fn.fn.fn.fn.net[0].weight.grad.clone())
    report('loss bf16 vs fp32', out[torch.bfloat16][0], out[torch.float32][0], 1e-2, 1e-2)
    g32, g16 = out[torch.float32][1], out[torch.bfloat16][1]
    report('ff1 grad bf16 vs fp32', g16, g32, 0.0, 0.05 * float(g32.abs().max()))


def test_generate_images_cached_and_uncached():
    """generate_images (dalle_pytorch.py:506-562) with and without the KV cache on an optimize_for_inference model: both run
    through the library kernels and, with the same Gumbel noise, produce the same image tokens."""
    import dalle_pytorch_b200 as D
    cfg = OracleConfig(dim=64, depth=2, heads=2, text_seq_len=8, fmap=4, num_text_tokens=50, num_image_tokens=32,
                       attn_types=('axial_row', 'axial_col'))
    sd = make_state_dict(cfg, seed=6)
    vae = D.TokenVAE(image_size=32, num_layers=3, num_tokens=32)
    m = D.DALLE(dim=64, vae=vae, num_text_tokens=50, text_seq_len=8, depth=2, heads=2, attn_types=('axial_row', 'axial_col'),
                optimize_for_inference=True)
    m.load_state_dict(sd)
    m = m.cuda()
    text, _ = make_inputs(cfg, 2, seed=7, pad_tail=False)
    toks = []
    for use_cache in (False, True):
        torch.manual_seed(123)
        out = m.generate_images(text.cuda(), use_cache=use_cache, filter_thres=0.9)
        assert out.shape == (2, 16) and out.min() >= 0 and out.max() < 32
        toks.append(out.cpu())
    assert torch.equal(toks[0], toks[1]), (toks[0], toks[1])          # same noise, same logits -> identical image tokens
    # teacher-forced: the logits of every decoding step with the KV cache equal the last row of the uncached forward
    img = toks[0].cuda()
    cache = {}
    with torch.no_grad():
        for cur in range(0, img.shape[1]):
            step = m(text.cuda(), img[:, :cur], cache=cache)[:, -1]
            full = m(text.cuda(), img[:, :cur])[:, -1]
            live = full > -1e30
            assert torch.equal(live, step > -1e30)
            report(f'cached logits step {cur}', step[live], full[live], 1e-4, 1e-5)


@pytest.mark.parametrize('mode', ['nest', 'flat', 'flat_bucket'])
@pytest.mark.parametrize('variant', ['full_shift_bf16', 'axial_static_fp32', 'full_noshift_stable_fp32'])
def test_generate_images_graph_replay(variant, mode, monkeypatch):
    """decode.GraphedDecoder: after the prompt pass every image token is one CUDA-graph replay whose position is a device tensor
    (rotary row, KV write, allowed keys, token-shift slot selected on the device).  Same seed -> the same image tokens as the
    host-indexed cached loop, and teacher-forced logits equal the host-indexed cached step at every position.
    mode: 'nest' = the captured step walks the module nest; 'flat' = GraphedDecoder walks the layers with the decode kernels
    (dalle_b200_decode_shift / decode_kv_append, LayerScale + residual in the projection epilogue); 'flat_bucket' = additionally one
    graph per 16-key bucket of visible cache rows."""
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import decode
    monkeypatch.setattr(decode, 'FLAT_DEFAULT', mode != 'nest')
    monkeypatch.setattr(decode, 'BUCKET_DEFAULT', 16 if mode == 'flat_bucket' else 0)
    kw = dict(full_shift_bf16=dict(attn_types=('full',), shift_tokens=True),
              axial_static_fp32=dict(attn_types=('axial_row', 'axial_col'), shift_tokens=True, optimize_for_inference=True),
              full_noshift_stable_fp32=dict(attn_types=('full',), shift_tokens=False, stable=True))[variant]
    dtype = torch.bfloat16 if variant.endswith('bf16') else torch.float32
    torch.manual_seed(21)
    vae = D.TokenVAE(image_size=64, num_layers=3, num_tokens=40)           # fmap 8: 64 image tokens, 61 graph replays
    wide = variant.endswith('bf16')                   # dim 256 / 4 heads: K = 256 and 1024, the shapes the small-M GEMM kernel takes
    m = D.DALLE(dim=256 if wide else 128, vae=vae, num_text_tokens=60, text_seq_len=12, depth=3, heads=4 if wide else 2, **kw).cuda().eval()
    for p in m.parameters():
        if p.dim() == 3:                                                   # LayerScale 0.1 -> every branch matters
            torch.nn.init.uniform_(p, 0.5, 1.0)
    text = torch.randint(1, 60, (3, 12), generator=torch.Generator().manual_seed(22)).cuda()
    made = []

    class Spy(decode.GraphedDecoder):
        def __init__(self, *a, **k):
            made.append(self)
            super().__init__(*a, **k)
    monkeypatch.setattr(decode, 'GraphedDecoder', Spy)
    toks = {}
    with D.compute_dtype_ctx(dtype):
        for on in (False, True):
            monkeypatch.setattr(decode, 'GRAPH_DEFAULT', on)
            torch.manual_seed(5)
            toks[on] = m.generate_images(text, use_cache=True, filter_thres=0.9).cpu()
        assert len(made) == 1 and made[0].graph is not None, 'the second call must have captured and replayed a graph'
        assert (made[0].plan is not None) == (mode != 'nest') and len(made[0].graphs) == (5 if mode == 'flat_bucket' else 1)
        assert int(made[0].pos_t) == m.total_seq_len and made[0].cache['offset'] == m.total_seq_len
        assert torch.equal(toks[False], toks[True]), (toks[False], toks[True])
        # teacher-forced logits, position by position (the decoder is driven by hand: eager warm-up steps, capture, replays)
        img = toks[False].cuda()
        host, dev = {}, {}
        with torch.no_grad():
            want = [m(text, img[:, :k], cache=host)[:, -1].float() for k in range(img.shape[1])]
            got = [m(text, img[:, :0], cache=dev)[:, -1].float()]
            dec = decode.GraphedDecoder(m, dev)
            for k in range(1, img.shape[1]):
                got.append(dec.step(img[:, k - 1]).float().clone())
        assert dec.graph is not None
    tol = (2e-2, 2e-2) if dtype == torch.bfloat16 else (1e-4, 1e-5)
    for k, (a, b) in enumerate(zip(want, got)):
        live = a > -1e30
        assert torch.equal(live, b > -1e30), k
        report(f'{variant}: graph-replayed logits step {k}', b[live], a[live], *tol)


def test_edge_shapes():
    """Ragged / degenerate shapes through the module API: batch 1, a text-only prefix shorter than text_len (no shift), one
    image token, n not a multiple of any tile."""
    import dalle_pytorch_b200 as D
    torch.manual_seed(13)
    t = D.Transformer(dim=64, depth=2, seq_len=24, heads=2, image_fmap_size=4, shift_tokens=True,
                      attn_types=('full', 'axial_row')).cuda()
    cfg = OracleConfig(dim=64, depth=2, heads=2, text_seq_len=8, fmap=4, num_text_tokens=50, num_image_tokens=32,
                       attn_types=('full', 'axial_row'))
    sd = {('transformer.' + k): v.detach().cpu() for k, v in t.state_dict().items()}
    for n in (1, 5, 9, 10, 23):
        x = torch.randn(1, n, 64)
        got = t(x.cuda())
        want = transformer_forward(x, sd, cfg)
        report(f'transformer n={n}', got, want, RTOL, ATOL)


@pytest.mark.parametrize('reversible', [False, True])
def test_flat_gradient_buffer_direct_write_world1(reversible):
    """Data-parallel plumbing on one GPU (NCCL world of 1): with NCCLBackend attached, the weight-gradient GEMMs write
    straight into the flat all-reduce buffer (functional._slot/_commit), shared weights fall back to autograd
    accumulation, the reversible executor hands gradients over as they complete; the gradients must equal the ones
    computed without the reducer."""
    import os
    import torch.distributed as dist
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200.distributed import NCCLBackend
    created = False
    if not dist.is_initialized():
        os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', '29617'
        os.environ['RANK'], os.environ['WORLD_SIZE'], os.environ['LOCAL_RANK'] = '0', '1', '0'
        created = True
    try:
        kw = dict(dim=256, num_text_tokens=200, text_seq_len=16, depth=4, heads=4, dim_head=64, reversible=reversible,
                  attn_types=('full', 'axial_row'), shared_attn_ids=(0, 1, 0, 1), shared_ff_ids=(0, 1, 2, 3))
        torch.manual_seed(3)
        m = D.DALLE(vae=D.TokenVAE(image_size=64, num_layers=3, num_tokens=64), **kw).cuda().train()
        text = torch.randint(1, 200, (2, 16)).cuda()
        image = torch.randint(0, 64, (2, 64)).cuda()
        for dt in (torch.float32, torch.bfloat16):
            with D.compute_dtype_ctx(dt):
                m.zero_grad(set_to_none=True)
                m(text, image, return_loss=True).backward()
                want = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
                be = NCCLBackend()
                be.initialize()
                be.distribute(model=m)
                red = m.grad_reducer
                for _ in range(2):                                           # two steps: the per-step reset must work
                    red.zero_grad()
                    m(text, image, return_loss=True).backward()
                    red.finish()
                direct = 0
                for k, p in m.named_parameters():
                    if k not in want:
                        continue
                    assert p.grad.data_ptr() == red.views[p].data_ptr(), k
                    tol = 1e-5 if dt == torch.float32 else 2e-2
                    report(f'{"rev" if reversible else "seq"} {dt} grad {k}', p.grad, want[k], tol, tol * float(want[k].abs().max()) + 1e-7)
                    direct += int(getattr(p, '_b200_uses', 0) == 1 and p.dim() == 2)
                assert direct >= 4
                red.remove()
                for p in m.parameters():
                    p.grad = None
    finally:
        if created and dist.is_initialized():
            dist.destroy_process_group()


def test_graphed_step_matches_eager():
    """GraphedStep (fwd+bwd captured into one CUDA graph, replayed): same loss and gradients as the eager step, for new inputs
    written into the static buffers, and stable over replays (gradients are overwritten, not accumulated)."""
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import ops
    cfg = OracleConfig(dim=256, depth=2, heads=4, text_seq_len=64, fmap=16, num_text_tokens=1000, num_image_tokens=512,
                       attn_types=('full', 'axial_col'))
    sd = make_state_dict(cfg, seed=2)
    m = build(cfg, sd).train()
    text, image = make_inputs(cfg, 4, seed=3)
    text2, image2 = make_inputs(cfg, 4, seed=4)
    with D.compute_dtype_ctx(torch.bfloat16):
        def eager(t, i):
            for p in m.parameters():
                p.grad = None
            loss = m(t.cuda(), i.cuda(), return_loss=True)
            loss.backward()
            return loss.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        l1, g1 = eager(text, image)
        l2, g2 = eager(text2, image2)
        step = D.GraphedStep(m, text.cuda(), image.cuda())
        assert step.kernels_per_step > 20
        n0 = ops.launches()
        for (t, i, lw, gw) in ((text, image, l1, g1), (text2, image2, l2, g2), (text, image, l1, g1)):
            loss = step(t.cuda(), i.cuda())
            report('graphed loss', loss, lw, 1e-5, 1e-6)
            for k, p in m.named_parameters():
                if k in gw:
                    report(f'graphed grad {k}', p.grad, gw[k], 1e-3, 1e-3 * float(gw[k].abs().max()) + 1e-7)
        assert ops.launches() == n0, 'a replay must not launch anything from Python'


@pytest.mark.parametrize('reversible', [False, True])
def test_dropout_matches_torch_with_the_same_masks(reversible):
    """attn_dropout / ff_dropout > 0 (attention.py:53-56, transformer.py:117) on the fused sub-layers, sequential and reversible:
    the masks are a pure function of the (seed, offset) pairs drawn from torch's generator, so a plain-torch evaluation of the
    reference math (the oracle's sub-layers with explicit masks) reproduces loss and gradients; the reversible executor must
    replay the same masks in its recomputation (reference Deterministic, reversible.py:20-50) -- its gradients equal those of
    autograd through the same two-stream forward."""
    import dalle_pytorch_b200 as D
    from dalle_pytorch_b200 import ops, functional
    cfg = OracleConfig(dim=64, depth=2, heads=2, text_seq_len=8, fmap=4, num_text_tokens=50, num_image_tokens=32,
                       attn_types=('full', 'axial_row'), reversible=reversible)
    sd = make_state_dict(cfg, seed=11)
    pa, pf = 0.25, 0.4
    vae = D.TokenVAE(image_size=32, num_layers=3, num_tokens=32)
    m = D.DALLE(dim=64, vae=vae, num_text_tokens=50, text_seq_len=8, depth=2, heads=2, attn_types=cfg.attn_types, reversible=reversible,
                attn_dropout=pa, ff_dropout=pf)
    m.load_state_dict(sd)
    m = m.cuda().train()
    text, image = make_inputs(cfg, 2, seed=12)
    b, n, d, H = 2, cfg.seq_len, 64, 256
    # the pairs the model will draw: replay torch's generator
    torch.manual_seed(99)
    pairs = [functional.DropoutRNG.draw(1) for _ in range(2 * cfg.depth)]
    masks = []
    for li in range(cfg.depth):
        ma = ops.dropout_(torch.ones(b * n * d, device='cuda'), pa, *pairs[2 * li]).view(b, n, d).cpu()
        mf = ops.dropout_(torch.ones(b * n * H, device='cuda'), pf, *pairs[2 * li + 1]).view(b, n, H).cpu()
        masks.append((ma, mf))
    torch.manual_seed(99)
    with D.compute_dtype_ctx(torch.float32):
        loss = m(text.cuda(), image.cuda(), return_loss=True)
        loss.backward()
    got = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    # plain torch evaluation of the same math with the same masks
    params = {k: v.clone().requires_grad_(k != 'transformer.pos_emb') for k, v in sd.items()}
    want = dalle_forward(text, image, params, cfg, return_loss=True, dropout_masks=masks)
    want.backward()
    report('loss with dropout', loss.detach(), want.detach(), RTOL, ATOL)
    for k, g in got.items():
        report(f'grad {k}', g, params[k].grad, RTOL, 2e-5)
    # a second step draws new masks
    torch.manual_seed(100)
    with D.compute_dtype_ctx(torch.float32):
        loss2 = m(text.cuda(), image.cuda(), return_loss=True)
    assert abs(float(loss2) - float(loss)) > 1e-6


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_sparse_attention_matches_dense_evaluation_of_its_layout(dtype):
    """SparseAttention (reference attention.py:339-398, DeepSpeed block-sparse): forward and gradients equal a dense torch
    evaluation of softmax over the documented layout (block_layout() expanded to tokens, intersected with causality), incl. a
    sequence that is not a multiple of the block size (the reference pads and slices, :369-376, :398)."""
    import dalle_pytorch_b200 as D
    from dalle_oracle import attention_core
    torch.manual_seed(17)
    dim, heads, seq_len, n = 128, 2, 297, 297                    # 41 text positions + 16 x 16 image tokens; not a multiple of the block
    a = D.SparseAttention(dim, seq_len, causal=True, heads=heads, block_size=16, text_seq_len=40).cuda()
    x = torch.randn(2, n, dim, device='cuda', requires_grad=True)
    ang = rotary_angle_table(41, 16, 64)[:n]
    with D.compute_dtype_ctx(dtype):
        out = a(x, rotary_pos_emb=ang.cuda()[None])
        g = torch.randn_like(out)
        out.backward(g)
    allow = a.static_mask.cpu()[:n, :n] & torch.ones(n, n).tril().bool()
    xr = x.detach().cpu().requires_grad_()
    P = {k: v.detach().cpu().requires_grad_() for k, v in dict(w_qkv=a.to_qkv.weight, w_out=a.to_out[0].weight, b_out=a.to_out[0].bias).items()}
    want = attention_core(xr, P['w_qkv'], P['w_out'], P['b_out'], heads, ang, allow, False)
    want.backward(g.cpu())
    tol = dict(rtol=RTOL, atol=2e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    report('sparse attn out', out, want, **tol)
    if dtype == torch.float32:
        report('sparse attn dx', x.grad, xr.grad, RTOL, 2e-5)
        report('sparse attn dWqkv', a.to_qkv.weight.grad, P['w_qkv'].grad, RTOL, 1e-4)
        report('sparse attn dWout', a.to_out[0].weight.grad, P['w_out'].grad, RTOL, 1e-4)
    else:
        report('sparse attn dx', x.grad, xr.grad, 0.0, 0.05 * float(xr.grad.abs().max()))
