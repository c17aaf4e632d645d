"""Pins oracle/dalle_oracle.py (the CPU restatement) against the golden vectors generated from the
UNMODIFIED reference by oracle/make_golden.py, and against the live reference when it is present.
Tolerance: BASELINE.json north_star — rtol 1e-3 / atol 1e-5 in fp32 (the oracle usually agrees to ~1e-6)."""
import pytest
import torch

from conftest import load_golden, golden_names
from dalle_oracle import (OracleConfig, make_state_dict, make_inputs, dalle_forward, allowed_mask,
                          attention_core, rotary_angle_table, token_shift)
import ref_import

RTOL, ATOL = 1e-3, 1e-5


def _cfg(rec):
    c = dict(rec['cfg'])
    c['attn_types'] = tuple(c['attn_types'])
    return OracleConfig(**c)


def _oracle_run(rec):
    cfg = _cfg(rec)
    sd = make_state_dict(cfg, seed=rec['seed'])
    for k, v in rec['weight_checksums'].items():
        assert abs(float(sd[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), f'synthetic weight drift in {k}'
    params = {k: v.clone().requires_grad_(k != 'transformer.pos_emb') for k, v in sd.items()}
    loss = dalle_forward(rec['text'].clone(), rec['image'].clone(), params, cfg, return_loss=True)
    loss.backward()
    with torch.no_grad():
        logits = dalle_forward(rec['text'].clone(), rec['image'].clone(), sd, cfg)
    grads = {k: p.grad for k, p in params.items() if p.grad is not None}
    return loss.detach(), logits, grads


# fixtures whose weights are tied (stored with the fixture): the oracle restates the un-tied block stack; these goldens pin
# the product directly (tests/test_parity_gpu.py::test_tied_weight_goldens_fp32)
TIED = {'tiny_shared', 'tiny_tied_emb'}


@pytest.mark.parametrize('name', [n for n in golden_names('tiny_') if n not in TIED])
def test_oracle_matches_reference_tiny(name):
    rec = load_golden(name)
    loss, logits, grads = _oracle_run(rec)
    torch.testing.assert_close(loss, rec['loss'], rtol=RTOL, atol=ATOL)
    masked = rec['logits'] < -1e30
    assert torch.equal(masked, logits < -1e30)
    assert torch.equal(logits[masked], rec['logits'][masked])          # exactly -fp32max (dalle_pytorch.py:651-652)
    torch.testing.assert_close(logits[~masked], rec['logits'][~masked], rtol=RTOL, atol=ATOL)
    assert set(grads) == set(rec['grads'])
    for k, g in rec['grads'].items():
        torch.testing.assert_close(grads[k], g, rtol=RTOL, atol=ATOL, msg=lambda m: f'{k}: {m}')


@pytest.mark.parametrize('name', golden_names('c1_') + ['c3_geom'])
def test_oracle_matches_reference_c1(name):
    rec = load_golden(name)
    loss, logits, grads = _oracle_run(rec)
    torch.testing.assert_close(loss, rec['loss'], rtol=RTOL, atol=ATOL)
    samp = logits[..., ::rec['logits_stride']]
    masked = rec['logits_sample'] < -1e30
    assert torch.equal(samp[masked], rec['logits_sample'][masked])
    torch.testing.assert_close(samp[~masked], rec['logits_sample'][~masked], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(torch.logsumexp(logits.double(), -1).float(), rec['logits_lse'], rtol=RTOL, atol=1e-4)
    for k, (vals, step) in rec['grad_samples'].items():
        mine = grads[k].reshape(-1)[::step][:vals.numel()]
        torch.testing.assert_close(mine, vals, rtol=RTOL, atol=ATOL, msg=lambda m: f'{k}: {m}')
        assert abs(float(grads[k].double().norm()) - rec['grad_norms'][k]) <= 1e-3 * rec['grad_norms'][k] + 1e-7


@pytest.mark.skipif(not ref_import.reference_available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('kind,axis', [('axial_row', 0), ('axial_col', 1)])
@pytest.mark.parametrize('stable', [False, True])
@pytest.mark.parametrize('n', [24, 19])
def test_oracle_attention_matches_live_reference_axial(kind, axis, stable, n):
    """Module-level pin: SparseAxialCausalAttention (attention.py:225-335) == oracle attention_core with the
    allowed-key predicate, including n < seq_len (padded tail sliced off, attention.py:255-258, 335)."""
    R = ref_import.import_reference()
    torch.manual_seed(0)
    dim, heads, fmap, text_seq = 64, 2, 4, 8
    seq_len = text_seq + fmap * fmap
    text_len = seq_len - fmap * fmap + 1
    m = R.SparseAxialCausalAttention(dim, seq_len, image_size=fmap, axis=axis, heads=heads, dim_head=64, stable=stable)
    x = torch.randn(2, n, dim)
    ang = rotary_angle_table(text_len, fmap, 64)
    ref = m(x, rotary_pos_emb=ang[None])
    allow = allowed_mask(kind, n, n, text_len, fmap)
    mine = attention_core(x, m.to_qkv.weight, m.to_out[0].weight, m.to_out[0].bias, heads, ang, allow, stable)
    torch.testing.assert_close(mine, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.skipif(not ref_import.reference_available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('kernel_size,dilation', [(3, 1), (5, 1), (3, 2)])
def test_oracle_attention_matches_live_reference_conv_like(kernel_size, dilation):
    R = ref_import.import_reference()
    torch.manual_seed(0)
    dim, heads, fmap, text_seq = 64, 2, 6, 5
    seq_len = text_seq + fmap * fmap
    text_len = text_seq + 1
    m = R.SparseConvCausalAttention(dim, seq_len, image_size=fmap, kernel_size=kernel_size, dilation=dilation,
                                    heads=heads, dim_head=64)
    x = torch.randn(2, seq_len, dim)
    ang = rotary_angle_table(text_len, fmap, 64)
    ref = m(x, rotary_pos_emb=ang[None])
    allow = allowed_mask('conv_like', seq_len, seq_len, text_len, fmap, kernel_size=kernel_size, dilation=dilation)
    mine = attention_core(x, m.to_qkv.weight, m.to_out[0].weight, m.to_out[0].bias, heads, ang, allow, False)
    torch.testing.assert_close(mine, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.skipif(not ref_import.reference_available(), reason='/root/reference not present (GPU box)')
def test_oracle_token_shift_matches_live_reference():
    R = ref_import.import_reference()
    torch.manual_seed(0)
    fmap, text_seq, dim = 4, 8, 32
    seq_len = text_seq + fmap * fmap
    sh = R.transformer.PreShiftToken(lambda x, **kw: x, image_size=fmap, seq_len=seq_len)
    for n in (seq_len, seq_len - 3, text_seq + 1, 5):
        x = torch.randn(2, n, dim)
        assert torch.equal(sh(x), token_shift(x, text_seq + 1, fmap)), n


@pytest.mark.skipif(not ref_import.reference_available(), reason='/root/reference not present (GPU box)')
def test_static_mask_equals_predicate():
    """The reference's own cross-check (transformer.py:333-350): static axial masks AND causal == predicate."""
    R = ref_import.import_reference()
    t = R.Transformer(dim=64, depth=1, seq_len=24, heads=2, dim_head=64, image_fmap_size=4, attn_types=('full',))
    caus = torch.ones(24, 24).tril().bool()
    for kind in ('axial_row', 'axial_col'):
        sm = t._get_attention_mask(kind)
        assert torch.equal(sm & caus, allowed_mask(kind, 24, 24, 9, 4))
