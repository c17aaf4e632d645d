"""CPU-side checks (run with -m "not gpu"): the C-ABI library loads and exports every symbol the header declares,
ctypes struct layouts match the C structs, and the host logic of the package (module tree / state-dict keys, rotary
table, static masks, token-shift restatement, argument routing) agrees with the oracle / reference."""
import ctypes
import os
import re

import pytest
import torch

import dalle_pytorch_b200 as D
from dalle_pytorch_b200 import _lib
from dalle_pytorch_b200.attention import rotary_tables
from dalle_pytorch_b200.transformer import PreShiftToken, build_rotary_angle_table
from dalle_pytorch_b200.reversible import route_args
from dalle_oracle import OracleConfig, make_state_dict, rotary_angle_table, token_shift, allowed_mask

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.lib()
    assert lib.dalle_b200_version() == 112
    hdr = open(os.path.join(ROOT, 'include', 'dalle_b200.h')).read()
    declared = set(re.findall(r'\b(dalle_b200_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    for name in declared:
        assert hasattr(lib, name), f'{name} not exported by libdalle_b200.so'


def test_struct_layouts_match_c():
    lib = _lib.lib()
    sizes = (ctypes.c_int * 8)()
    n = lib.dalle_b200_abi_sizes(sizes, 8)
    assert n == len(_lib._STRUCTS)
    for i, st in enumerate(_lib._STRUCTS):
        assert ctypes.sizeof(st) == sizes[i], st.__name__


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = _lib.lib()
    P = _lib.GemmParams(M=4, N=3, K=8, dtype=0, epilogue=0)     # odd N, null operands
    rc = lib.dalle_b200_gemm(ctypes.byref(P), None)
    assert rc == -1 and b'null operand' in lib.dalle_b200_last_error()
    A = _lib.AttnFwdParams(batch=1, heads=1, n_q=4, n_k=4, dim_head=32)
    rc = lib.dalle_b200_attn_fwd(ctypes.byref(A), None)
    assert rc == -2 and b'dim_head' in lib.dalle_b200_last_error()
    # optimizer / embedding entry points validate before touching the device as well
    O = _lib.AdamParams(count=16, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, step=1)
    assert lib.dalle_b200_adam(ctypes.byref(O), None) == -1 and b'adam' in lib.dalle_b200_last_error()
    assert lib.dalle_b200_sumsq(None, 4, None, None) == -1
    assert lib.dalle_b200_embed_fwd(None, None, None, 1, 1, 1, 0, 4, 8, None) == -1 and b'embed_fwd' in lib.dalle_b200_last_error()
    L = _lib.LnShiftBwdParams(batch=1, n=4, d=64, do_ln=0, do_shift=0, dout_dtype=0)
    assert lib.dalle_b200_ln_shift_bwd(ctypes.byref(L), None) < 0


@pytest.mark.parametrize('kw', [dict(), dict(reversible=True), dict(shift_tokens=False), dict(sandwich_norm=True),
                                dict(attn_types=('full', 'axial_row', 'axial_col', 'conv_like'), depth=4)])
def test_state_dict_keys_match_reference_format(kw):
    base = dict(dim=64, depth=2, heads=2, text_seq_len=8, fmap=4, num_text_tokens=50, num_image_tokens=32)
    base.update(kw)
    cfg = OracleConfig(**base)
    sd = make_state_dict(cfg)       # keys asserted equal to the live reference's in oracle/make_golden.py
    m = D.DALLE(dim=cfg.dim, vae=D.TokenVAE(image_size=8 * cfg.fmap, num_layers=3, num_tokens=cfg.num_image_tokens),
                num_text_tokens=cfg.num_text_tokens, text_seq_len=cfg.text_seq_len, depth=cfg.depth, heads=cfg.heads,
                reversible=cfg.reversible, attn_types=cfg.attn_types, sandwich_norm=cfg.sandwich_norm, shift_tokens=cfg.shift_tokens)
    own = m.state_dict()
    assert set(own) == set(sd)
    for k in own:
        assert own[k].shape == sd[k].shape, k
    assert torch.equal(own['transformer.pos_emb'], sd['transformer.pos_emb'])
    m.load_state_dict(sd)


def test_rotary_table_and_cos_sin():
    for T, fm in ((9, 4), (65, 8), (257, 32)):
        assert torch.equal(build_rotary_angle_table(T, fm, 64)[0], rotary_angle_table(T, fm, 64))
    ang = build_rotary_angle_table(9, 4, 64)
    cos, sin = rotary_tables(ang, 64)
    assert cos.shape == (25, 32)
    assert torch.allclose(cos[:, :30], ang[0, :, 0::2].double().cos().float()) and torch.all(cos[:, 30:] == 1)
    assert torch.all(sin[:, 30:] == 0)


def test_token_shift_restatement_matches_oracle():
    ps = PreShiftToken(lambda x, **kw: x, image_size=4, seq_len=24)
    for n in (24, 20, 9):
        x = torch.randn(2, n, 32)
        assert torch.equal(ps(x), token_shift(x, 9, 4))
    x = torch.randn(2, 5, 32)
    assert torch.equal(ps(x), x)


def test_static_masks_match_predicate():
    t = D.Transformer(dim=64, depth=1, seq_len=24, heads=2, image_fmap_size=4)
    caus = torch.ones(24, 24).tril().bool()
    for kind in ('axial_row', 'axial_col'):
        assert torch.equal(t._get_attention_mask(kind) & caus, allowed_mask(kind, 24, 24, 9, 4))


def test_route_args():
    r = route_args({'mask': ((True, False),) * 2, 'cache': ((True, True),) * 2}, dict(mask=1, cache=2, other=3), 2)
    assert r == [({'mask': 1, 'cache': 2}, {'cache': 2})] * 2


def test_plan_resolves_fused_sublayers():
    t = D.Transformer(dim=64, depth=2, seq_len=24, heads=2, image_fmap_size=4, shift_tokens=True,
                      attn_types=('axial_row', 'conv_like'))
    x = torch.zeros(1, 24, 64)
    f0, g0 = t.layers.layers[0]
    p = f0.plan(x, rotary_pos_emb=t.pos_emb)
    assert p.kind == 'attn' and p.geom.do_shift and p.geom.text_len == 9 and p.geom.attn_spec.pattern == _lib.ATTN_AXIAL_ROW
    assert g0.plan(x).kind == 'ff'
    assert t.layers.layers[1][0].plan(x).geom.attn_spec.pattern == _lib.ATTN_CONV_LIKE
    assert f0.plan(x, cache={}) is None           # inference cache -> module-by-module path


def test_ops_refuse_cpu_tensors():
    from dalle_pytorch_b200 import ops
    with pytest.raises(AssertionError):
        ops.colsum(torch.zeros(4, 4))


def test_patch_dalle_pytorch_rebinds_the_live_reference():
    """`patch_dalle_pytorch()` (SURVEY.md §7 drop-in installer): an UNMODIFIED reference DALLE built after patching carries this
    package's modules with the reference's state-dict keys; undo() restores the reference classes."""
    import ref_import
    if not ref_import.reference_available():
        pytest.skip('reference not present')
    ref = ref_import.import_reference()
    undo = D.patch_dalle_pytorch()
    try:
        vae = ref.DiscreteVAE(image_size=32, num_layers=3, num_tokens=32, codebook_dim=16, hidden_dim=8)
        kw = dict(dim=64, num_text_tokens=50, text_seq_len=8, depth=4, heads=2, dim_head=64,
                  attn_types=('full', 'axial_row', 'axial_col', 'conv_like'))
        for rev in (False, True):
            m = ref.DALLE(vae=vae, reversible=rev, **kw)
            ours = D.DALLE(vae=D.TokenVAE(image_size=32, num_layers=3, num_tokens=32), reversible=rev, **kw)
            assert type(m).__module__.startswith('dalle_pytorch.')                 # the reference's own DALLE / Transformer ...
            assert type(m.transformer).__module__ == 'dalle_pytorch.transformer'
            assert type(m.transformer.layers).__module__ == 'dalle_pytorch_b200.reversible'   # ... around this package's blocks
            kinds = {type(x).__name__: type(x).__module__ for x in m.transformer.modules()}
            for name in ('Attention', 'SparseAxialCausalAttention', 'SparseConvCausalAttention', 'FeedForward', 'LayerScale', 'PreNorm',
                         'PreShiftToken'):
                assert kinds[name].startswith('dalle_pytorch_b200.'), (name, kinds[name])
            own = {k: v.shape for k, v in ours.state_dict().items()}
            got = {k: v.shape for k, v in m.state_dict().items() if not k.startswith('vae.')}
            assert own == got
            ours.load_state_dict({k: v for k, v in m.state_dict().items() if not k.startswith('vae.')})
    finally:
        undo()
    m = ref.DALLE(vae=vae, **kw)
    assert type(m.transformer.layers).__module__ == 'dalle_pytorch.reversible'


def test_sparse_attention_layout_spec():
    """The written layout spec of SparseAttention (attention.py docstring; reference attention.py:339-365): global text blocks,
    local window, `num_random_blocks` unidirectional random blocks, reproducible from layout_seed."""
    a = D.SparseAttention(64, 304, causal=True, heads=2, block_size=16, text_seq_len=48)
    L = a.block_layout()
    nb = 19
    assert L.shape == (nb, nb) and a.num_global_blocks == 3 and a.num_random_blocks == 304 // 16 // 4
    assert not torch.triu(L, 1).any()                                          # unidirectional
    for r in range(nb):
        assert L[r, :min(3, r + 1)].all()                                      # global (text) blocks at or before the row
        assert L[r, 4 * (r // 4):r + 1].all()                                  # local window
        extra = L[r].clone()
        extra[:3] = False
        extra[4 * (r // 4):r + 1] = False
        assert int(extra.sum()) <= a.num_random_blocks                         # at most num_random_blocks more
    assert torch.equal(L, D.SparseAttention(64, 304, heads=2, block_size=16, text_seq_len=48).block_layout())      # reproducible
    assert not torch.equal(L, D.SparseAttention(64, 304, heads=2, block_size=16, text_seq_len=48, layout_seed=1).block_layout())
    m = a.static_mask
    assert m.shape == (304, 304) and torch.equal(m, L.repeat_interleave(16, 0).repeat_interleave(16, 1))
    assert isinstance(a, D.Attention)                                          # transformer.py:279 decides cache support by this


def test_small_m_gemm_dispatch_rule():
    """ops._small_m: which problems the host sends to DB200_GEMM_SMALLM (bf16, K-major, 1 <= M <= 16, K % 256 == 0, N % 16 == 0)."""
    import torch
    from dalle_pytorch_b200 import ops, _lib
    assert _lib.GEMM_SMALLM == 3
    bf = lambda m, k: torch.empty(m, k, dtype=torch.bfloat16)
    assert ops._small_m(bf(16, 1024), 3072) and ops._small_m(bf(1, 256), 16)
    assert not ops._small_m(bf(17, 1024), 3072)                      # more rows than one mma fragment
    assert not ops._small_m(bf(16, 1000), 3072) and not ops._small_m(bf(16, 128), 3072)
    assert not ops._small_m(bf(16, 1024), 3080)
    assert not ops._small_m(torch.empty(16, 1024), 3072)             # fp32 parity mode keeps its own GEMM evaluation
    assert not ops._small_m(bf(16, 1024), 3072, a_mn=True) and not ops._small_m(bf(16, 1024), 3072, b_mn=True)
    was, ops.SMALL_M = ops.SMALL_M, False
    try:
        assert not ops._small_m(bf(16, 1024), 3072)                  # DALLE_B200_SMALLM=0
    finally:
        ops.SMALL_M = was


def test_header_is_plain_c_and_a_c_client_links_against_the_library(tmp_path):
    """include/dalle_b200.h is the drop-in boundary: it must compile as C (no C++ / CUDA / torch types) and a C program that only
    knows the header must link against libdalle_b200.so and read the version and struct sizes (no GPU needed: no compute call)."""
    import os
    import shutil
    import subprocess
    from dalle_pytorch_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, 'include')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-fsyntax-only', '-x', 'c', os.path.join(hdr, 'dalle_b200.h')],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = tmp_path / 'client.c'
    src.write_text('#include <stdio.h>\n#include "dalle_b200.h"\n'
                   'int main(void) {\n'
                   '  int sizes[8]; int n = dalle_b200_abi_sizes(sizes, 8);\n'
                   '  db200_gemm_params g; db200_attn_fwd_params a;\n'
                   '  printf("%d %d %d %d %d\\n", dalle_b200_version(), n, sizes[2] == (int)sizeof g, sizes[3] == (int)sizeof a, DB200_GEMM_SMALLM);\n'
                   '  return dalle_b200_gemm(0, 0) == DB200_OK;   /* NULL params are rejected with an error code, not a crash */\n'
                   '}\n')
    exe = tmp_path / 'client'
    _lib.lib()                                         # built in-tree
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(['gcc', '-std=c99', '-I', hdr, str(src), '-o', str(exe), '-L', libdir, '-ldalle_b200', '-Wl,-rpath,' + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    ver, n, gemm_ok, attn_ok, smallm = map(int, r.stdout.split())
    assert ver == 112 and n >= 4 and gemm_ok == 1 and attn_ok == 1 and smallm == 3
