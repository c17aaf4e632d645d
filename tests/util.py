import torch


def report(name, got, want, rtol, atol):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    assert got.shape == want.shape, f'{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}'
    if not torch.isfinite(got).all():
        bad = (~torch.isfinite(got)).sum().item()
        raise AssertionError(f'{name}: {bad} non-finite values of {got.numel()}')
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    worst = (err - tol).max().item()
    if worst > 0:
        idx = (err - tol).argmax().item()
        flat_g, flat_w = got.reshape(-1), want.reshape(-1)
        nbad = (err > tol).sum().item()
        raise AssertionError(f'{name}: {nbad}/{got.numel()} outside rtol={rtol} atol={atol}; max abs err {err.max().item():.3e} '
                             f'(|want| max {want.abs().max().item():.3e}); worst at flat {idx}: got {flat_g[idx].item():.6e} '
                             f'want {flat_w[idx].item():.6e}')
    return err.max().item()
