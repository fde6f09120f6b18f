"""The fused residual-tower kernel (dbaz_nn_tower, tcgen05 / tensor memory / TMA) against plain PyTorch float32 on
bf16-rounded weights with bf16 activations between stages (nn.tower_reference).  Reference: nn.py:16-58."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(eng, n, S, hc, seed):
    from dotsboxesaz_b200.nn import tower_pack, tower_reference
    dev = eng.device
    H, W = eng.rows, eng.cols
    g = torch.Generator(device="cpu").manual_seed(seed)
    w3 = (torch.randn(S, 64, 64, 3, 3, generator=g) * (0.045 if S > 4 else 0.06)).to(dev)
    b3 = (torch.randn(S, 64, generator=g) * 0.1).to(dev)
    wh = (torch.randn(hc, 64, generator=g) * 0.2).to(dev) if hc else None
    bh = (torch.randn(hc, generator=g) * 0.1).to(dev) if hc else None
    x = torch.rand(n, H, W, 64, generator=g).to(dev).to(torch.bfloat16)
    packed, bias = tower_pack(w3, b3, wh, bh)
    tiles = eng.tower_tiles(n)
    eng.tower_planarize(x, tiles)
    out = torch.full((n, H, W, hc or 64), 7.0, dtype=torch.bfloat16, device=dev)
    eng.tower(tiles, packed, bias, S, hc, out)
    torch.cuda.synchronize()
    ref = tower_reference(x, w3, b3, wh, bh).float()
    return out.float(), ref


@pytest.mark.parametrize("board", [(5, 5), (3, 3), (2, 3), (5, 2)])
def test_tower_short(board):
    """1, 2 and 4 stages, with and without the fused head: at most one bf16 step away from the reference."""
    from dotsboxesaz_b200 import engine
    eng = engine.Engine(board, n_games=8, max_nodes=16)
    nb = eng.tower_geometry()["nb"]
    try:
        for (n, S, hc) in [(nb - 2, 1, 0), (2 * nb + 3, 2, 0), (nb * 149 + 1, 2, 32), (3 * nb, 4, 16)]:
            out, ref = _case(eng, n, S, hc, seed=n + S)
            scale = ref.abs().max().item()
            err = (out - ref).abs()
            # bf16 has 8 significant bits: one rounding step of the largest values, a few after four stages
            assert err.max().item() <= (2.0 ** -7) * scale * (1 if S <= 2 else 2), (board, n, S, hc, err.max().item(), scale)
            assert err.mean().item() <= 2e-3 * ref.abs().mean().item() + 1e-6
    finally:
        eng.close()


def test_tower_full_depth():
    """20 residual blocks + head on 5x5 and 3x3 boards: errors stay at the rounding level of bf16 activations."""
    from dotsboxesaz_b200 import engine
    for board in ((5, 5), (3, 3)):
        eng = engine.Engine(board, n_games=8, max_nodes=16)
        try:
            n = 5 * eng.tower_geometry()["nb"] + 2
            out, ref = _case(eng, n, 40, 32, seed=3)
            err = (out - ref).abs()
            assert err.max().item() <= 0.05 * ref.abs().max().item()
            assert err.mean().item() <= 0.01 * ref.abs().mean().item()
        finally:
            eng.close()


def test_tower_rejects_unsupported():
    from dotsboxesaz_b200 import engine
    eng = engine.Engine((6, 6), n_games=8, max_nodes=16)
    try:
        assert eng.tower_geometry()["ok"] == 0
    finally:
        eng.close()


@pytest.mark.parametrize("board", [(5, 5), (3, 3)])
def test_tower_full_size_is_position_independent(board):
    """BASELINE configs[3] size (16 384 leaves, 20 blocks + head): a board's result must not depend on where it sits in the
    batch (tile, h-block row, CTA, wave of the persistent grid).  64 distinct boards repeated 256 times: every copy
    bit-equal to the first, and the 64 distinct results equal to the reference."""
    from dotsboxesaz_b200 import engine
    from dotsboxesaz_b200.nn import tower_pack, tower_reference
    eng = engine.Engine(board, n_games=8, max_nodes=16)
    try:
        dev, H, W = eng.device, eng.rows, eng.cols
        g = torch.Generator(device="cpu").manual_seed(11)
        S, hc, n, k = 40, 32, 16384, 64
        w3 = (torch.randn(S, 64, 64, 3, 3, generator=g) * 0.045).to(dev)
        b3 = (torch.randn(S, 64, generator=g) * 0.1).to(dev)
        wh, bh = (torch.randn(hc, 64, generator=g) * 0.2).to(dev), (torch.randn(hc, generator=g) * 0.1).to(dev)
        base = torch.rand(k, H, W, 64, generator=g).to(dev).to(torch.bfloat16)
        x = base.repeat(n // k, 1, 1, 1).contiguous()
        packed, bias = tower_pack(w3, b3, wh, bh)
        tiles = eng.tower_tiles(n)
        eng.tower_planarize(x, tiles)
        out = torch.empty((n, H, W, hc), dtype=torch.bfloat16, device=dev)
        eng.tower(tiles, packed, bias, S, hc, out)
        torch.cuda.synchronize()
        first = out[:k].view(torch.int16)
        assert torch.equal(out.view(torch.int16).reshape(n // k, k, H, W, hc), first.unsqueeze(0).expand(n // k, k, H, W, hc))
        ref = tower_reference(base, w3, b3, wh, bh).float()
        err = (out[:k].float() - ref).abs()
        assert err.max().item() <= 0.05 * ref.abs().max().item() and err.mean().item() <= 0.01 * ref.abs().mean().item()
        # a second launch on the same inputs: bit-identical (no dependence on timing between the roles)
        out2 = torch.empty_like(out)
        eng.tower(tiles, packed, bias, S, hc, out2)
        torch.cuda.synchronize()
        assert torch.equal(out.view(torch.int16), out2.view(torch.int16))
    finally:
        eng.close()
