"""GPU probe of the fused residual-tower kernel: correctness against nn.tower_reference with error maps, then timing.
    python tools/tower_probe.py [--board 5x5] [--time 16384]
"""
import argparse
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dotsboxesaz_b200 import engine  # noqa: E402
from dotsboxesaz_b200.nn import tower_pack, tower_reference  # noqa: E402


def run(eng, x, w3, b3, wh, bh):
    packed, bias = tower_pack(w3, b3, wh, bh)
    tiles = eng.tower_tiles(x.shape[0])
    eng.tower_planarize(x, tiles)
    hc = 0 if wh is None else wh.shape[0]
    out = torch.full((x.shape[0], x.shape[1], x.shape[2], hc or 64), 7.0, dtype=torch.bfloat16, device=x.device)
    eng.tower(tiles, packed, bias, w3.shape[0], hc, out)
    torch.cuda.synchronize()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", default="5x5")
    ap.add_argument("--time", type=int, default=0)
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--trace-boards", type=int, default=0)
    args = ap.parse_args()
    if args.trace:
        return trace(args.board, n_boards=args.trace_boards or None, blocks=args.blocks)
    L, C = (int(v) for v in args.board.split("x"))
    dev = torch.device("cuda:0")
    eng = engine.Engine((L, C), n_games=64, max_nodes=64, device=dev)
    g = eng.tower_geometry()
    print("geometry", g, flush=True)
    H, W = L + 1, C + 1
    torch.manual_seed(0)
    bad = 0
    for (n, S, hc) in [(g["nb"] - 3, 1, 0), (2 * g["nb"] + 5, 1, 0), (2 * g["nb"] + 5, 2, 0), (g["nb"] * 3, 2, 32), (g["nb"] * 150 + 7, 4, 32),
                       (g["nb"] * 2 + 1, 2 * args.blocks, 32)]:
        w3 = (torch.randn(S, 64, 64, 3, 3, device=dev) * (0.045 if S > 4 else 0.06))
        b3 = torch.randn(S, 64, device=dev) * 0.1
        wh = torch.randn(hc, 64, device=dev) * 0.2 if hc else None
        bh = torch.randn(hc, device=dev) * 0.1 if hc else None
        x = torch.rand(n, H, W, 64, device=dev).to(torch.bfloat16)
        ref = tower_reference(x, w3, b3, wh, bh).float()
        out = run(eng, x, w3, b3, wh, bh).float()
        err = (out - ref).abs()
        scale = ref.abs().max().item()
        tol = (0.01 if S <= 4 else 0.05) * scale
        nbad = int((err > tol).sum())
        print(f"n={n} stages={S} head={hc}: max|err| {err.max().item():.4g} mean|err| {err.mean().item():.3g} max|ref| {scale:.3g} "
              f"mean|ref| {ref.abs().mean().item():.3g}  bad {nbad}/{err.numel()}", flush=True)
        if nbad:
            bad += 1
            e = (err > tol).float()
            print("  bad fraction per h:", [round(v, 3) for v in e.mean((0, 2, 3)).tolist()])
            print("  bad fraction per w:", [round(v, 3) for v in e.mean((0, 1, 3)).tolist()])
            print("  bad fraction per channel group:", [round(v, 3) for v in e.reshape(n, H, W, -1, 8).mean((0, 1, 2, 4)).tolist()])
            print("  bad fraction per board (first 40):", [round(v, 2) for v in e.mean((1, 2, 3)).tolist()[:40]])
            print("  sample out", out[0, 0, 0, :8].tolist(), "ref", ref[0, 0, 0, :8].tolist())
            print("  sample out", out[0, 2, 3, :8].tolist(), "ref", ref[0, 2, 3, :8].tolist())
    if bad:
        print("TOWER PROBE: MISMATCH")
        sys.exit(1)
    print("TOWER PROBE: all configurations match", flush=True)
    if args.time:
        n, S, hc = args.time, 2 * args.blocks, 32
        w3 = torch.randn(S, 64, 64, 3, 3, device=dev) * 0.045
        b3 = torch.randn(S, 64, device=dev) * 0.1
        wh, bh = torch.randn(hc, 64, device=dev) * 0.2, torch.randn(hc, device=dev) * 0.1
        packed, bias = tower_pack(w3, b3, wh, bh)
        for nn_ in sorted({n, g["nb"] * 148, g["nb"] * 148 * 2, g["nb"] * 148 * 6}):
            x = torch.rand(nn_, H, W, 64, device=dev).to(torch.bfloat16)
            tiles = eng.tower_tiles(nn_)
            eng.tower_planarize(x, tiles)
            out = torch.empty((nn_, H, W, hc), dtype=torch.bfloat16, device=dev)
            for _ in range(3):
                eng.tower(tiles, packed, bias, S, hc, out)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                eng.tower(tiles, packed, bias, S, hc, out)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 100.0
            flops = nn_ * H * W * (S * 64 * 64 * 9 + hc * 64) * 2.0
            print(f"tower {nn_} boards x {S} stages: {us:.1f} us = {flops / us / 1e6:.1f} TFLOP/s (dense-equivalent)", flush=True)
    eng.close()




def trace(board="5x5", n_boards=None, blocks=20):
    """Timeline of CTA 0's first tile (clock64 cycles relative to the first stage)."""
    L, C = (int(v) for v in board.split("x"))
    dev = torch.device("cuda:0")
    eng = engine.Engine((L, C), n_games=64, max_nodes=64, device=dev)
    g = eng.tower_geometry()
    H, W = L + 1, C + 1
    S, hc = 2 * blocks, 32
    n = n_boards or g["nb"] * 148
    w3 = torch.randn(S, 64, 64, 3, 3, device=dev) * 0.045
    b3 = torch.randn(S, 64, device=dev) * 0.1
    wh, bh = torch.randn(hc, 64, device=dev) * 0.2, torch.randn(hc, device=dev) * 0.1
    packed, bias = tower_pack(w3, b3, wh, bh)
    x = torch.rand(n, H, W, 64, device=dev).to(torch.bfloat16)
    tiles = eng.tower_tiles(n)
    eng.tower_planarize(x, tiles)
    out = torch.empty((n, H, W, hc), dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        eng.tower(tiles, packed, bias, S, hc, out)
    tl = torch.zeros((64, 16), dtype=torch.int64, device=dev)
    eng.lib.dbaz_nn_tower_trace(eng._h, tl.data_ptr())
    eng.tower(tiles, packed, bias, S, hc, out)
    torch.cuda.synchronize()
    eng.lib.dbaz_nn_tower_trace(eng._h, None)
    t = tl.cpu()
    t0 = int(t[0, 0])
    print("stage | mma: start w0 w1 issued | epilogue per h-block: (acc ready, written) ... [cycles since stage 0 start]")
    for s in range(min(S, 8)):
        r = [int(v) - t0 for v in t[s].tolist()]
        print(s, r[:4], [(r[4 + 2 * h], r[5 + 2 * h]) for h in range(H)])
    eng.close()


if __name__ == "__main__":
    main()
