"""CPU emulation of k_resnet_tower's data movement (dotsboxesaz_b200/csrc/dbaz_tower.cu): the planar shared-memory
layout, the row-shifted A operands, the three dy taps stacked along N, the packed weight chunks and the epilogue --
everything except the tensor-core hardware itself -- checked against nn.tower_reference.  Run on the CPU:
    python tools/tower_emulate.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dotsboxesaz_b200.nn import tower_pack, tower_reference  # noqa: E402


def bf16(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).float().numpy()


def emulate(x_nhwc, w3, b3, w_head=None, b_head=None):
    n, H, W, _ = x_nhwc.shape
    WP, nb = W + 1, 128 // (W + 1)
    plane_slots = H * 128 + 8
    buf_slots = 8 + 8 * plane_slots
    packed, bias = tower_pack(w3, b3, w_head, b_head)
    pk = packed.view(torch.bfloat16).float().numpy()
    bias = bias.numpy()
    S3 = w3.shape[0]
    hc = 0 if w_head is None else w_head.shape[0]
    S = S3 + (1 if hc else 0)
    n_tiles = -(-n // nb)
    out = np.zeros((n, H, W, hc or 64), np.float32)
    xs = x_nhwc.float().numpy()
    for tile in range(n_tiles):
        X = np.zeros((buf_slots, 8), np.float32)
        Y = np.zeros((buf_slots, 8), np.float32)
        for bi in range(nb):                       # planarize
            board = tile * nb + bi
            if board >= n:
                break
            for h in range(H):
                for w in range(W):
                    for cg in range(8):
                        X[8 + cg * plane_slots + h * 128 + bi * WP + w] = xs[board, h, w, 8 * cg:8 * cg + 8]
        for s in range(S):
            src, dst = (Y, X) if s & 1 else (X, Y)
            acc = np.zeros((H, 128, 64), np.float32)
            if s < S3:
                for p in range(12):
                    chunk = pk[(s * 12 + p) * 3072:(s * 12 + p + 1) * 3072].reshape(2, 192, 8)
                    dx, k = (p >> 2) - 1, p & 3
                    for hp in range(H):
                        a = np.stack([src[8 + (2 * k + kh) * plane_slots + hp * 128 + dx:8 + (2 * k + kh) * plane_slots + hp * 128 + dx + 128]
                                      for kh in range(2)], 0)            # [kh, 128 rows, 8]
                        lo, hi = max(hp - 1, 0), min(hp + 1, H - 1)
                        brow = (lo - (hp - 1)) * 64
                        nn_ = 64 * (hi - lo + 1)
                        bmat = chunk[:, brow:brow + nn_]                # [kh, N, 8]
                        d = np.einsum("kre,kne->rn", a, bmat)          # [128, N]
                        for j in range(hi - lo + 1):
                            acc[lo + j] += d[:, 64 * j:64 * j + 64]
            else:
                off = S3 * 12 * 3072
                chunk = pk[off:off + 4 * 2 * hc * 8].reshape(4, 2, hc, 8)
                for hp in range(H):
                    for k in range(4):
                        a = np.stack([src[8 + (2 * k + kh) * plane_slots + hp * 128:8 + (2 * k + kh) * plane_slots + hp * 128 + 128] for kh in range(2)], 0)
                        acc[hp, :, :hc] += np.einsum("kre,kne->rn", a, chunk[k])
            head, last = s >= S3, s == S - 1
            for h in range(H):
                for row in range(128):
                    bi, w = divmod(row, WP)
                    ok = bi < nb and w < W
                    v = acc[h, row] + bias[s]
                    if (not head) and (s & 1):
                        v = v + np.concatenate([dst[8 + cg * plane_slots + h * 128 + row] for cg in range(8)])
                    v = bf16(np.maximum(v, 0.0)) if ok else np.zeros(64, np.float32)
                    if not last:
                        for cg in range(8):
                            dst[8 + cg * plane_slots + h * 128 + row] = v[8 * cg:8 * cg + 8]
                    elif ok and tile * nb + bi < n:
                        out[tile * nb + bi, h, w] = v[:hc] if head else v
    return out


def main():
    torch.manual_seed(1)
    for (H, W, n, S, hc) in [(6, 6, 20, 1, 0), (6, 6, 20, 2, 0), (4, 4, 27, 4, 32), (3, 5, 23, 2, 16), (6, 6, 19, 4, 32)]:
        w3 = torch.randn(S, 64, 64, 3, 3) * 0.06
        b3 = torch.randn(S, 64) * 0.1
        wh = torch.randn(hc, 64) * 0.2 if hc else None
        bh = torch.randn(hc) * 0.1 if hc else None
        x = torch.rand(n, H, W, 64).to(torch.bfloat16)
        ref = tower_reference(x, w3, b3, wh, bh).float().numpy()
        got = emulate(x, w3, b3, wh, bh)
        err = np.abs(got - ref)
        print(f"H={H} W={W} n={n} stages={S} head={hc}: max |err| {err.max():.4g} (max |ref| {np.abs(ref).max():.3g}), "
              f"mismatching elements {(err > 0.02 * np.abs(ref).max()).sum()}")
        assert err.max() <= 0.02 * np.abs(ref).max() + 1e-3


if __name__ == "__main__":
    main()
