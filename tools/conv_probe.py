#!/usr/bin/env python
"""Probe: cuDNN fused conv+bias+ReLU variants vs plain conv on the SimpleNN trunk shape (bf16, NHWC)."""
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = False
dev = "cuda"


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n // 10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for N in (4096, 2304, 1024):
    x = torch.randn(N, 256, 4, 4, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(256, 256, 3, 3, device=dev, dtype=torch.bfloat16) * 0.02).contiguous(memory_format=torch.channels_last)
    b = torch.randn(256, device=dev, dtype=torch.bfloat16)
    z = torch.randn(N, 256, 4, 4, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t0 = timeit(lambda: F.conv2d(x, w, None, padding=1))
    t1 = timeit(lambda: F.conv2d(x, w, b, padding=1))
    try:
        t2 = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1))
    except Exception as e:
        t2 = str(e)[:80]
    try:
        t3 = timeit(lambda: torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, (1, 1), (1, 1), (1, 1), 1))
    except Exception as e:
        t3 = str(e)[:80]
    xp = F.pad(x, (1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
    t4 = timeit(lambda: F.conv2d(xp, w, None, padding=0))
    try:
        t5 = timeit(lambda: torch.cudnn_convolution_relu(xp, w, b, (1, 1), (0, 0), (1, 1), 1))
    except Exception as e:
        t5 = str(e)[:80]
    print(N, "conv %.1f | conv+bias %.1f | cudnn_conv_relu %s | cudnn_conv_add_relu %s | valid conv on padded 6x6 %.1f | conv_relu padded %s"
          % (t0, t1, t2, t3, t4, t5), flush=True)
    y = torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
    ref = F.relu(F.conv2d(x.float(), w.float(), b.float(), padding=1))
    print("   max err conv_relu vs fp32:", (y.float() - ref).abs().max().item(), "stride", y.stride())
