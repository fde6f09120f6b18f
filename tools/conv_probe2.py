#!/usr/bin/env python
"""Probe: does cudnn_convolution_add_relu accept a batch-broadcast (stride-0) z, and what does it cost?"""
import torch
import torch.nn.functional as F
from conv_probe import timeit

dev = "cuda"
for N in (4096, 1024):
    x = torch.randn(N, 256, 4, 4, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(256, 256, 3, 3, device=dev, dtype=torch.bfloat16) * 0.02).contiguous(memory_format=torch.channels_last)
    b = torch.randn(256, device=dev, dtype=torch.bfloat16)
    zpos = torch.randn(256, 4, 4, device=dev, dtype=torch.bfloat16)
    z_full = zpos.unsqueeze(0).expand(N, -1, -1, -1).contiguous(memory_format=torch.channels_last)
    z_b = zpos.permute(1, 2, 0).contiguous().permute(2, 0, 1).unsqueeze(0).expand(N, -1, -1, -1)
    print("z_b strides", z_b.stride(), "z_full strides", z_full.stride())
    ref = torch.cudnn_convolution_add_relu(x, w, z_full, 1.0, b, (1, 1), (1, 1), (1, 1), 1)
    try:
        out = torch.cudnn_convolution_add_relu(x, w, z_b, 1.0, b, (1, 1), (1, 1), (1, 1), 1)
        print(N, "stride-0 z accepted; equal:", torch.equal(out, ref), "max diff", (out.float() - ref.float()).abs().max().item())
        print(N, "time full z %.1f us, stride-0 z %.1f us" % (
            timeit(lambda: torch.cudnn_convolution_add_relu(x, w, z_full, 1.0, b, (1, 1), (1, 1), (1, 1), 1)),
            timeit(lambda: torch.cudnn_convolution_add_relu(x, w, z_b, 1.0, b, (1, 1), (1, 1), (1, 1), 1))))
    except Exception as e:
        print(N, "stride-0 z rejected:", str(e)[:200])
    # FC: fused relu epilogue
    a = torch.randn(N, 1024, device=dev, dtype=torch.bfloat16)
    w0 = torch.randn(1024, 512, device=dev, dtype=torch.bfloat16) * 0.02
    b0 = torch.randn(512, device=dev, dtype=torch.bfloat16)
    t_mm = timeit(lambda: torch.mm(a, w0))
    t_act = timeit(lambda: torch._addmm_activation(b0, a, w0))
    r1 = torch._addmm_activation(b0, a, w0)
    r2 = F.relu(torch.addmm(b0, a, w0))
    print(N, "mm %.1f us, _addmm_activation(relu) %.1f us, max diff %.4f" % (t_mm, t_act, (r1.float() - r2.float()).abs().max().item()))
    # 64-channel ResNet shapes on 6x6
    x6 = torch.randn(N, 64, 6, 6, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w6 = (torch.randn(64, 64, 3, 3, device=dev, dtype=torch.bfloat16) * 0.05).contiguous(memory_format=torch.channels_last)
    b6 = torch.randn(64, device=dev, dtype=torch.bfloat16)
    print(N, "resnet 64ch 6x6: conv %.1f | conv_relu %.1f | conv_add_relu %.1f" % (
        timeit(lambda: F.conv2d(x6, w6, None, padding=1)),
        timeit(lambda: torch.cudnn_convolution_relu(x6, w6, b6, (1, 1), (1, 1), (1, 1), 1)),
        timeit(lambda: torch.cudnn_convolution_add_relu(x6, w6, x6, 1.0, b6, (1, 1), (1, 1), (1, 1), 1))))
