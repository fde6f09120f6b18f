import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n*1e3
N=4096
for (cin,H,pad) in ((256,4,1),(256,4,0),(3,4,1),(8,4,1),(64,6,1)):
    cout = 256 if cin!=64 else 64
    x=torch.randn(N,cin,H,H,device='cuda',dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w=torch.randn(cout,cin,3,3,device='cuda',dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b=torch.randn(cout,device='cuda',dtype=torch.bfloat16)
    Ho=H if pad else H-2
    z=torch.randn(N,cout,Ho,Ho,device='cuda',dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t0=timeit(lambda: F.conv2d(x,w,None,padding=pad))
    t1=timeit(lambda: F.conv2d(x,w,b,padding=pad))
    try: t2=timeit(lambda: torch.cudnn_convolution_relu(x,w,b,[1,1],[pad,pad],[1,1],1))
    except Exception as e: t2=str(e)[:80]
    try: t3=timeit(lambda: torch.cudnn_convolution_add_relu(x,w,z,1.0,b,[1,1],[pad,pad],[1,1],1))
    except Exception as e: t3=str(e)[:80]
    print(f"cin={cin} H={H} pad={pad}: conv {t0:.1f} us, conv+bias {t1:.1f}, cudnn_conv_relu {t2}, cudnn_conv_add_relu {t3}", flush=True)
# linear fused
x=torch.randn(N,1024,device='cuda',dtype=torch.bfloat16); w=torch.randn(1024,512,device='cuda',dtype=torch.bfloat16); b=torch.randn(512,device='cuda',dtype=torch.bfloat16)
print("mm", timeit(lambda: torch.mm(x,w)), "addmm", timeit(lambda: torch.addmm(b,x,w)), "addmm_act", timeit(lambda: torch._addmm_activation(b,x,w)))
