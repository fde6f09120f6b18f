"""GPU numerics of the leaf-evaluation pipeline: fused inference plans (library convs/GEMMs + the
engine's fused epilogue / heads kernels) against the plain PyTorch fp32 modules in eval mode.
Tolerances: fp32 plan 2e-5 abs on probabilities/values (north_star: Q/priors within 1e-5 relative is
for the TREE given identical net outputs; the net itself is floating point and its plan re-associates
sums); bf16 plans: BF16_BOUNDS below (relative error of every prior above 1e-3, per-row KL, tanh value) against the
fp32 module on bf16-rounded weights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.2)


def _rounded_copy(model, dtype):
    """The fp32 module with every conv / linear weight rounded to `dtype` (what a 16-bit plan can represent at best)."""
    import copy
    m = copy.deepcopy(model)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):
                mod.weight.copy_(mod.weight.to(dtype).float())
    return m


def _net_errors(p, v, p_ref, v_ref):
    """max relative error over every prior above 1e-3, largest per-row KL(ref || p), max |v - v_ref|, max abs prior error."""
    big = p_ref > 1e-3
    rel = ((p - p_ref).abs() / p_ref)[big].max().item()
    kl = (p_ref * (torch.log(p_ref.clamp_min(1e-30)) - torch.log(p.clamp_min(1e-30)))).sum(1).max().item()
    return {"rel": rel, "kl": kl, "v": (v - v_ref).abs().max().item(), "abs": (p - p_ref).abs().max().item()}


def _record(name, errs):
    """Achieved errors go to gpurun_out/net_errors.json (scratch) so that the bounds below can be read against them."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "net_errors.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[name] = errs
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


# bf16 bounds per net, every one about 3x what the plans achieve on a B200 (profiles/r02_net_errors.md; BatchNorm
# statistics and weights are random here, which makes the nets far more sensitive than trained ones): relative error of
# every prior above 1e-3, per-row KL divergence, tanh value.  A wrong fold of one BatchNorm moves the KL by orders of magnitude.
BF16_BOUNDS = {("simple", 3): {"rel": 1.5e-2, "kl": 1e-5, "v": 8e-3}, ("simple", 5): {"rel": 1.5e-2, "kl": 1e-5, "v": 8e-3},
               ("resnet", 3): {"rel": 1e-1, "kl": 3e-4, "v": 2e-2}, ("resnet", 5): {"rel": 3e-1, "kl": 2.5e-3, "v": 5e-2}}
# the plain bf16 module (no fused plan) exponentiates a bf16 log_softmax: looser
MODULE_BOUNDS = {"rel": 1.5e-1, "kl": 2e-2, "v": 3e-2}


@pytest.mark.parametrize("net,board,blocks", [("simple", (3, 3), 0), ("simple", (5, 5), 0), ("resnet", (3, 3), 4), ("resnet", (5, 5), 4),
                                              ("resnet", (3, 3), 20), ("resnet", (5, 5), 20)])
def test_fused_plans_match_torch_modules(net, board, blocks):
    """fp32 plan: 2e-5 abs against the fp32 module.  bf16 plans: against the fp32 module evaluated on bf16-rounded
    weights -- a relative bound on every prior above 1e-3, a per-row KL bound and an absolute bound on the tanh value
    (BF16_BOUNDS); ResNetZero also at its full depth of 20 blocks (configuration.py:133-155), with the tcgen05 tower
    kernel and with the library convolutions."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dotsboxesaz_b200 import engine
    from dotsboxesaz_b200.nn import (DeviceEvaluator, FusedResNetZero, FusedSimpleNN, ResNetZero, resnet_zero_parameters)
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    from dotsboxesaz_b200.utils.utils import DotDict
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = 512
    eng = engine.Engine(board, n_games=n, max_nodes=64)
    torch.manual_seed(1)
    if net == "simple":
        model = SimpleNN(board=board)
        plan_cls = FusedSimpleNN
    else:
        model = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters(board, nb_blocks=blocks)}}))
        plan_cls = FusedResNetZero
    _randomise_bn(model, 2)
    model = model.cuda().eval()
    # leaves of a real search as inputs
    st = eng.new_states(n)
    plies, _ = eng.random_rollout(st.clone(), seed=3, record_moves=True)
    for ply in range(6):
        legal = eng.valid_moves(st).float()
        mv = torch.multinomial(legal + 1e-9, 1).reshape(-1).int()
        eng.play(st, torch.where(torch.arange(n, device=st.device) % 7 > ply, mv, torch.full_like(mv, -1)))
    x32 = eng.features(st, torch.float32)
    with torch.no_grad():
        logp, v = model(x32)
        logp16, v16 = _rounded_copy(model, torch.bfloat16)(x32)
    refs = {torch.float32: (torch.exp(logp), v.reshape(-1)), torch.bfloat16: (torch.exp(logp16), v16.reshape(-1))}
    tag = "%s-%dx%d-%d" % (net, board[0], board[1], blocks)

    def check(name, dtype):
        torch.cuda.synchronize()
        p_ref, v_ref = refs[dtype]
        assert torch.isfinite(eng.priors).all()
        # the plans' heads kernel normalises in float32; the plain bf16 module exponentiates a bf16 log_softmax
        assert (eng.priors.sum(1) - 1).abs().max() < (1e-3 if name.startswith("plan") else 1e-2)
        errs = _net_errors(eng.priors, eng.values, p_ref, v_ref)
        _record("%s %s" % (tag, name), errs)
        if dtype == torch.float32:
            assert errs["abs"] < 2e-5 and errs["v"] < 2e-5, (tag, name, errs)
        else:
            bounds = BF16_BOUNDS[(net, board[0])] if name.startswith("plan") else MODULE_BOUNDS
            for k, bound in bounds.items():
                assert abs(errs[k]) < bound, (tag, name, k, errs)

    for dtype in (torch.float32, torch.bfloat16):
        dn = "fp32" if dtype == torch.float32 else "bf16"
        plan = plan_cls(model, eng, dtype=dtype)
        eng.planes.copy_(x32.to(dtype))
        eng.leaf_states.copy_(st)
        plan(eng)
        check("plan " + dn, dtype)
        if dtype != torch.float32:
            assert plan.stem_mma is not None  # the default stem is the tensor-core kernel
            if net == "resnet":
                assert plan.tower is not None  # ... and the default tower the tcgen05 kernel
                plan_lib = plan_cls(model, eng, dtype=dtype, use_tower=False)
                assert plan_lib.tower is None
                plan_lib(eng)
                check("plan bf16, library convolutions", dtype)
            if blocks <= 4:  # the library-conv0 and the CUDA-core-stem variants of the same plan must agree as well
                for variant in (False, "fma"):
                    plan2 = plan_cls(model, eng, dtype=dtype, use_stem=variant)
                    assert plan2.stem_mma is None
                    eng.planes.copy_(x32.to(dtype))
                    plan2(eng)
                    check("plan bf16, stem %s" % variant, dtype)
        if blocks <= 4:
            # the un-fused evaluator (plain module on the same planes) must agree too
            ev = DeviceEvaluator(model, eng, dtype=dtype, channels_last=True)
            eng.planes.copy_(x32.to(dtype))
            ev(eng)
            check("module " + dn, dtype)
    eng.close()


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_epilogue_kernel_matches_torch(dtype):
    """dbaz_nn_epilogue (bias + ReLU + eval-BatchNorm in one pass; the three modes of include/dbaz_b200.h) against the
    same expression in PyTorch fp32.  Tolerance: one rounding of the storage dtype (bf16: 2^-8 relative)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dotsboxesaz_b200 import engine
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    eng = engine.Engine((3, 3), n_games=4, max_nodes=8)
    g = torch.Generator(device="cuda").manual_seed(0)
    rows, ch = 1000, 64
    x = torch.randn(rows, ch, device="cuda", generator=g)
    res = torch.randn(rows, ch, device="cuda", generator=g)
    bias, scale, shift = (torch.randn(ch, device="cuda", generator=g) for _ in range(3))
    for mode in (0, 1, 2):
        for use_res in ((False, True) if mode == 1 else (False,)):
            xr, rr = x.to(dt).float(), res.to(dt).float()
            if mode == 0:
                ref = scale * torch.relu(xr + bias) + shift
            elif mode == 1:
                ref = torch.relu(scale * (xr + bias) + shift + (rr if use_res else 0))
            else:
                ref = scale * (xr + bias) + shift
            y = x.to(dt).clone()
            eng.nn_epilogue(y, bias, scale, shift, mode=mode, res=res.to(dt) if use_res else None)
            tol = 2 ** -7 if dtype == "bf16" else 1e-5
            assert ((y.float() - ref).abs() <= tol * (1 + ref.abs())).all(), (dtype, mode, use_res)
    eng.close()


@pytest.mark.parametrize("board,cout,dtype", [((3, 3), 256, "bf16"), ((5, 5), 64, "bf16"), ((2, 3), 128, "fp16"), ((4, 6), 512, "bf16")])
def test_stem_mma_kernel_matches_conv(board, cout, dtype):
    """dbaz_nn_stem_mma (leaf gather + conv0 + folded affines + ReLU as one tensor-core implicit GEMM from packed states)
    against F.conv2d in fp32 on the feature planes of the same states, with the weights rounded to the storage dtype as
    the kernel sees them.  Tolerance: fp32 accumulation order + one output rounding (2^-8 relative for bf16)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.nn.functional as F
    from dotsboxesaz_b200 import engine
    from dotsboxesaz_b200.nn import _stem_mma_table
    dt = torch.bfloat16 if dtype == "bf16" else torch.float16
    n = 777  # not a multiple of the 16-row tile
    eng = engine.Engine(board, n_games=4, max_nodes=8)
    g = torch.Generator(device="cuda").manual_seed(5)
    st = eng.new_states(n)
    for ply in range(9):
        legal = eng.valid_moves(st).float()
        mv = torch.multinomial(legal + 1e-9, 1, generator=g).reshape(-1).int()
        eng.play(st, torch.where((torch.arange(n, device="cuda") % 10 > ply) & (legal.sum(1) > 0), mv, torch.full_like(mv, -1)))
    conv = torch.nn.Conv2d(3, cout, 3, padding=1).cuda()
    s_in, t_in = torch.rand(3, device="cuda", generator=g) + 0.5, torch.randn(3, device="cuda", generator=g) * 0.3
    s_o, t_o = torch.rand(cout, device="cuda", generator=g) + 0.5, torch.randn(cout, device="cuda", generator=g) * 0.2
    tab = _stem_mma_table(conv, s_in, t_in, s_o, t_o).to(dt)
    out = torch.empty((n, eng.rows, eng.cols, cout), dtype=dt, device="cuda")
    eng.nn_stem_mma(st, eng.nn_stem_mma_pack(tab), out)
    # reference from the ROUNDED table: un-fold it into a conv over [plane0, plane1, plane2, in-board indicator]
    t32 = tab.float()
    w = torch.zeros((cout, 4, 3, 3), device="cuda")
    w[:, :2] = t32[0:18].reshape(2, 3, 3, cout).permute(3, 0, 1, 2)
    w[:, 2] = t32[18:27].reshape(3, 3, cout).permute(2, 0, 1)
    w[:, 3] = t32[27:36].reshape(3, 3, cout).permute(2, 0, 1)
    x = eng.features(st, torch.float32)
    x4 = torch.cat([x, torch.ones_like(x[:, :1])], 1)
    ref = F.relu(F.conv2d(x4, w, t32[36], padding=1)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert (err <= 2 ** -7 * (1 + ref.abs())).all(), err.max().item()
    # and the fold itself: relu(s_o * (conv(s_in * x + t_in) + b) + t_o) in fp32, up to the rounding of the table
    with torch.no_grad():
        full = F.relu(s_o.view(1, -1, 1, 1) * conv(x * s_in.view(1, 3, 1, 1) + t_in.view(1, 3, 1, 1)) + t_o.view(1, -1, 1, 1))
    assert (out.float() - full.permute(0, 2, 3, 1)).abs().max().item() < 0.25
    eng.close()
