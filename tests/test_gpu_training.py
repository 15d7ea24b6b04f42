"""B200: the training step of cfg 5 -- flat AdamW / non-finite-check kernels through the C ABI against
torch.optim.AdamW, and two full steps against parameters produced by the reference (tests/golden/train_*.npz)."""
import numpy as np
import pytest
import torch

from test_training_host import load_train_case
from turtlevsr_b200 import capi
from turtlevsr_b200.training import FlatAdamW, FlatParams, TrainStep

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OPTIM = dict(type="Adam", lr=4e-4, weight_decay=0, betas=[0.9, 0.99])         # Turtle_Derain.yml:90-94


@pytest.fixture(autouse=True)
def true_fp32_library_kernels():
    """The autograd graph runs on cuDNN/cuBLAS: for parity with the reference's CPU fp32 gradients switch their TF32
    paths off ("oracle numerics")."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("wd,n", [(0.0, 1 << 20), (0.05, 1000003)])
def test_adamw_flat_kernel_matches_torch_adamw(wd, n):
    g = torch.Generator(device=DEV).manual_seed(5)
    p0 = torch.randn(n, device=DEV, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    ref = torch.optim.AdamW([ref_p], lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    lin = torch.nn.Linear(1, 1, bias=False).to(DEV)
    lin.weight = torch.nn.Parameter(p0.clone().view(n, 1))
    flat = FlatParams(lin)
    opt = FlatAdamW(flat, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    for step in range(4):
        grad = torch.randn(n, device=DEV, generator=g) * (10.0 ** (step - 2))
        ref_p.grad = grad.clone()
        ref.step()
        flat.grad[:n].copy_(grad * 128.0)                       # as if the loss had been scaled by 128 on 1 rank
        opt.step(grad_scale=1.0 / 128.0)
        assert (flat.data[:n] - ref_p.detach()).abs().max().item() < 2e-6
    assert opt.steps == 4
    st = ref.state[ref_p]
    assert torch.allclose(opt.exp_avg[:n], st["exp_avg"], rtol=1e-5, atol=1e-6)   # torch lerps, the kernel does b1*m+(1-b1)*g
    assert torch.allclose(opt.exp_avg_sq[:n], st["exp_avg_sq"], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,C,H,W", [(2, 64, 17, 33), (1, 512, 8, 8), (3, 8, 5, 300)])
def test_channel_layernorm_forward_backward_kernels(dtype, tol, B, C, H, W):
    """turtle_ln2d_fwd / turtle_ln2d_bwd behind autograd vs the reference's formula (T1:96-112) in float64."""
    from turtlevsr_b200.training import _ChannelLayerNorm
    g = torch.Generator(device=DEV).manual_seed(B * C + H)
    x = (torch.randn(B, C, H, W, device=DEV, generator=g) * 1.7 + 0.4).to(dtype).requires_grad_()
    w = (torch.rand(C, device=DEV, generator=g) + 0.5).requires_grad_()
    b = torch.randn(C, device=DEV, generator=g).requires_grad_()
    dy = torch.randn(B, C, H, W, device=DEV, generator=g)
    y = _ChannelLayerNorm.apply(x, w, b)
    assert y.dtype == torch.float32
    y.backward(dy)
    xd = x.detach().double().requires_grad_()
    wd, bd = w.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    mu = xd.mean(1, keepdim=True)
    var = (xd - mu).pow(2).mean(1, keepdim=True)
    yd = (xd - mu) / torch.sqrt(var + 1e-5) * wd.view(1, -1, 1, 1) + bd.view(1, -1, 1, 1)
    yd.backward(dy.double())
    assert (y.double() - yd).abs().max() < 1e-5                      # x is the same (already rounded) tensor on both sides
    assert (x.grad.double() - xd.grad).abs().max() < tol * max(1.0, xd.grad.abs().max().item())
    n = B * H * W
    assert (w.grad.double() - wd.grad).abs().max() < 1e-5 * n ** 0.5 * 4
    assert (b.grad.double() - bd.grad).abs().max() < 1e-5 * n ** 0.5 * 4
    y2 = _ChannelLayerNorm.apply(x, w, b)                            # deterministic weight gradients
    w.grad = None
    y2.backward(dy)
    w1 = w.grad.clone()
    w.grad = None
    _ChannelLayerNorm.apply(x, w, b).backward(dy)
    assert torch.equal(w1, w.grad)


@pytest.mark.parametrize("od", [torch.float16, torch.bfloat16])
def test_channel_layernorm_cast_variants(od):
    """y written / dy read in the autocast dtype == the fp32 kernel followed by ATen's casts (one rounding either way)."""
    from turtlevsr_b200.training import _ChannelLayerNorm
    B, C, H, W = 2, 48, 16, 24
    g = torch.Generator(device=DEV).manual_seed(5)
    x = (torch.randn(B, C, H, W, device=DEV, generator=g) * 1.3 - 0.2).requires_grad_()
    w = (torch.rand(C, device=DEV, generator=g) + 0.5).requires_grad_()
    b = torch.randn(C, device=DEV, generator=g).requires_grad_()
    dy = torch.randn(B, C, H, W, device=DEV, generator=g).to(od)
    y = _ChannelLayerNorm.apply(x, w, b, od)
    assert y.dtype == od
    y.backward(dy)
    got = (y.detach().clone(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    y32 = _ChannelLayerNorm.apply(x, w, b)
    y32.to(od).backward(dy)
    assert torch.equal(got[0], y32.detach().to(od))
    assert torch.equal(got[1], x.grad) and torch.equal(got[2], w.grad) and torch.equal(got[3], b.grad)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,C,H,W,bias", [(2, 24, 37, 70, True), (1, 640, 32, 32, False), (3, 5, 64, 33, True),
                                            (2, 6, 70, 128, True), (1, 9, 24, 64, False), (2, 3, 256, 256, True)])   # W % 64 == 0: two-column kernels
def test_depthwise3x3_forward_backward_kernels(dtype, tol, B, C, H, W, bias):
    """turtle_dwconv3x3_nchw / _wgrad behind autograd vs F.conv2d(groups=C) in float64."""
    import torch.nn.functional as F
    from turtlevsr_b200.training import _Depthwise3x3
    g = torch.Generator(device=DEV).manual_seed(C + H)
    x = torch.randn(B, C, H, W, device=DEV, generator=g).to(dtype).requires_grad_()
    w = (torch.randn(C, 1, 3, 3, device=DEV, generator=g) / 3).requires_grad_()
    b = torch.randn(C, device=DEV, generator=g).requires_grad_() if bias else None
    dy = torch.randn(B, C, H, W, device=DEV, generator=g).to(dtype)
    y = _Depthwise3x3.apply(x, w, b)
    assert y.dtype == dtype
    y.backward(dy)
    xd, wd = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
    bd = b.detach().double().requires_grad_() if bias else None
    yd = F.conv2d(xd, wd, bd, padding=1, groups=C)
    yd.backward(dy.double())
    assert (y.double() - yd).abs().max() < tol * max(1.0, yd.abs().max().item())
    assert (x.grad.double() - xd.grad).abs().max() < tol * max(1.0, xd.grad.abs().max().item())
    n = B * H * W
    assert (w.grad.double() - wd.grad).abs().max() < 2e-5 * n ** 0.5 * 4
    if bias:
        assert (b.grad.double() - bd.grad).abs().max() < 2e-5 * n ** 0.5 * 4
    w1 = w.grad.clone()
    w.grad = None
    _Depthwise3x3.apply(x, w, b).backward(dy)
    assert torch.equal(w1, w.grad)                                   # deterministic weight gradient


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,Ch,H,W", [(2, 5, 8, 12), (1, 96, 32, 32), (3, 7, 64, 40)])
def test_gelu_gate_forward_backward_kernels(dtype, tol, B, Ch, H, W):
    """turtle_gelu_gate_nchw / _bwd behind autograd vs `gelu(a) * g` on the chunk views: float64 reference within the
    dtype's tolerance, and bit-identical to the ATen chain in the same dtype (it rounds where ATen rounds)."""
    import torch.nn.functional as F
    from turtlevsr_b200.training import _GeluGate
    g = torch.Generator(device=DEV).manual_seed(Ch + W)
    u = (torch.randn(B, 2 * Ch, H, W, device=DEV, generator=g) * 1.5).to(dtype).requires_grad_()
    dy = torch.randn(B, Ch, H, W, device=DEV, generator=g).to(dtype)
    y = _GeluGate.apply(u)
    assert y.dtype == dtype and y.shape == (B, Ch, H, W)
    y.backward(dy)
    ud = u.detach().double().requires_grad_()
    a, gg = ud.chunk(2, dim=1)
    yd = F.gelu(a) * gg
    yd.backward(dy.double())
    assert (y.double() - yd).abs().max() < tol * max(1.0, yd.abs().max().item())
    assert (u.grad.double() - ud.grad).abs().max() < 2 * tol * max(1.0, ud.grad.abs().max().item())
    ut = u.detach().clone().requires_grad_()
    a, gg = ut.chunk(2, dim=1)
    yt = F.gelu(a) * gg
    yt.backward(dy)
    if dtype == torch.float32:
        assert (y - yt).abs().max() < 1e-6 and (u.grad - ut.grad).abs().max() < 2e-6
    else:
        # the same intermediate roundings as the ATen chain: at most one ulp of the 16-bit type apart (erff / expf variants)
        ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
        assert ((y.float() - yt.float()).abs() <= ulp * yt.float().abs() + 1e-6).all()
        assert ((u.grad.float() - ut.grad.float()).abs() <= 2 * ulp * ut.grad.float().abs() + 1e-5).all()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 1e-3), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("b,hd,cc,h,w", [(2, 2, 6, 8, 12), (1, 4, 16, 32, 32), (3, 1, 5, 4, 300)])
def test_unit_rows_forward_backward_kernels(dtype, tol, b, hd, cc, h, w):
    """turtle_rownorm_fwd / _bwd behind autograd on the q chunk of a qkv map vs F.normalize in float64."""
    import torch.nn.functional as F
    from turtlevsr_b200.training import _UnitRows, _unit_rows
    c = hd * cc
    g = torch.Generator(device=DEV).manual_seed(c + w)
    qkv = torch.randn(b, 3 * c, h, w, device=DEV, generator=g).to(dtype).requires_grad_()
    dy = torch.randn(b, hd, cc, h * w, device=DEV, generator=g)
    q = qkv.chunk(3, dim=1)[1].reshape(b, hd, cc, h * w)               # the k chunk: a strided view
    assert q.data_ptr() != qkv.data_ptr() and (b == 1 or not q.is_contiguous())
    y = _unit_rows(q)
    assert y.dtype == torch.float32 and isinstance(y.grad_fn, _UnitRows._backward_cls)
    y.backward(dy)
    qd = qkv.detach().double().requires_grad_()
    yd = F.normalize(qd.chunk(3, dim=1)[1].reshape(b, hd, cc, h * w), dim=-1)
    yd.backward(dy.double())
    assert (y.double() - yd).abs().max() < 2e-6
    assert (qkv.grad.double() - qd.grad).abs().max() < tol * max(1.0, qd.grad.abs().max().item())
    z = torch.zeros(1, 1, 2, 8, device=DEV, requires_grad=True)        # zero rows: the eps clamp, no NaN
    yz = _unit_rows(z)
    yz.sum().backward()
    assert torch.equal(yz, torch.zeros_like(yz)) and torch.isfinite(z.grad).all()


def test_non_finite_gradients_skip_the_update():
    lin = torch.nn.Linear(1000, 37).to(DEV)
    flat = FlatParams(lin)
    opt = FlatAdamW(flat, lr=1e-2)
    before = flat.data.clone()
    flat.grad.normal_()
    flat.grad[12345] = float("inf")
    opt.step(check_finite=True)
    assert opt.found_inf.item() == 1.0 and torch.equal(flat.data, before)
    flat.grad[12345] = float("nan")
    opt.step(check_finite=True)
    assert opt.found_inf.item() == 1.0 and torch.equal(flat.data, before)
    flat.grad[12345] = 0.5
    opt.step(check_finite=True)
    assert opt.found_inf.item() == 0.0 and not torch.equal(flat.data, before)


@pytest.mark.parametrize("case", ["train_tiny_t0.npz", "train_tiny_t1.npz"])
def test_two_training_steps_match_reference(case):
    net, lq, gt, z = load_train_case(case)
    net = net.to(DEV)
    n0 = capi.launch_count
    ts = TrainStep(net, OPTIM)
    lq, gt = lq.to(DEV), gt.to(DEV)
    l1 = ts.step(lq, gt).item()
    l2 = ts.step(lq, gt).item()
    assert capi.launch_count - n0 > 2                           # LayerNorm fwd/bwd launches + one AdamW launch per step
    assert abs(l1 - float(z["losses"][0])) < 1e-5
    assert abs(l2 - float(z["losses"][1])) < 2e-5               # second loss is evaluated on the updated weights
    worst = 0.0
    for name, p in net.named_parameters():
        want = torch.from_numpy(z["after2::" + name]).to(DEV)
        worst = max(worst, (p.detach() - want).abs().max().item())
    # measured 2.6e-7 (t0) / 1.5e-7 (t1) on B200 with the library TF32 paths off; one Adam step moves a weight by ~4e-4,
    # so 5e-6 still pins every element's update direction and size
    assert worst <= 5e-6, worst


def test_graph_replayed_steps_match_eager_steps():
    """TrainStep(cuda_graph=True): forward+backward replayed from one CUDA graph == the eager step, step by step."""
    nets = []
    for _ in range(2):
        net, lq, gt, _ = load_train_case("train_tiny_t1.npz")
        nets.append(net.to(DEV))
    lq, gt = lq.to(DEV), gt.to(DEV)
    eager, graphed = TrainStep(nets[0], OPTIM), TrainStep(nets[1], OPTIM, cuda_graph=True)
    g = torch.Generator(device=DEV).manual_seed(3)
    for i in range(TrainStep.GRAPH_WARMUP + 3):
        a = torch.rand(lq.shape, device=DEV, generator=g)
        b = torch.rand(gt.shape, device=DEV, generator=g)
        le, lg = eager.step(a, b).item(), graphed.step(a, b).item()
        assert abs(le - lg) < 1e-6, i
    assert graphed._graph is not None and eager._graph is None
    # same kernels, but cuDNN may pick another algorithm under capture: gradients agree to rounding, and Adam turns a
    # rounding-level gradient difference on a near-zero gradient into a fraction of its lr = 4e-4 step
    for (n, p), (_, q) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        assert (p - q).abs().max().item() < 5e-5, n


def test_autocast_steps_loss_scaler_and_eval_after_training():
    net, lq, gt, z = load_train_case("train_tiny_t1.npz")
    net = net.to(DEV)
    lq, gt = lq.to(DEV), gt.to(DEV)
    # fp16 autocast (the reference's mode, VRM:80): this fixture's live gates overflow fp16 inside the gated FFN, which
    # is exactly what the GradScaler protocol is for -- the non-finite check must skip the update and halve the scale
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    ts = TrainStep(net, OPTIM, amp="fp16")
    losses = [ts.step(lq, gt).item() for _ in range(3)]
    assert ts.opt.steps + ts.skipped_steps == 3
    if not all(np.isfinite(losses)):
        assert ts.skipped_steps == 3 and ts.scaler.scale == 65536.0 / 8
        assert all(torch.equal(p.detach(), before[n]) for n, p in net.named_parameters())
    # bf16 autocast: finite, close to the fp32 loss, and the optimizer advances
    ts.amp, ts.scaler = "bf16", None
    losses = [ts.step(lq, gt).item() for _ in range(2)]
    assert all(np.isfinite(losses)) and abs(losses[0] - float(z["losses"][0])) < 2e-2
    assert ts.opt.steps >= 2
    # the inference engine picks up the updated (flat-buffer) weights: eval forward == autograd graph, no_grad
    net.eval()
    x = torch.stack([lq[:1, 0], lq[:1, 1]], 1)
    with torch.no_grad():
        got, _, _ = net.set_precision("fp32")(x)
        from turtlevsr_b200.training import autograd_forward
        want, _, _ = autograd_forward(net, x)
    assert (got - want).abs().max().item() < 1e-4
