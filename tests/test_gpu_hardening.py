"""GPU: the drop-in contract under misuse (ADVICE r01 / VERDICT r01 weak item 5): parameters updated in place between
frames, caches that went stale or belong to another geometry, workspace growth while graphs are cached."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_case  # noqa: E402
from turtlevsr_b200.archs import create_video_model  # noqa: E402
from turtlevsr_b200.clip import run_clip  # noqa: E402
from turtlevsr_b200.history import RING_PERIOD, StaleCacheError  # noqa: E402


def build(opt, sd, precision="tf32"):
    net = create_video_model(opt)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval().set_precision(precision)


def test_in_place_parameter_update_between_frames_is_picked_up():
    """An optimizer step followed by eval() validation (VRM:110-129) mutates parameters in place: the packed fp16 / tap-major
    copies and the captured graphs must not survive it."""
    opt, sd, clip, _, _ = load_case("tiny_t1_live.npz")
    net = build(opt, sd)
    net.enable_cuda_graphs()
    g = torch.Generator().manual_seed(3)
    long_clip = torch.rand(1, 3 * RING_PERIOD + 2, *clip.shape[2:], generator=g).cuda()
    run_clip(net, long_clip)                                   # packs weights, captures graphs
    with torch.no_grad():
        for p in net.parameters():                             # what torch.optim.AdamW.step does: in-place update
            p.mul_(1.01).add_(0.001)
    got, _, _ = run_clip(net, long_clip)
    fresh = build(opt, {k: v.detach().clone() for k, v in net.state_dict().items()})
    want, _, _ = run_clip(fresh, long_clip)
    assert torch.equal(got, want)


def test_stale_cache_view_raises_and_clone_replays():
    opt, sd, clip, _, _ = load_case("tiny_t1_live.npz")
    net = build(opt, sd, "fp32")
    g = torch.Generator().manual_seed(4)
    c = torch.rand(1, RING_PERIOD + 6, *clip.shape[2:], generator=g).cuda()
    k = v = None
    kept = None
    outs = []
    for j in range(c.shape[1]):
        x = torch.stack([c[:, max(j - 1, 0)], c[:, j]], 1)
        o, k, v = net(x, k, v)
        outs.append(o)
        if j == 2:
            kept = (list(k), list(v))                                            # views of ring memory
            cloned = ([None if t is None else t.clone() for t in k], [None if t is None else t.clone() for t in v])
    x3 = torch.stack([c[:, 2], c[:, 3]], 1)
    with pytest.raises(StaleCacheError):                       # the rings have compacted since frame 2
        net(x3, kept[0], kept[1])
    o3, _, _ = net(x3, cloned[0], cloned[1])                   # a clone is an ordinary foreign tensor: replay works
    assert torch.equal(o3, outs[3])


def test_cache_geometry_mismatch_raises_like_the_reference():
    opt, sd, clip, _, _ = load_case("tiny_t1_live.npz")
    net = build(opt, sd, "fp32")
    x = torch.rand(1, 2, 3, 64, 96).cuda()
    _, k, v = net(x, None, None)
    k = [None if t is None else t.clone() for t in k]
    v = [None if t is None else t.clone() for t in v]
    with pytest.raises(ValueError):                            # another resolution: the reference fails in torch.cat
        net(torch.rand(1, 2, 3, 32, 64).cuda(), k, v)
    with pytest.raises(ValueError):                            # another batch size
        net(torch.rand(2, 2, 3, 64, 96).cuda(), k, v)


def test_workspace_growth_drops_graphs_instead_of_replaying_freed_memory():
    opt, sd, clip, _, _ = load_case("tiny_t1_live.npz")
    net = build(opt, sd)
    net.enable_cuda_graphs()
    ref = build(opt, sd)
    g = torch.Generator().manual_seed(5)
    small = torch.rand(1, 3 * RING_PERIOD, 3, 32, 64, generator=g).cuda()
    big = torch.rand(1, 4, 3, 96, 128, generator=g).cuda()
    want_a, _, _ = run_clip(ref, small)
    k = v = None
    outs = []
    for j in range(2 * RING_PERIOD + 1):                       # graphs captured for the small clip
        o, k, v = net(torch.stack([small[:, max(j - 1, 0)], small[:, j]], 1), k, v)
        outs.append(o)
    assert net._engine.graph_captures > 0
    run_clip(net, big)                                         # every workspace buffer grows: old graphs must go
    for j in range(2 * RING_PERIOD + 1, small.shape[1]):       # continue the small clip with its caches
        o, k, v = net(torch.stack([small[:, j - 1], small[:, j]], 1), k, v)
        outs.append(o)
    assert torch.equal(torch.stack(outs, 1), want_a)
