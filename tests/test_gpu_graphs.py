"""B200: CUDA-graph replay of steady-state frames (model.enable_cuda_graphs()) against the eager launch path."""
import pytest
import torch

from helpers import load_case
from turtlevsr_b200.archs import create_video_model
from turtlevsr_b200.clip import run_clip
from turtlevsr_b200.history import RING_PERIOD

pytestmark = pytest.mark.gpu


def _nets(case, mode):
    opt, sd, clip, _, _ = load_case(case)
    nets = []
    for _ in range(2):
        n = create_video_model(opt)
        n.load_state_dict(sd, strict=True)
        nets.append(n.cuda().eval().set_precision(mode))
    return nets, clip


@pytest.mark.parametrize("case,mode", [("tiny_t1_live.npz", "tf32"), ("tiny_t1_live.npz", "fp32"),
                                       ("tiny_super_live.npz", "tf32")])
def test_graph_replay_is_bit_identical_to_eager(case, mode):
    (eager, graphed), clip = _nets(case, mode)
    graphed.enable_cuda_graphs()
    T = 5 * RING_PERIOD + 3
    g = torch.Generator().manual_seed(5)
    long_clip = torch.rand(1, T, *clip.shape[2:], generator=g).cuda()
    want, kw, vw = run_clip(eager, long_clip)
    got, kg, vg = run_clip(graphed, long_clip)
    assert torch.equal(got, want)
    for a, b in zip(list(kw) + list(vw), list(kg) + list(vg)):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape and torch.equal(a, b)
    eng = graphed._engine
    assert 1 <= eng.graph_captures <= RING_PERIOD + 1          # one start-up state + the steady cycle
    assert eng.graph_replays - eng.graph_captures >= 2 * RING_PERIOD
    assert eager._engine.graph_captures == 0


def test_two_interleaved_histories_and_foreign_caches_with_graphs_enabled():
    """Independent histories (tiles / clips) get their own graphs; caches that are not the ring's current window
    (clones, as after a .cpu() round trip, INF:227-237) fall back to an eager frame and still give the same result."""
    (eager, graphed), clip = _nets("tiny_t1_live.npz", "tf32")
    graphed.enable_cuda_graphs()
    T = 3 * RING_PERIOD + 6
    g = torch.Generator().manual_seed(6)
    clips = [torch.rand(1, T, *clip.shape[2:], generator=g).cuda() for _ in range(2)]
    want = [run_clip(eager, c)[0] for c in clips]
    state = [(None, None), (None, None)]
    outs = [[], []]
    with torch.no_grad():
        for j in range(T):
            for i, c in enumerate(clips):
                k, v = state[i]
                if i == 1 and j == T - 3:                  # break the zero-copy chain once
                    k = [None if t is None else t.clone() for t in k]
                    v = [None if t is None else t.clone() for t in v]
                o, k, v = graphed(torch.stack([c[:, j if j == 0 else j - 1], c[:, j]], 1), k, v)
                outs[i].append(o)
                state[i] = (k, v)
    for i in range(2):
        assert torch.equal(torch.stack(outs[i], 1), want[i])
    assert graphed._engine.graph_captures >= 2 * RING_PERIOD          # both histories were captured separately
