"""CPU: host-side logic of the drop-in -- parameter tree, cache protocol, history rings, sharding."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_case
from oracle.turtle_oracle import ArchSpec, Oracle
from turtlevsr_b200.archs import create_video_model
from turtlevsr_b200.clip import shard_clips
from turtlevsr_b200.configs import shipped
from turtlevsr_b200.history import FhrRing, SabRing, resolve_ring


def test_state_dict_schema_full_config():
    net = create_video_model(shipped("Turtle_Deblur_Gopro"))
    sd = net.state_dict()
    assert len(sd) == 633
    assert sum(v.numel() for v in sd.values()) == 59_079_548
    assert sd["decoder_level1.transformer_blocks.1.attn.spatial_aligner.k2_dwconv.weight"].shape == (128, 1, 16, 16)
    assert sd["latent.transformer_blocks.0.attn.temperature"].shape == (8, 1, 1)
    assert "encoder_level1.transformer_blocks.0.attn.beta" in sd and "ending.bias" in sd
    # checkpoints saved from DDP carry a "module." prefix that the reference loader strips (BM:281-284)
    net.load_state_dict({k: v for k, v in sd.items()}, strict=True)


def test_davis_aliases_and_bad_types():
    net = create_video_model(shipped("Turtle_Denoise_Davis"))      # MEST/CTS names build
    assert net.decoder_level3.transformer_blocks[-1].attention_type == "CHM"
    assert net.latent.transformer_blocks[0].attention_type == "FHR"
    bad = shipped("Turtle_Deblur_Gopro")
    bad["encoder1_attn_type1"] = "Nope"
    with pytest.raises(SystemExit):
        create_video_model(bad)
    for name, cls in [("Turtle_Derain", "Turtle"), ("Turtle_SR_MVSR", "TurtleSuper_t1")]:
        assert type(create_video_model(shipped(name))).__name__ == cls
    assert type(create_video_model(shipped("Turtle_Deblur_Gopro"), "SR")).__name__ == "TurtleSuper_t1"


@pytest.mark.parametrize("name", ["tiny_t1_live.npz", "tiny_t0_live.npz", "tiny_super_live.npz"])
def test_cache_protocol_matches_oracle_shapes(name):
    """Dry run of the engine (no kernels): the list-of-8 cache protocol, ring windows growing 1..K and
    output shapes must equal the oracle's for every frame (T1:1045-1132 row a12)."""
    opt, sd, clip, _, _ = load_case(name)
    net = create_video_model(opt)
    net.load_state_dict(sd)
    net.eval()
    net._dry_run = True
    orc = Oracle(ArchSpec.from_opt(opt), sd)
    k = v = ok = ov = None
    for j in range(clip.shape[1]):
        x = torch.stack([clip[:, max(j - 1, 0)], clip[:, j]], 1)
        out, k, v = net(x, k, v)
        oo, ok, ov = orc.forward(x, ok, ov)
        assert out.shape == oo.shape
        assert len(k) == 8 and len(v) == 8
        for a, b in zip(k + v, ok + ov):
            assert (a is None) == (b is None)
            if a is not None:
                assert a.shape == b.shape
    launches = net._engine.launch_log
    assert "turtle_gemm" in launches and "turtle_sab_aggregate" in launches


def test_sab_ring_window_and_compaction():
    r = SabRing(B=1, N=4, Dk=2, Dv=3, keep=3, device="cpu", slots=8)
    seen = []
    for t in range(30):
        slot = r.begin_push()
        first, F = r.first_live, r.count + 1
        assert slot - first + 1 == F and F == min(t + 1, 4)
        # history frames the kernels would read are the last <=3 pushed, oldest first
        assert [float(r.kbuf[0, s, 0, 0]) for s in range(first, slot)] == seen[-3:]
        r.kbuf[:, slot] = float(t)
        r.vbuf[:, slot] = float(t) + 0.5
        seen.append(float(t))
        r.commit()
        k, v = r.views()
        assert k.shape == (1, min(t + 1, 3), 1, 4, 2) and v.shape == (1, min(t + 1, 3), 1, 4, 3)
        assert k[0, :, 0, 0, 0].tolist() == seen[-3:]
        assert resolve_ring(k, v) is r                       # handed-out window is recognised ...
        assert resolve_ring(k.clone(), v.clone()) is None    # ... copies are foreign tensors
    assert r.slots == 8


def test_fhr_ring_views_are_reference_shaped():
    heads, ch, P = 2, 4, 5
    r = FhrRing(B=1, P=P, heads=heads, ch=ch, keep=3, device="cpu", slots=8)
    frames = []
    for t in range(20):
        slot = r.begin_push()
        kf = torch.randn(P, heads * ch)
        frames.append(kf)
        # a kernel writes frame `slot` of head h at columns [slot*ch, (slot+1)*ch) of that head
        r.kbuf[0, :, :, slot * ch:(slot + 1) * ch] = kf.view(P, heads, ch)
        r.vbuf[0, :, :, slot * ch:(slot + 1) * ch] = kf.view(P, heads, ch) * 2
        r.commit()
        k, v = r.views()
        n = min(t + 1, 3)
        assert k.shape == (1, heads, n * ch, P)
        want = torch.stack(frames[-n:], 0).view(n, P, heads, ch).permute(2, 0, 3, 1).reshape(heads, n * ch, P)
        assert torch.equal(k[0], want) and torch.equal(v[0], want * 2)
    adopted = FhrRing.adopt(k.clone(), v.clone(), keep=3, ch=ch, device="cpu")
    k2, _ = adopted.views()
    assert torch.equal(k2, k)


def test_shard_clips_partition():
    for n, w in [(8, 8), (10, 4), (3, 8), (100, 2)]:
        got = [shard_clips(n, r, w) for r in range(w)]
        assert sorted(i for g in got for i in g) == list(range(n))
        assert all(i % w == r for r, g in enumerate(got) for i in g)       # VRM:162-164


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_clips(5, rank, world)
    # per-rank "frames processed" and "elapsed": job throughput = sum(frames) / max(elapsed), as bench.py does
    t = torch.tensor([float(len(mine)) * 3, 10.0 + rank], dtype=torch.float64)
    frames = t[:1].clone()
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    elapsed = t[1:].clone()
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((frames.item(), elapsed.item(), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_clip_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    frames, elapsed, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames == 15.0 and elapsed == 11.0
    assert sorted(gathered[0] + gathered[1]) == [0, 1, 2, 3, 4]
