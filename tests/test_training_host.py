"""CPU: the training-step host logic (cfg 5) -- the differentiable graph against gradients produced by the
reference itself (tests/golden/train_*.npz, written by oracle/make_golden_train.py), the flat parameter/gradient
layout, and the bucketed gradient all-reduce on a 2-rank gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import yaml

from helpers import GOLDEN
from turtlevsr_b200.archs import create_video_model
from turtlevsr_b200.training import FlatAdamW, FlatParams, GradBuckets, LossScaler, TrainStep, autograd_forward


def load_train_case(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    opt = yaml.safe_load(str(z["opt_yaml"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    net = create_video_model(opt)
    net.load_state_dict(sd, strict=True)
    return net.train(), torch.from_numpy(z["lq"]), torch.from_numpy(z["gt"]), z


def clip_loss(net, lq, gt):
    k = v = None
    total = 0
    for j in range(lq.shape[1]):
        pre = lq[:, j if j == 0 else j - 1]
        out, k, v = net(torch.stack([pre, lq[:, j]], 1), k, v)          # forward routes to the autograd graph
        total = total + torch.nn.functional.l1_loss(out, gt[:, j])
    return total / lq.shape[1]


@pytest.mark.parametrize("case", ["train_tiny_t0.npz", "train_tiny_t1.npz"])
def test_autograd_graph_matches_reference_loss_and_gradients(case):
    net, lq, gt, z = load_train_case(case)
    loss = clip_loss(net, lq, gt)
    assert abs(loss.item() - float(z["losses"][0])) < 2e-6
    loss.backward()
    named = dict(net.named_parameters())
    names = [str(n) for n in z["grad_names"]]
    assert names == list(named)
    for n, (s, a) in zip(names, z["grad_digest"]):
        g = named[n].grad
        if g is None:                      # the reference's `0 * sum(p.sum())` term gives untouched weights a zero grad
            assert a == 0.0, n
            continue
        assert abs(float(g.double().abs().sum()) - a) <= 2e-4 * max(a, 1e-3), n
        assert abs(float(g.double().sum()) - s) <= 2e-4 * max(a, 1e-3), n
    for key in z.files:
        if key.startswith("g::"):
            want = torch.from_numpy(z[key])
            got = named[key[3:]].grad
            assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()), key


def test_flat_params_are_views_and_grads_accumulate_in_place():
    net, lq, gt, _ = load_train_case("train_tiny_t0.npz")
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    flat = FlatParams(net)
    assert flat.numel % FlatParams.ALIGN == 0 and all(o % FlatParams.ALIGN == 0 for o in flat.offsets)
    for (n, p), o in zip(net.named_parameters(), flat.offsets):
        assert torch.equal(p, before[n]) and p.data_ptr() == flat.data.data_ptr() + 4 * o
    clip_loss(net, lq[:, :2], gt[:, :2]).backward()
    assert flat.grad.abs().sum() > 0
    for p, o in zip(flat.params, flat.offsets):
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o       # autograd accumulated into the views
    flat.zero_grad()
    assert flat.grad.abs().sum() == 0


def test_bucket_ranges_tile_the_flat_buffer_in_reverse_order():
    net, *_ = load_train_case("train_tiny_t1.npz")
    flat = FlatParams(net)
    b = GradBuckets(flat, bucket_bytes=64 << 10)
    assert len(b.ranges) > 3
    assert b.ranges[0][1] == flat.numel and b.ranges[-1][0] == 0
    for (lo, hi), (lo2, hi2) in zip(b.ranges[1:], b.ranges[:-1]):
        assert hi == lo2 and lo < hi                                     # contiguous, descending
    assert sum(b.members) == len(flat.params)
    assert b.bucket_of[-1] == 0 and b.bucket_of[0] == len(b.ranges) - 1  # last-constructed parameter reduces first


def test_loss_scaler_schedule():
    s = LossScaler(init_scale=8.0, growth_interval=3)
    s.update(True)
    assert s.scale == 4.0
    for _ in range(3):
        s.update(False)
    assert s.scale == 8.0 and s.clean == 0


def test_optimizer_has_no_cpu_path():
    net, *_ = load_train_case("train_tiny_t0.npz")
    with pytest.raises(RuntimeError):
        FlatAdamW(FlatParams(net))
    with pytest.raises(RuntimeError):
        TrainStep(net)


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    net, lq, gt, _ = load_train_case("train_tiny_t0.npz")
    flat = FlatParams(net)
    buckets = GradBuckets(flat, bucket_bytes=64 << 10)
    flat.zero_grad()
    clip_loss(net, lq[rank:rank + 1, :2], gt[rank:rank + 1, :2]).backward()     # each rank: its own sample
    launched_in_backward = sum(buckets.launched)
    buckets.finish()
    mean = flat.grad * buckets.scale
    # the logging reduce of BM:340-365: rank r contributes r + 1, rank 0 ends up with the mean
    red = buckets.reduce_loss(torch.tensor(float(rank + 1)))
    if rank == 0:
        assert abs(red.item() - (world + 1) / 2) < 1e-6
        q.put((mean.clone(), launched_in_backward, len(buckets.ranges), buckets.bytes_reduced))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, early, nb, nbytes = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, both samples in one batch: the L1 mean over the batch == mean of the per-sample losses
    net, lq, gt, _ = load_train_case("train_tiny_t0.npz")
    flat = FlatParams(net)
    clip_loss(net, lq[:, :2], gt[:, :2]).backward()
    assert (mean - flat.grad).abs().max().item() <= 1e-6 * max(1.0, flat.grad.abs().max().item())
    assert nbytes == 4 * flat.numel                                   # every gradient crossed the wire exactly once
    assert 0 < early <= nb                                            # buckets were launched from the grad hooks
