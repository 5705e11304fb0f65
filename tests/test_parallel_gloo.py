"""world_size-2 data-parallel host logic on CPU (gloo): batch sharding, bucketed gradient averaging
with and without the backward-overlap hooks, parameter broadcast.  The reducer is model-agnostic,
so a small stock-torch model stands in for PGUNet (whose kernels need a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ug_pg_unet_b200.parallel import GradBucketReducer, broadcast_module_state, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 8, 3, padding=1),
                               torch.nn.ReLU(), torch.nn.Conv2d(8, 1, 1))


def _worker(rank, world, port, overlap, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(seed=100 + rank)               # ranks start different on purpose
        broadcast_module_state(model, src=0)
        ref = _model(seed=100)                        # what rank 0 had
        for a, b in zip(model.parameters(), ref.parameters()):
            assert torch.equal(a, b)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(8, 3, 16, 16, generator=g)
        y = torch.randn(8, 1, 16, 16, generator=g)
        xs, ys = shard_batch(x, rank, world), shard_batch(y, rank, world)
        assert xs.shape[0] == 4 and torch.equal(xs, x[rank * 4:(rank + 1) * 4])
        # tiny buckets -> several buckets, exercised in reverse parameter order
        reducer = GradBucketReducer(list(model.parameters()), bucket_mb=0.0005, overlap=overlap)
        assert len(reducer.buckets) >= 2
        for step in range(2):                         # two steps: hooks must re-arm
            model.zero_grad(set_to_none=True)
            loss = ((model(xs) - ys) ** 2).mean()     # per-shard mean, as UG_unet.py:92
            loss.backward()
            reducer.wait()
            # oracle: N replicas each fed its shard, gradients averaged
            ref.zero_grad(set_to_none=True)
            total = 0
            for r in range(world):
                total = total + ((ref(shard_batch(x, r, world)) - shard_batch(y, r, world)) ** 2).mean() / world
            total.backward()
            for (n, p), q in zip(model.named_parameters(), ref.parameters()):
                torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-7, msg=lambda m: f"{n}: {m}")
        # average=False: .grad is a view of the bucket holding the SUM over ranks (the optimiser kernel
        # applies 1/world_size), and no per-tensor scale runs after the all-reduce
        reducer.remove()
        summed = GradBucketReducer(list(model.parameters()), bucket_mb=0.0005, overlap=overlap, average=False)
        model.zero_grad(set_to_none=True)
        ((model(xs) - ys) ** 2).mean().backward()
        summed.wait()
        for (n, p), q in zip(model.named_parameters(), ref.parameters()):
            torch.testing.assert_close(p.grad, q.grad * world, rtol=1e-5, atol=1e-7, msg=lambda m: f"{n}: {m}")
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


class _FakePack:
    """Stand-in for engine.PackedWeights with a deferred (packed) gradient slice in a flat arena."""

    def __init__(self, conv, buf):
        self.conv, self._dwp, self.on_grad_ready = conv, buf, None


def _arena_worker(rank, world, port, overlap, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(seed=5)
        convs = [m for m in model if isinstance(m, torch.nn.Conv2d)]
        deferred = list(reversed(convs[:2]))                 # arena order = reverse forward order
        arena = torch.zeros(sum(c.weight.numel() for c in deferred))
        packs, off = [], 0
        for c in deferred:
            n = c.weight.numel()
            packs.append(_FakePack(c, arena[off:off + n]))
            off += n
        reducer = GradBucketReducer(list(model.parameters()), bucket_mb=0.0005, overlap=overlap,
                                    conv_packs=packs, arena=arena)
        assert len(reducer.arena_buckets) == 2
        # deferred weights are not in the `.grad` buckets
        in_buckets = {id(p) for b in reducer.buckets for p in b.params}
        assert all(id(c.weight) not in in_buckets for c in deferred)
        for step in range(2):
            # "backward": rank-dependent packed gradients land in the arena layer by layer
            for i, pk in enumerate(packs):
                pk._dwp.copy_(torch.full_like(pk._dwp, float((rank + 1) * (i + 1) + step)))
                if pk.on_grad_ready is not None:
                    pk.on_grad_ready(pk)
            for p in model.parameters():
                if id(p) not in {id(c.weight) for c in deferred}:
                    p.grad = torch.full_like(p, float(rank + 1))
            if overlap:
                # launched from the hooks - except the light tail bucket, which wait() sends together
                # with the flat small-tensor bucket
                assert all(ab.work is not None for ab in reducer.arena_buckets if ab is not reducer._tail_arena)
                assert reducer._tail_arena is reducer.arena_buckets[-1] and reducer._tail_arena.work is None
            reducer.wait()
            # arena holds the SUM (the optimiser kernel applies 1/world); `.grad`s hold the mean
            for i, pk in enumerate(packs):
                want = sum((r + 1) * (i + 1) + step for r in range(world))
                assert torch.all(pk._dwp == float(want)), (i, pk._dwp[:4], want)
            for p in model.parameters():
                if p.grad is not None:
                    assert torch.allclose(p.grad, torch.full_like(p, sum(range(1, world + 1)) / world))
        reducer.remove()
        assert all(pk.on_grad_ready is None for pk in packs)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_packed_gradient_arena_buckets_world2(tmp_path, overlap):
    port = _free_port()
    mp.spawn(_arena_worker, args=(2, port, overlap, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


@pytest.mark.parametrize("overlap", [True, False])
def test_bucketed_gradient_averaging_world2(tmp_path, overlap):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, overlap, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_batch_requires_equal_shards():
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(5, 1), 0, 2)


def _herlev_worker(rank, world, port, out_dir):
    """Data-parallel HerlevTrainer host logic (no kernels): epoch metrics identical on every rank, reducer
    rebuilt when requires_grad flags change (ranks must not pick different learning rates or
    leave the epoch loop apart)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import contextlib
        import io
        import ug_pg_unet_b200.herlev as H

        with contextlib.redirect_stdout(io.StringIO()):
            tr = H.HerlevTrainer({"device": "cpu", "epochs_per_stage": 1, "num_classes": 7, "log_interval": 0,
                                  "stages": [1]})
        assert tr.distributed and tr.world_size == world
        # rank-dependent per-batch numbers, as rank-local BatchNorm buffers produce them in validation
        calls = {"n": 0}

        def fake_forward_loss(data, target, stage):
            calls["n"] += 1
            f = torch.tensor(1.0 + rank + 0.1 * calls["n"])
            logits = torch.zeros(target.shape[0], 7)
            logits[:, 0] = 1.0
            return f, f * 0.5, logits, None, None

        tr._forward_loss = fake_forward_loss
        loader = [(torch.zeros(4, 3, 32, 32), torch.tensor([0, 0, 1, 2]))] * (2 + rank)   # ragged: 2 vs 3 batches
        res = tr.validate_epoch(loader, 1)
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        assert all(g == gathered[0] for g in gathered), gathered          # same numbers everywhere
        # loss = mean over ALL batches of all ranks; accuracy over all samples
        per_rank = [[1.0 + r + 0.1 * (i + 1) for i in range(2 + r)] for r in range(world)]
        want = sum(sum(v) for v in per_rank) / sum(len(v) for v in per_rank)
        assert abs(res[0] - want) < 1e-6 and abs(res[2] - 50.0) < 1e-6
        # reducer follows requires_grad
        with contextlib.redirect_stdout(io.StringIO()):
            tr.setup_optimizer_scheduler(1)
        n0 = sum(len(b.params) for b in tr._reducers[1].buckets)
        tr.models[1]._freeze_encoder()
        assert tr._reducer_sig[1] != tr._grad_signature(1)
        tr._build_reducer(1)
        n1 = sum(len(b.params) for b in tr._reducers[1].buckets)
        assert n1 < n0 and tr._reducer_sig[1] == tr._grad_signature(1)
        tr.close()
        assert not tr._reducers and not tr._graphs
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_herlev_data_parallel_host_logic_world2(tmp_path):
    port = _free_port()
    mp.spawn(_herlev_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

