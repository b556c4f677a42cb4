"""World-size-2 checks of the data-parallel host logic on CPU (gloo): batch sharding, the flat
gradient bucket + single all-reduce, and rank-synchronous branch selection.  No kernels run here
(the compute path has no CPU implementation); parameters are plain CPU tensors."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import xggm_b200 as X
        from xggm_b200.ddp import BranchSchedule, FlatGrads, shard_range
        torch.manual_seed(9595)  # same init on every rank
        model = X.XGGMHeads(32, "GCN", 1, n_nodes=6)
        fg = FlatGrads(model.parameters())
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in fg.params)
        # rank-dependent "gradients": parameter i on rank r gets the constant (i + 1) * (r + 1)
        for i, p in enumerate(fg.params):
            p.grad.fill_(float((i + 1) * (rank + 1)))
        fg.all_reduce(average=True)
        mean_scale = sum(r + 1 for r in range(world)) / world
        ok = all(torch.allclose(p.grad, torch.full_like(p, (i + 1) * mean_scale)) for i, p in enumerate(fg.params))
        # bucketed variant: the early bucket is reduced from a backward hook at the layer boundary, the rest at
        # the end; the result must equal the single-collective average of the same per-rank gradients
        from xggm_b200 import ddp
        torch.manual_seed(1)
        l1, l2 = torch.nn.Linear(8, 8), torch.nn.Linear(8, 4)
        fg2 = FlatGrads(list(l1.parameters()) + list(l2.parameters()), early=list(l2.parameters()))
        ok = ok and fg2.split > 0 and fg2.params[0] is l2.weight
        xin = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank))
        fired = []
        with fg2.overlap(average=True):
            h = torch.tanh(l1(xin))
            ddp.notify_layer_boundary(h)
            h.register_hook(lambda g: fired.append(fg2._early_in_flight))
            l2(h).square().sum().backward()
        ok = ok and fired == [True]            # the early bucket went out before backward finished
        local = fg2.flat.clone()
        # (the early part of `local` is already averaged; rebuild the pure local gradient for the reference)
        l1.zero_grad(set_to_none=False); l2.zero_grad(set_to_none=False)
        fg2.flat.zero_()
        l2(torch.tanh(l1(xin))).square().sum().backward()
        ref = fg2.flat.clone()
        dist.all_reduce(ref)
        ref /= world
        fg2.flat.copy_(local)
        fg2._early_in_flight = True
        fg2.all_reduce(average=True)
        ok = ok and torch.allclose(fg2.flat, ref, rtol=1e-6, atol=1e-7)
        lo, hi = shard_range(37, rank, world)
        sched = BranchSchedule(delta=5, seed=9595)
        picks = [sched.next() for _ in range(32)]
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, picks, float(fg.flat.sum())))
        if rank == 0:
            out.put((ok, gathered))
        else:
            out.put((ok, None))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_sharding_and_branch_sync_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in results)
    gathered = next(g for _, g in results if g is not None)
    (lo0, hi0, picks0, s0), (lo1, hi1, picks1, s1) = gathered
    assert (lo0, hi0, lo1, hi1) == (0, 19, 19, 37)      # disjoint, covering, sizes differ by <= 1
    assert picks0 == picks1 and {"relation", "node"} == set(picks0)  # same branch on every rank
    assert s0 == s1                                       # identical reduced gradients


def test_shard_range_properties():
    from xggm_b200.ddp import shard_range
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
