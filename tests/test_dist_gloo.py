"""world_size-2 gloo test of the N>1 path's host logic: contiguous edge-balanced graph shards,
no data-path collective in the forward, sharded outputs concatenate to the unsharded result.
The per-shard forward here is the CPU oracle (the test checks the sharding plumbing; the CUDA
forward on shards is covered by tests/test_gpu_forward.py::test_sharded_equals_unsharded)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import graphnets_b200 as gn
    import workloads as W
    from oracle import gn_oracle as O
    w = W.make_workload("cfg3", B=12, n_nodes=(2, 7))
    layers = [("block", W.block_params(np.random.default_rng(1), (0, 100, 0), (6, 5, 4))),
              ("core", W.core_params(np.random.default_rng(2), (6, 5, 4)))]
    shard, (lo, hi) = gn.shard_batch(W.as_batch_input(w), rank, world)
    g = O.lower(shard["graphs"])
    nf = np.concatenate([x.T for x in shard["nf"]]) if hi > lo else np.zeros((0, 100), np.float32)
    ye, yn, yg = O.forward_sparse(layers, g, None, nf, None)
    # gather variable-size shards: sizes first, then padded tensors (no collective is needed by the
    # forward itself; this is the result collection a caller would do)
    outs = []
    for y in (ye, yn, yg):
        t = torch.from_numpy(np.ascontiguousarray(y))
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]]))
        mx = max(int(s) for s in sizes)
        pad = torch.zeros((mx, t.shape[1]), dtype=t.dtype)
        pad[:t.shape[0]] = t
        bufs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        outs.append(torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)]).numpy())
    ranges = [None] * world
    dist.all_gather_object(ranges, (lo, hi))
    if rank == 0:
        gfull = O.lower(w["graphs"])
        full = O.forward_sparse(layers, gfull, None, np.concatenate([x.T for x in w["nf"]]), None)
        # float64 BLAS blocks differently for different row counts: equal to rounding, same shapes
        ok = all(a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-12) for a, b in zip(outs, full))
        q.put((ok, ranges))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_forward_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, ranges = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 12


def _flatten(tree):
    out = []
    for kind, p in tree:
        def rec(o):
            if isinstance(o, dict):
                for k in sorted(o):
                    rec(o[k])
            elif isinstance(o, list):
                for v in o:
                    rec(v)
            elif isinstance(o, np.ndarray) and o.dtype.kind == "f":
                out.append(o.reshape(-1))
        rec(p)
    return np.concatenate(out)


def _grad_worker(rank, world, port, q):
    """Data-parallel training step over graph shards: every rank differentiates the loss of ITS shard, one
    all-reduce(sum) of the flat gradient buffer gives the gradient of the whole batch.  Per-shard gradients here come from the
    CPU oracle (torch float64 autograd); the CUDA backward is covered by tests/test_gpu_train.py."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import graphnets_b200 as gn
    import workloads as W
    from oracle import gn_oracle as O, gn_grad_oracle as G
    allreduce_flat = gn.pkg.train.allreduce_flat
    w = W.make_workload("cfg3", B=10, n_nodes=(2, 6))
    layers = [("block", W.block_params(np.random.default_rng(1), (0, 100, 0), (6, 5, 4))),
              ("core", W.core_params(np.random.default_rng(2), (6, 5, 4)))]

    def loss_grads(graphs, nfs):      # L = sum of all outputs (cotangent of ones): additive over graphs
        g = O.lower(graphs)
        nf = np.concatenate([x.T for x in nfs]) if len(nfs) else np.zeros((0, 100), np.float32)
        ones = lambda rows, d: np.ones((rows, d))
        _, pg, _ = G.forward_and_grads(layers, g, None, nf, None, [ones(g["E"], 6), ones(g["N"], 5), ones(g["B"], 4)])
        return _flatten(pg)
    shard, (lo, hi) = gn.shard_batch(W.as_batch_input(w), rank, world)
    flat = torch.from_numpy(loss_grads(shard["graphs"], shard["nf"]))
    allreduce_flat(flat, average=False)
    if rank == 0:
        full = loss_grads(w["graphs"], w["nf"])
        q.put(bool(np.allclose(flat.numpy(), full, rtol=1e-10, atol=1e-10)) and flat.numel() > 500)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
