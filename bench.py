#!/usr/bin/env python
"""bench.py - GNCore-model forward throughput (edges/s, graphs/s) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg4|cfg5]
                    [--precision auto|fp32|bf16] [--graphs B] [--no-extra] [--no-cpu-baseline] [--no-parity]

A "step" is one forward of the whole model (encoder GNBlock -> n x GNCore -> decoder GNBlock; BASELINE.json config 4 by
default: hidden 128, 4 cores, 4096 random graphs of 64 nodes / 512 edges per GPU) over one batch of synthetic inputs.
  value : inputs already resident in HBM, K steps between CUDA events on the launching stream, no per-launch instrumentation.
  e2e   : the same batch through the host-buffer C-ABI calls (gnb_graph_lower + gnb_model_forward_host) from pinned host
          memory - H2D of the bit-packed adjacency and the features, GPU lowering, forward, D2H of the outputs inside the
          timed region, K steps, double-buffered over two contexts (two host threads alternate batches).
  kernels / roofline : a separate, instrumented pass (CUDA events around every launch, gnb_ctx_set_profiling).
  extra : bounded legs for what north_star names beside the headline: the hidden-256 shard of config 5, config 3 at full size,
          config 2 (launch-bound) with and without a CUDA graph, the fp32-precision path.
N > 1: one process per GPU under torchrun, every rank owns its own shard of the graph batch (weak scaling), no data-path
collective; time = max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads as W  # noqa: E402

METRIC = "edges/sec, GNCore-model forward (graphs/sec alongside)"
H128_EXECUTED = 18.0 / 24.0      # the fused edge kernel executes 18 H^2 of the canonical 24 H^2 flop per edge (linearity split)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


def pin_host_threads(local_rank, local_world):
    """Give every rank its own slice of the host cores NVML reports as local to its GPU (NUMA affinity), so the 2 e2e host
    threads + the NCCL / sampler threads of 8 ranks do not migrate over each other.  Returns the core list (or None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if len(allowed) < 2 * local_world:
            return None
        per = len(allowed) // local_world
        # GPUs that share an affinity set split it by their index among the GPUs with the same set
        mine = allowed[(local_rank % local_world) * per:(local_rank % local_world + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def synth(name, B, seed):
    """Compact synthetic inputs of a vector-mode config (cfg4 / cfg5 shape), no per-graph Python."""
    cfg = W.CONFIGS[name]
    rng = np.random.default_rng(seed)
    n, m = 64, 512
    adj = W.random_cells_adj(rng, B, n, m)
    ef = rng.random((B * m, cfg["enc"][0]), dtype=np.float32)
    nf = rng.random((B * n, cfg["enc"][1]), dtype=np.float32)
    return adj, ef, nf


# ------------------------------------------------------------------------------------ reference
def cpu_reference_time(name, Bs, steps, warmup, seed=123):
    """Times the reference formulation (dense broadcasters, padded slots) on the host cores."""
    import torch
    from oracle import gn_oracle as O, ref_cpu as R
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    layers = R.to_torch_params(W.model_params(name))
    adj, ef, nf = synth(name, Bs, seed)
    adjs = [adj[b] for b in range(Bs)]
    g = O.lower(adjs)
    d = R.TorchDenseBatch(adjs)
    efp, nfp, gfp = R.pad_inputs(adjs, ef, nf, None, g)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            R.forward(layers, d, efp, nfp, gfp)
            t1 = time.perf_counter()
            if i >= warmup:
                times.append(t1 - t0)
    return float(np.mean(times)), g["E"], Bs, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    Bs = args.ref_graphs
    t, E, B, threads = cpu_reference_time(args.config, Bs, args.steps, args.warmup)
    val = E / t
    sample = "%d graphs of the %s workload per step (reference formulation holds (4H,PN^2,B) in memory; scaled linearly, graphs are independent)" % (Bs, args.config)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "edges/s", "graphs_per_sec": B / t,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_graphs": Bs},
        "cpu_baseline": {"value": val, "unit": "edges/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def workload_name(args):
    c = W.CONFIGS[args.config]
    return "%s: enc%s -> %dx GNCore%s -> dec%s, %d random graphs/GPU of 64 nodes / 512 edges" % (
        args.config, c["enc"], c["cores"], c["hidden"], c["dec"], args.graphs)


# ------------------------------------------------------------------------------------ ours
class Timer:
    """K forwards between two CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def time(self, fn, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        self.barrier()
        ms = ev0.elapsed_time(ev1) / steps
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms


def oracle_parity(gn, args, adj, ef, nf, model, layers, graphs):
    """The timed kernels against the float64 oracle on a slice of this very batch (>= 512 graphs: every persistent CTA of the
    fused kernel runs several passes), BEFORE anything is timed.  The oracle is the checker here, nothing else."""
    from oracle import gn_oracle as O
    import torch
    Bp = min(args.graphs, 512)
    m, n = 512, 64
    xs = gn.batch_compact(adj[:Bp], ef[:Bp * m], nf[:Bp * n])
    y = model(xs, precision=args.precision)
    xs.graphs.engine.sync()
    g = O.lower([adj[b] for b in range(Bp)])
    ref = O.forward_sparse(layers, g, ef[:Bp * m], nf[:Bp * n], None)
    tol = 1e-5 if args.precision == "fp32" else 1e-2
    errs = {}
    for name, f, r in zip(("ef", "nf", "gf"), (y.ef, y.nf, y.gf), ref):
        errs[name] = O.rel_err(f.compact.cpu().numpy(), r)
        assert errs[name] <= tol, "bench parity: %s rel err %.3e > %.0e against the float64 oracle" % (name, errs[name], tol)
    del xs, y
    torch.cuda.empty_cache()
    return {"graphs": Bp, "tolerance": tol, "max_norm_rel_err": errs, "against": "float64 oracle (oracle/gn_oracle.py), same inputs and weights"}


def extra_legs(gn, args, T, rank, world, local_rank, peaks):
    """Bounded legs beside the headline (north_star): every number is K_x forwards between CUDA events after 3 warm-ups."""
    import torch
    out = {}
    steps = max(3, min(args.steps, 5))

    def run(name, model, x, prec, steps_=steps):
        fn = lambda: model(x, precision=prec)
        ms = T.time(fn, steps_, 3)
        g = x.graphs
        return {"ms_per_forward": ms, "edges": g.E, "nodes": g.N, "graphs": g.B, "edges_per_sec": world * g.E / (ms * 1e-3),
                "graphs_per_sec": world * g.B / (ms * 1e-3), "precision": prec, "steps": steps_}

    # ---- config 5's per-GPU shard: hidden 256, 65 536 graphs over 8 GPUs = 8192 graphs per GPU (every rank runs its shard)
    try:
        adj, ef, nf = synth("cfg5", 8192, 5000 + rank)
        layers = W.model_params("cfg5")
        x = gn.batch_compact(adj, ef, nf, device=local_rank)
        leg = run("cfg5", W.to_gn_model(gn, layers), x, "auto", steps_=3)
        fl, by = W.canonical_work(layers, x.graphs.E, x.graphs.N, x.graphs.B)
        leg.update({"workload": "cfg5 shard: enc -> 4x GNCore(256) -> dec, 8192 graphs/GPU (65 536 graphs over 8 GPUs)",
                    "tensor_frac_of_burst": fl / (leg["ms_per_forward"] * 1e-3) / 1e12 / peaks["bf16"],
                    "hbm_frac": by / (leg["ms_per_forward"] * 1e-3) / 1e9 / peaks["hbm"]})
        out["cfg5_shard_8192"] = leg
        del x
        torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001
        out["cfg5_shard_8192"] = {"error": repr(e)[:300]}
    # ---- config 5's TRAINING step (every rank: the gradient all-reduce is a collective), bounded size
    try:
        torch.cuda.set_device(local_rank)
        # 1024 graphs per GPU = an eighth of config 5's per-GPU shard (8192): one micro-batch of the step (activations are kept)
        out["cfg5_train_step"] = train_leg(torch, gn, W, T.dist, world, 1024)
        torch.cuda.empty_cache()
        out["cfg5_train_step_bf16"] = train_leg(torch, gn, W, T.dist, world, 1024, precision="bf16")
        torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001
        out["cfg5_train_step"] = {"error": repr(e)[:300]}
    if rank != 0:
        return out
    # ---- the headline model on the fp32 CUDA-core path (1e-5 parity): what the default precision of the drop-in costs
    try:
        adj, ef, nf = synth("cfg4", 4096, 1000)
        x = gn.batch_compact(adj, ef, nf, device=local_rank)
        model = W.to_gn_model(gn, W.model_params("cfg4"))
        # single rank timing for the rank-0-only legs
        T1 = Timer(T.torch, T.dist, 1, T.dev)
        leg = {}
        for prec in ("fp32",):
            fn = lambda: model(x, precision=prec)
            ms = T1.time(fn, 3, 3)
            leg = {"ms_per_forward": ms, "edges_per_sec": x.graphs.E / (ms * 1e-3), "precision": prec, "steps": 3,
                   "workload": "cfg4 at 4096 graphs, fp32 CUDA-core path (the drop-in's default precision)"}
        out["cfg4_fp32"] = leg
        # ---- CUDA graph of the whole forward (bf16 path): launch overhead of the ~35 launches removed
        try:
            out["cfg4_cuda_graph"] = cuda_graph_leg(torch, T1, model, x, "auto", steps)
        except Exception as e:      # noqa: BLE001
            out["cfg4_cuda_graph"] = {"error": repr(e)[:300]}
        del x
        torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001
        out["cfg4_fp32"] = {"error": repr(e)[:300]}
    # ---- config 2: 1024 same-structure 16-node graphs, dims (10,5,3): launch-bound; with and without a CUDA graph
    try:
        w = W.make_workload("cfg2", B=1024)
        x = gn.batch(W.as_batch_input(w), device=local_rank)
        model = W.to_gn_model(gn, W.model_params("cfg2"))
        T1 = Timer(T.torch, T.dist, 1, T.dev)
        ms = T1.time(lambda: model(x, precision="fp32"), 20, 5)
        leg = {"us_per_forward": ms * 1e3, "edges": x.graphs.E, "graphs": x.graphs.B, "graphs_per_sec": x.graphs.B / (ms * 1e-3),
               "precision": "fp32", "steps": 20, "workload": "cfg2: enc -> 2x GNCore(10,5,3) -> dec, 1024 x 16-node same-structure graphs"}
        try:
            cg = cuda_graph_leg(torch, T1, model, x, "fp32", 20)
            leg["us_per_forward_cuda_graph"] = cg["ms_per_forward"] * 1e3
            leg["launches_per_forward"] = cg["launches_per_forward"]
        except Exception as e:      # noqa: BLE001
            leg["cuda_graph_error"] = repr(e)[:300]
        out["cfg2"] = leg
    except Exception as e:      # noqa: BLE001
        out["cfg2"] = {"error": repr(e)[:300]}
    # ---- config 3 at full size: sort model, hidden 384, 4096 fully connected graphs of 8-64 nodes (variable structure)
    try:
        w = W.make_workload("cfg3")
        x = gn.batch(W.as_batch_input(w), device=local_rank)
        layers = W.model_params("cfg3")
        T1 = Timer(T.torch, T.dist, 1, T.dev)
        model = W.to_gn_model(gn, layers)
        ms = T1.time(lambda: model(x, precision="auto"), 3, 3)
        fl, by = W.canonical_work(layers, x.graphs.E, x.graphs.N, x.graphs.B)
        # edge balance of a contiguous 8-way graph split by edge count (shard.py), for the multi-GPU case
        ep = x.graphs.index()["graph_edge_ptr"].astype(np.int64)
        rng8 = gn.shard_ranges(np.diff(ep), 8)
        per = [int(ep[b1] - ep[b0]) for b0, b1 in rng8]
        out["cfg3_full"] = {"ms_per_forward": ms, "edges": x.graphs.E, "nodes": x.graphs.N, "graphs": x.graphs.B,
                            "edges_per_sec": x.graphs.E / (ms * 1e-3), "graphs_per_sec": x.graphs.B / (ms * 1e-3), "precision": "auto",
                            "steps": 3, "tensor_frac_of_burst": fl / (ms * 1e-3) / 1e12 / peaks["bf16"],
                            "edge_balance_8way": {"max_over_mean": max(per) / (sum(per) / len(per)), "edges_per_shard": per},
                            "workload": "cfg3: enc(0,100,0) -> 2x GNCore(384) -> dec(2,2,0), 4096 fully connected graphs of 8-64 nodes"}
    except Exception as e:      # noqa: BLE001
        out["cfg3_full"] = {"error": repr(e)[:300]}
    return out


def train_leg(torch, gn, W, dist, world, graphs, steps=3, lr=5e-6, precision="fp32"):
    """Config 5's training step (BASELINE configs[4]: hidden 256, gradient all-reduce) on the fp32 path, at a bounded number of
    graphs per GPU: forward with kept activations + cross-entropy on node and edge outputs + backward + all-reduce of the flat
    gradient buffer (NCCL) + AdamW (graphnets.jl_b200/train.py; gradients checked against torch float64 autograd in
    tests/test_gpu_train.py)."""
    rank_ = dist.get_rank() if (dist is not None and world > 1) else 0
    adj, ef, nf = synth("cfg5", graphs, 77 + rank_)      # every rank its own shard (weak scaling); weights replicated
    x = gn.batch_compact(adj, ef, nf, device=torch.cuda.current_device())
    tr = gn.Trainer(W.model_params("cfg5"), engine=x.graphs.engine, precision=precision)
    dev = tr.eng.torch_device
    g = x.graphs
    rng = np.random.default_rng(5)
    te = torch.from_numpy(np.eye(3, dtype=np.float32)[rng.integers(0, 3, g.E)]).to(dev)
    tn = torch.from_numpy(np.eye(4, dtype=np.float32)[rng.integers(0, 4, g.N)]).to(dev)

    def step():
        y = tr.forward(x)
        le, de = tr.cross_entropy(y[0], te)
        ln, dn = tr.cross_entropy(y[1], tn)
        tr.backward(de, dn, None)
        tr.step(lr=lr)
        return le + ln
    l0 = float(step().cpu())      # warm-up (workspace growth, index upload)
    step()                        # second warm-up: the caching allocator of the tensors the step allocates settles
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        l = step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return {"ms_per_step": ms, "graphs_per_gpu": graphs, "edges_per_gpu": int(g.E), "edges_per_sec": world * g.E / (ms * 1e-3),
            "parameters": int(tr.params.numel()), "allreduce_bytes_per_step": int(tr.params.numel()) * 4 if world > 1 else 0,
            "world": world, "precision": precision, "loss_first": l0, "loss_last": float(l.cpu()), "steps": steps,
            "workload": "cfg5 training step: enc -> 4x GNCore(256) -> dec, forward + cross-entropy + backward + gradient all-reduce + AdamW; " +
                        ("fp32 CUDA-core path" if precision == "fp32" else "GEMMs with bf16 operands on the tensor cores (k_tc_lin), weight gradients / LayerNorm / optimiser fp32")}


def cuda_graph_leg(torch, T1, model, x, prec, steps):
    """Captures one forward into a CUDA graph (the library launches on torch's capture stream) and replays it."""
    eng = x.graphs.engine
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            model(x, precision=prec)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    l0 = eng.launches
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = model(x, precision=prec)
    launches = eng.launches - l0
    eng.bind_stream()
    ms = T1.time(graph.replay, steps, 3)
    del y
    return {"ms_per_forward": ms, "launches_per_forward": launches, "precision": prec}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    cores = pin_host_threads(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    import graphnets_b200 as gn
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()
    T = Timer(torch, dist, world, dev)
    B = args.graphs
    adj, ef, nf = synth(args.config, B, 1000 + rank)
    layers = W.model_params(args.config)
    model = W.to_gn_model(gn, layers)
    t0 = time.perf_counter()
    x = gn.batch_compact(adj, ef, nf, device=local_rank)
    torch.cuda.synchronize()
    batch_first_ms = (time.perf_counter() - t0) * 1e3
    g = x.graphs
    eng = g.engine
    E, N = g.E, g.N
    in_bytes = (ef.nbytes + nf.nbytes)

    # ---- parity of the timed path against the oracle, before timing (rank 0)
    parity = None
    if rank == 0 and not args.no_parity:
        parity = oracle_parity(gn, args, adj, ef, nf, model, layers, g)

    # ---- device-resident throughput (`value`): no per-launch instrumentation inside the timed region
    fwd = lambda: model(x, precision=args.precision)
    for _ in range(args.warmup):
        y = fwd()
    del y
    for _ in range(3):      # same call pattern as the timed loop (result dropped -> same output buffers -> same key): the library
        fwd()               # captures the forward into a CUDA graph on the second call with a key and replays it from then on
    T.barrier()
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_max = T.time(fwd, args.steps, 0)
    clocks = sampler.stop()
    launches = eng.launches - launches0
    y = fwd()
    out_bytes = sum(int(f.compact.numel()) * 4 for f in (y.ef, y.nf, y.gf) if f is not None)

    # ---- instrumented pass: CUDA events around every launch (library profile), for the kernel table and the roofline
    psteps = max(3, min(args.steps, 10))
    eng.set_profiling(True)
    eng.read_profile()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(psteps):
        y = fwd()
    ev1.record()
    torch.cuda.synchronize()
    prof = eng.read_profile()
    eng.set_profiling(False)
    ms_prof = ev0.elapsed_time(ev1) / psteps

    # ---- end to end through the host-buffer ABI (`e2e`), every step from pinned host memory:
    #      H2D bit-packed adjacency + GPU lowering + H2D features + forward + D2H outputs
    t0 = time.perf_counter()
    mask = np.ascontiguousarray((adj == 1).transpose(0, 2, 1)).astype(np.uint8)
    bits_np = gn.pack_adjacency_bits(mask)
    host_pack_ms = (time.perf_counter() - t0) * 1e3
    bits = torch.from_numpy(bits_np).pin_memory()
    h_ef, h_nf = torch.from_numpy(ef).pin_memory(), torch.from_numpy(nf).pin_memory()
    dout = model._out_dims()
    mk_out = lambda: (torch.empty((E, dout[0]), dtype=torch.float32).pin_memory(), torch.empty((N, dout[1]), dtype=torch.float32).pin_memory(),
                      torch.empty((B, dout[2]), dtype=torch.float32).pin_memory())
    nn = (C.c_int32 * B)(*([64] * B))
    mh = model._model(eng)
    L = gn.pkg._lib
    prec = L.PRECISIONS[args.precision]
    P = lambda t: C.c_void_p(t.data_ptr())

    def make_step(engine, bufs):
        h_oe_, h_on_, h_og_ = bufs

        def step():
            h = C.c_void_p()
            L.check(gn.lib.gnb_graph_lower(engine.ctx, P(bits), L.ADJ_BITS, 0, nn, 64, B, B, C.byref(h)))
            L.check(gn.lib.gnb_model_forward_host(engine.ctx, mh, h, P(h_ef), P(h_nf), None, P(h_oe_), P(h_on_), P(h_og_), prec))
            gn.lib.gnb_graph_destroy(h)
        return step

    bufs1 = mk_out()
    e2e_step = make_step(eng, bufs1)
    e2e_steps = args.steps + (args.steps % 2)
    for _ in range(3):      # first calls grow / coalesce the workspace arenas (cudaMalloc / cudaFree)
        e2e_step()
    T.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    T.barrier()
    e2e_single_s = (time.perf_counter() - t0) / e2e_steps
    for hb, f in zip(bufs1, (y.ef, y.nf, y.gf)):      # the host path must agree with the device path
        assert torch.equal(hb, f.compact.cpu()), "host-ABI result differs from the device-resident result"

    e2e_s, e2e_mode = e2e_single_s, "single context, strictly sequential steps"
    if args.e2e_streams >= 2:
        # Double-buffered input pipeline: two host threads, each with its own context + stream + pinned output buffers, alternate
        # batches through the same synchronous calls, so the PCIe copies and the lowering of one batch overlap the forward of the
        # other.  Every step still uploads its own inputs and downloads its own results.
        eng2 = gn.pkg.engine.Engine(local_rank)
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        engs = [eng, eng2]
        for e_, s_ in zip(engs, streams):
            L.check(gn.lib.gnb_ctx_set_stream(e_.ctx, C.c_void_p(s_.cuda_stream)))
        bufs2 = mk_out()
        steps2 = [make_step(eng, bufs1), make_step(eng2, bufs2)]
        errors = []

        def run_pipelined(n):
            def worker(k):
                try:
                    torch.cuda.set_device(local_rank)
                    for _ in range(n // 2):
                        steps2[k]()
                except Exception as e:      # noqa: BLE001
                    errors.append(e)
            ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
            for t_ in ts:
                t_.start()
            for t_ in ts:
                t_.join()
            if errors:
                raise errors[0]
        run_pipelined(4)      # warm-up of the second context (arena growth)
        T.barrier()
        t0 = time.perf_counter()
        run_pipelined(e2e_steps)
        T.barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e_mode = "2 host threads x (context + stream) alternate batches: copies / lowering of one batch overlap the forward of the other"
        for hb, f in (list(zip(bufs1, (y.ef, y.nf, y.gf))) + list(zip(bufs2, (y.ef, y.nf, y.gf)))):
            assert torch.equal(hb, f.compact.cpu()), "pipelined host-ABI result differs from the device-resident result"
        eng.bind_stream()

    # ---- the same through the repo's Python API (batch_compact + model + .cpu()): includes the host-side mask packing
    t0 = time.perf_counter()
    for _ in range(2):
        xx = gn.batch_compact(adj, ef, nf, device=local_rank)
        yy = model(xx, precision=args.precision)
        _ = [f.compact.cpu() for f in (yy.ef, yy.nf, yy.gf) if f is not None]
    py_api_ms = (time.perf_counter() - t0) / 2 * 1e3
    del xx, yy

    # ---- extra legs (all ranks take part in the sharded one)
    extra = None
    if not args.no_extra and args.config == "cfg4":
        extra = extra_legs(gn, args, T, rank, world, local_rank, peaks)

    # ---- reduce over ranks
    t = torch.tensor([e2e_s * 1e3, e2e_single_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(E), float(B), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    e2e_ms_max, e2e_single_ms_max = float(t[0]), float(t[1])
    E_all, B_all = float(tot[0]), float(tot[1])
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (instrumented pass)
    top = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roof = None
    if top[0] is not None:
        name, p = top
        per_ms = p["ms"] / p["launches"]
        alg_fl, alg_by = p["alg_flops"] / p["launches"], p["alg_bytes"] / p["launches"]
        if name.startswith("tc_"):
            # timed per launch with CUDA events at full clocks: the burst cuBLAS figure is the denominator (the sustained one
            # was measured power-throttled at ~1340 MHz); both fractions are given
            ach = alg_fl / (per_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": ach / peaks["bf16"],
                    "frac_of_sustained_peak": ach / peaks["bf16_sustained"]}
            if name == "tc_edge_core":
                ex = alg_fl * H128_EXECUTED
                roof.update({"executed_flops_per_launch": ex, "executed_tflops": ex / (per_ms * 1e-3) / 1e12,
                             "executed_frac": ex / (per_ms * 1e-3) / 1e12 / peaks["bf16"],
                             "executed_note": "18 H^2 of the canonical 24 H^2 flop per edge run in this kernel; the sender / receiver projections (6 H^2) run once per NODE in tc_node_proj"})
        else:
            ach = alg_by / (per_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"]}
        # DRAM bytes (read + write) of one launch of this kernel: the committed ncu --set full capture of the same workload
        traffic, tsrc = None, None
        for cand in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            tp = os.path.join(ROOT, "profiles", cand)
            if os.path.exists(tp) and args.config == "cfg4" and args.graphs == 4096:
                tj = json.load(open(tp))
                if name in tj.get("kernels", {}):
                    traffic, tsrc = tj["kernels"][name]["traffic"], "profiles/" + cand
                    break
        roof.update({"traffic": traffic, "traffic_source": tsrc, "alg_bytes_per_launch": alg_by, "alg_flops_per_launch": alg_fl,
                     "hbm_frac_of_kernel": alg_by / (per_ms * 1e-3) / 1e9 / peaks["hbm"], "kernel": name,
                     "launches_per_step": p["launches"] / psteps, "avg_launch_ms": per_ms, "share_of_step": p["ms"] / (ms_prof * psteps),
                     "peak_source": "of " + peaks["src"], "timed": "CUDA events around every launch, separate instrumented pass of %d steps" % psteps})
    fl, by = W.canonical_work(layers, E, N, B)
    model_roof = {"canonical_flops": fl, "canonical_bytes": by, "hbm_frac": by / (ms_max * 1e-3) / 1e9 / peaks["hbm"],
                  "tensor_frac_of_burst": fl / (ms_max * 1e-3) / 1e12 / peaks["bf16"],
                  "tensor_frac_of_sustained": fl / (ms_max * 1e-3) / 1e12 / peaks["bf16_sustained"], "peak_source": "of " + peaks["src"]}
    kern = {k: {"ms_per_step": v["ms"] / psteps, "launches_per_step": v["launches"] / psteps}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- CPU baseline: reference formulation on the host cores, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        tcpu, Ecpu, Bcpu, threads = cpu_reference_time(args.config, args.ref_graphs, 2, 1)
        cpu = {"value": Ecpu / tcpu, "unit": "edges/s", "cores": threads, "kind": "port", "graphs_per_sec": Bcpu / tcpu,
               "sample": "%d graphs of the same workload, dense-broadcaster formulation on torch CPU, mean of 2 passes" % Bcpu}

    act_mb = 4.0 * (E + N) * max(W.CONFIGS[args.config]["hidden"]) / 1e6
    out = {
        "metric": METRIC, "value": E_all / (ms_max * 1e-3), "unit": "edges/s",
        "graphs_per_sec": B_all / (ms_max * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if (args.precision != "fp32" and any(k.startswith("tc_") for k in prof)) else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "precision": args.precision, "graphs_per_gpu": B,
                   "edges_per_gpu": E, "nodes_per_gpu": N,
                   "launch": "forward replayed as a CUDA graph by the library (gnb_model_forward captures the second call with a key); GNB_CUDA_GRAPH=0: eager",
                   "l2": "no flush: every core layer streams %.0f MB of activations in and out (inputs %.0f MB), far beyond the 126 MB L2, so each step runs cold" % (act_mb, in_bytes / 1e6),
                   "parallelism": "graph-sharded x%d, no data-path collective" % world, "host_cores_of_rank0": cores},
        "e2e": {"value": E_all / (e2e_ms_max * 1e-3), "unit": "edges/s", "ms_per_step": e2e_ms_max, "steps": e2e_steps,
                "h2d_bytes_per_step": int(bits_np.nbytes + in_bytes), "d2h_bytes_per_step": int(out_bytes),
                "includes": "H2D bit-packed adjacency + GPU lowering + H2D features + forward + D2H outputs (gnb_graph_lower + gnb_model_forward_host), every step",
                "pipelining": e2e_mode, "ms_per_step_single_context": e2e_single_ms_max,
                "host_bit_packing_ms": host_pack_ms, "python_api_ms_per_step": py_api_ms,
                "python_api_note": "gn.batch_compact (numpy mask + bit packing + pageable H2D) -> model -> .cpu(), rank 0, 2 steps: the host-side numpy work is visible here and is NOT part of e2e.value",
                "first_batch_ms": batch_first_ms},
        "gpu_launches": int(float(tot[2])),
        "parity": parity,
        "clocks": clocks, "roofline": roof, "model_roofline": model_roof, "kernels": kern, "kernels_pass_ms_per_step": ms_prof,
        "cpu_baseline": cpu, "extra": extra,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--graphs", type=int, default=4096, help="graphs per GPU")
    ap.add_argument("--ref-graphs", type=int, default=32, help="graphs per step of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (cfg5 shard, cfg3, cfg2, fp32)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check before timing")
    ap.add_argument("--e2e-streams", type=int, default=2, help="2 (default): double-buffered over two contexts; 1: strictly sequential steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
