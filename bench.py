#!/usr/bin/env python
"""bench.py -- headline benchmark of the sampling / grouping / neighbour-aggregation path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], "C2"): the operator chain of one PointNet++ MSG part-seg
forward+backward pass over a ShapeNet-shaped batch of 16 clouds x 2048 points -- FPS 2048->512
and 512->128, five dense ball queries, the fused xyz+feature grouping of every scale, both
three_nn / three_interpolate layers, and the backward kernels of every grouping / interpolation
(shapes from conf/models/segmentation/pointnet2.yaml:95-130).  The MLPs between the operators
are stock cuDNN/cuBLAS in the reference and are not part of the path.  One step = one pass of
that chain over one batch; metric = clouds/s.  With N GPUs every rank runs its own batch
(clouds shard, no data-path collective; weak scaling).

The JSON line also carries, as extra keys: the roofline of the dominant kernel (measured on the launch that is in
the step), the CPU baseline, the KPConv layer numbers with their own live roofline and CPU baseline ("kpconv", C3
layer 32->32, ~50k points, <=34 neighbours), the KPConvPaper segmentation training step ("kpconv_stack", C3/C5; under
DDP when N > 1), C4 (FPS 4096 of 120k + knn k=16, and the RandLA chain on top) sharded over the ranks, the C1 chain,
the whole C2 network step against the published iteration time, and per-operator figures ("operators").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_CLOUDS, N_PTS = 16, 2048
METRIC = "PointNet++ MSG SA/FP operator chain fwd+bwd (FPS + ball_query + grouping + three_nn/interpolate)"
WORKLOAD = ("C2: PointNet++ MSG part-seg op chain fwd+bwd, batch 16 x 2048 pts per GPU "
            "(SA1 512 pts r[.1,.2,.4] k[32,64,128] C=3; SA2 128 pts r[.4,.8] k[64,128] C=320; "
            "FP 128->512 C=256, 512->2048 C=128)")
UNIT = "clouds/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------- inputs -----------
def make_host_inputs(seed):
    from torch_points3d_b200 import synthetic
    rng = np.random.default_rng(seed)
    pos = synthetic.surface_clouds(B_CLOUDS, N_PTS, seed)
    nrm = pos / (np.linalg.norm(pos, axis=-1, keepdims=True) + 1e-12)
    h = dict(pos=pos, x0=np.ascontiguousarray(nrm.transpose(0, 2, 1)).astype(np.float32))
    h["x1"] = rng.standard_normal((B_CLOUDS, 320, 512)).astype(np.float32)    # SA1 output features
    h["x2"] = rng.standard_normal((B_CLOUDS, 256, 128)).astype(np.float32)    # FP2 input (SA2 level)
    h["x3"] = rng.standard_normal((B_CLOUDS, 128, 512)).astype(np.float32)    # FP3 input (SA1 level)
    return h


# ------------------------------------------------------------------- the b200 arm -------------
class Chain:
    """the C2 operator chain on one GPU through the public API (torch_points3d_b200.ops)"""

    REAL_INPUTS = ("pos", "x0")  # what a training pipeline gets from the host: positions and the input features

    def __init__(self, dev, seed):
        import torch
        from torch_points3d_b200 import synthetic
        self.torch, self.dev = torch, dev
        self.sa = synthetic.C2_SA
        self.host_all = {k: torch.from_numpy(v).pin_memory() for k, v in make_host_inputs(seed).items()}
        self.host = self.host_all
        self.d = {k: torch.empty_like(v, device=dev) for k, v in self.host_all.items()}
        for k in ("x0", "x1", "x2", "x3"):
            self.d[k].requires_grad_(True)
        g = torch.Generator(device=dev).manual_seed(seed)
        # upstream gradients of every grouped / interpolated tensor (what the MLP backward would deliver)
        self.g_sa = [[torch.randn((B_CLOUDS, 3 + l["cin"], l["npoint"], k), device=dev, generator=g)
                      for k in l["nsample"]] for l in self.sa]
        self.g_fp = [torch.randn((B_CLOUDS, 256, 512), device=dev, generator=g),
                     torch.randn((B_CLOUDS, 128, 2048), device=dev, generator=g)]
        self.geo_stream = torch.cuda.Stream(device=dev)
        self.geo_stream2 = torch.cuda.Stream(device=dev)
        # result of a step read back by the host: sampled indices of both levels + the gradient w.r.t. the input features
        self.result_shapes = [((B_CLOUDS * (512 + 128),), torch.int64), ((B_CLOUDS, 3, N_PTS), torch.float32)]
        self.d2h_bytes = B_CLOUDS * (512 + 128) * 8 + B_CLOUDS * 3 * N_PTS * 4

    def upload(self):
        for k, v in self.host_all.items():
            self.d[k].data.copy_(v, non_blocking=True)

    def make_staging(self, keys=None):
        """two device staging sets: the H2D copy of step i+1 runs on a copy stream while step i computes.
        `keys` = which inputs come from the host every step (default: all of them)"""
        torch = self.torch
        self.host = {k: v for k, v in self.host_all.items() if keys is None or k in keys}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.stage = [{k: torch.empty_like(v, device=self.dev) for k, v in self.host.items()} for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.results = [[torch.empty(shape, dtype=dt).pin_memory() for shape, dt in self.result_shapes] for _ in range(2)]
        for slot in range(2):
            self.consumed[slot].record(torch.cuda.current_stream())

    def h2d_async(self, slot):
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])  # the compute stream has read this staging set
            for k, v in self.host.items():
                self.stage[slot][k].copy_(v, non_blocking=True)
            self.staged[slot].record(self.copy_stream)

    def stage_to_inputs(self, slot):
        torch = self.torch
        cur = torch.cuda.current_stream()
        cur.wait_event(self.staged[slot])
        for k in self.host:
            self.d[k].data.copy_(self.stage[slot][k], non_blocking=True)  # D2D into the graph's input buffers
        self.consumed[slot].record(cur)

    # The chain has two stages.  GEOMETRY reads positions only: FPS of both levels, the five ball queries, both
    # three_nn searches and the inverse index of every neighbour table (needed by the backward kernels).  FEATURES
    # reads the tables: fused xyz+feature grouping of every scale, both interpolations, then all backward kernels in
    # reverse layer order.  Geometry is latency-bound on a few SMs (FPS: one CTA per cloud, 511 dependent rounds),
    # features are bandwidth-bound on all of them, and geometry does not depend on anything the network computes --
    # so `step` software-pipelines them: one step = FEATURES of batch i on the main stream while GEOMETRY of batch
    # i+1 runs on two side streams (what a training loop does when positions arrive one batch ahead; the reference
    # precomputes the same tables in DataLoader workers, core/data_transform/transforms.py:603-651).  Every batch
    # still gets the whole chain, one batch completes per step, its latency is two steps.
    def geometry(self, pos):
        """positions [B,N,3] -> every table the feature stage reads (runs on self.geo_stream + self.geo_stream2).
        Measured (bench_tools/stage_times.py): forking every independent search -> inverse-index chain onto its own
        stream does not shorten the stage -- apart from FPS (16 CTAs) its kernels each fill the GPU."""
        torch = self.torch
        from torch_points3d_b200 import ops
        sa1, sa2 = self.sa
        g0, g1 = self.geo_stream, self.geo_stream2
        t = {"pos": pos.clone()}
        t["idx1"], t["pos1"] = ops.furthest_point_sample(pos, sa1["npoint"], return_xyz=True)
        g1.wait_stream(g0)
        with torch.cuda.stream(g1):
            t["idx2"], t["pos2"] = ops.furthest_point_sample(t["pos1"], sa2["npoint"], return_xyz=True)
            tables = ops.ball_query_multi(sa2["radii"], sa2["nsample"], t["pos1"], t["pos2"])  # all scales, one sweep
            for si in range(len(tables)):
                t["ridx2_%d" % si] = tables[si]
                t["prep2_%d" % si] = ops.prepare_grouping(t["ridx2_%d" % si], sa2["cin"], sa1["npoint"]).ws
            _, t["nn0_idx"], t["nn0_w"] = ops.three_nn(t["pos1"], t["pos2"], return_weight=True)
            t["prepi0"] = ops.prepare_three_interpolate(t["nn0_idx"], 256, sa2["npoint"]).ws
        tables = ops.ball_query_multi(sa1["radii"], sa1["nsample"], pos, t["pos1"])
        for si in range(len(tables)):
            t["ridx1_%d" % si] = tables[si]
            t["prep1_%d" % si] = ops.prepare_grouping(t["ridx1_%d" % si], sa1["cin"], N_PTS).ws
        _, t["nn1_idx"], t["nn1_w"] = ops.three_nn(pos, t["pos1"], return_weight=True)
        t["prepi1"] = ops.prepare_three_interpolate(t["nn1_idx"], 128, sa1["npoint"]).ws
        g0.wait_stream(g1)
        return t

    def features(self, t):
        """forward of every grouping / interpolation over the tables `t`, then the backward kernels in reverse
        layer order (the order a training step runs them in); returns the last gradient"""
        torch = self.torch
        from torch_points3d_b200 import ops
        d = self.d
        sa1, sa2 = self.sa
        P = lambda ws: ops.PreparedInverse(ws, ws.numel())  # noqa: E731
        tape = []
        for si in range(len(sa1["nsample"])):
            tape.append((ops.group_features(d["x0"], t["pos"], t["pos1"], t["ridx1_%d" % si],
                                            prepared=P(t["prep1_%d" % si])), d["x0"], self.g_sa[0][si]))
        for si in range(len(sa2["nsample"])):
            tape.append((ops.group_features(d["x1"], t["pos1"], t["pos2"], t["ridx2_%d" % si],
                                            prepared=P(t["prep2_%d" % si])), d["x1"], self.g_sa[1][si]))
        # FP2: 512 unknown (SA1 level) from 128 known (SA2 level); FP3: 2048 unknown from 512 known; the weights of
        # core/base_conv/dense.py:137-139 come out of the three_nn kernel
        tape.append((ops.three_interpolate(d["x2"], t["nn0_idx"], t["nn0_w"], prepared=P(t["prepi0"])), d["x2"], self.g_fp[0]))
        tape.append((ops.three_interpolate(d["x3"], t["nn1_idx"], t["nn1_w"], prepared=P(t["prepi1"])), d["x3"], self.g_fp[1]))
        gx = None
        for out, leaf, g in reversed(tape):
            (gx,) = torch.autograd.grad(out, leaf, g)
        return gx

    def prime(self):
        """geometry of the first batch (eager, outside the timed region): fills the table set the first step reads"""
        torch = self.torch
        cur = torch.cuda.current_stream()
        self.geo_stream.wait_stream(cur)
        with torch.cuda.stream(self.geo_stream):
            t = self.geometry(self.d["pos"])
        cur.wait_stream(self.geo_stream)
        self.tables = {k: v.clone() for k, v in t.items()}
        self.d["pos1"], self.d["pos2"] = self.tables["pos1"], self.tables["pos2"]

    def step(self):
        """FEATURES of the current batch (tables in self.tables) overlapped with GEOMETRY of the next batch
        (positions in d["pos"]); the new tables replace the old ones at the end.  Returns the step's result:
        the sampled indices of both levels (next batch) and the gradient w.r.t. the input features (current batch)."""
        torch = self.torch
        cur = torch.cuda.current_stream()
        self.geo_stream.wait_stream(cur)
        with torch.cuda.stream(self.geo_stream):
            new = self.geometry(self.d["pos"])
        gx = self.features(self.tables)
        cur.wait_stream(self.geo_stream)
        keys = list(self.tables)
        torch._foreach_copy_([self.tables[k] for k in keys], [new[k] for k in keys])
        return torch.cat([new["idx1"].reshape(-1), new["idx2"].reshape(-1)]), gx

    def step_serial(self):
        """the same chain without pipelining (geometry, then features, of ONE batch): used once to check that the
        pipelined step computes the same thing"""
        torch = self.torch
        cur = torch.cuda.current_stream()
        self.geo_stream.wait_stream(cur)
        with torch.cuda.stream(self.geo_stream):
            t = self.geometry(self.d["pos"])
        cur.wait_stream(self.geo_stream)
        gx = self.features(t)
        return torch.cat([t["idx1"].reshape(-1), t["idx2"].reshape(-1)]), gx

    @staticmethod
    def algorithmic_bytes_dominant(fused=True):
        """grouping of SA2 scale 2 (4*C*N + 8*S*K + 4*C*S*K per cloud).  The kernel in the step is the
        fused xyz+feature gather (p3d_group_xyz_feat_f32): it also reads the positions of both levels (12*(N+S)) and
        writes three more channels; `fused=False` is the plain 320-channel gather (p3d_group_fwd_f32)."""
        C, Np, S, K = 320, 512, 128, 128
        if fused:
            return B_CLOUDS * (4 * C * Np + 12 * (Np + S) + 8 * S * K + 4 * (C + 3) * S * K)
        return B_CLOUDS * (4 * C * Np + 8 * S * K + 4 * C * S * K)


class ClockSampler(threading.Thread):
    """`nvidia-smi -lms 100` running beside the timed regions (the recipe's clocks line)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if line.strip():
                    self.samples.append([c.strip() for c in line.strip().split(",")])
        except Exception:  # noqa: BLE001
            pass

    def summary(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def _watchdog(seconds, line_holder):
    """the secondary sections run after the headline is measured; if one of them wedges (e.g. a collective that one
    rank never reaches), print what is measured and leave instead of hanging the whole run"""
    def fire():
        if line_holder.get("line") is not None and int(os.environ.get("RANK", "0")) == 0:
            line_holder["line"]["aborted_extras"] = "watchdog after %d s" % seconds
            print(json.dumps(line_holder["line"]), flush=True)
        os._exit(0 if line_holder.get("line") is not None or int(os.environ.get("RANK", "0")) != 0 else 3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200 arm has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")  # required to capture NCCL work in CUDA graphs
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from torch_points3d_b200 import _cabi, ops

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    chain = Chain(dev, seed=1002 + 97 * rank)
    chain.upload()
    torch.cuda.synchronize()

    # ---- eager warm-up (also loads modules, sets kernel attributes), launch count per step ----
    chain.prime()
    serial_out, serial_gx = chain.step_serial()
    for _ in range(2):
        piped_out, piped_gx = chain.step()
    torch.cuda.synchronize()
    if not (torch.equal(serial_out, piped_out) and torch.equal(serial_gx, piped_gx)):
        raise SystemExit("bench.py: pipelined step disagrees with the serial chain")  # same inputs every step
    c0 = _cabi.launch_count()
    chain.step()
    torch.cuda.synchronize()
    launches_per_step = _cabi.launch_count() - c0

    # ---- capture one step in a CUDA graph (the chain is launch-bound from Python otherwise) ----
    graph, graph_out, use_graph = None, None, not args.no_graph
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                chain.step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_out = chain.step()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print("bench.py: CUDA graph capture failed (%s); timing eagerly" % e, file=sys.stderr)
            graph, use_graph = None, False
            torch.cuda.synchronize()

    def one_step():
        if graph is not None:
            graph.replay()
            return graph_out
        return chain.step()

    # ---- device-resident timed region ----
    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)  # let the sampler come up before the timed regions
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- end to end: every step copies ITS inputs from pinned host memory (copy stream, one step ahead of
    # the compute stream), runs the chain and reads its result back to pinned host memory.  Measured twice:
    # (1) the inputs a training pipeline really gets from the host (positions + input features; the SA/FP
    #     activations x1..x3 are produced on the device by the network's MLPs) -- the headline `e2e`;
    # (2) every tensor the chain reads, synthetic activations included, from the host (`e2e_all_from_host`,
    #     the round-1 configuration: an upper bound on input traffic no real pipeline has).
    def e2e_measure(keys):
        chain.make_staging(keys)

        def e2e_run(nsteps):
            chain.h2d_async(0)
            for i in range(nsteps):
                slot = i & 1
                if i + 1 < nsteps:
                    chain.h2d_async(slot ^ 1)
                chain.stage_to_inputs(slot)
                out = one_step()
                for dst, src in zip(chain.results[slot], out):
                    dst.copy_(src.reshape(dst.shape), non_blocking=True)

        e2e_run(max(3, args.warmup // 2))
        barrier()
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), chain.h2d_bytes

    e2e_ms_total, e2e_h2d = e2e_measure(Chain.REAL_INPUTS)
    e2e_all_ms_total, e2e_all_h2d = e2e_measure(None)
    chain.upload()
    clocks = sampler.summary() if sampler else None

    clouds = B_CLOUDS * world * args.steps
    line = {
        "metric": METRIC, "value": round(clouds / (ms_total * 1e-3), 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "clouds_per_gpu": B_CLOUDS, "points_per_cloud": N_PTS, "parallelism": "clouds sharded x%d" % world,
                   "cuda_graph": bool(use_graph),
                   "streams": "software-pipelined: each step runs the feature stage (grouping, interpolation, all "
                              "backward kernels) of batch i on the main stream while the positions-only stage (FPS, "
                              "ball queries, three_nn, inverse indices) of batch i+1 runs on two side streams inside "
                              "the same captured graph; every batch gets the whole chain, one batch completes per step",
                   "l2": "each step streams ~1.1 GB of grouped tensors (> 126 MB L2), no flush needed"},
        "e2e": {"value": round(clouds / (e2e_ms_total * 1e-3), 2), "unit": UNIT,
                "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": chain.d2h_bytes,
                "ms_per_step": round(e2e_ms_total / args.steps, 4),
                "inputs_from_host": "pos [16,2048,3] + input features x0 [16,3,2048] (what a data loader delivers); "
                                    "result read back: sampled indices of both levels + d(loss)/d(x0)"},
        "e2e_all_from_host": {"value": round(clouds / (e2e_all_ms_total * 1e-3), 2), "unit": UNIT,
                              "h2d_bytes_per_step": e2e_all_h2d, "d2h_bytes_per_step": chain.d2h_bytes,
                              "ms_per_step": round(e2e_all_ms_total / args.steps, 4),
                              "inputs_from_host": "pos, x0 and the synthetic SA/FP activations x1, x2, x3 (16.8 MB that a "
                                                  "real network produces on the device)"},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "clocks": clocks,
    }
    holder = {"line": line if rank == 0 else None}
    dog = _watchdog(args.extras_timeout, holder)

    def section(name, fn, all_ranks=False):
        """secondary sections never cost the headline line: errors are recorded, not raised"""
        if args.skip_extras or (rank != 0 and not all_ranks):
            return
        try:
            res = fn()
        except Exception as e:  # noqa: BLE001
            res = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        if rank == 0:
            line[name] = res

    # sections every rank takes part in (collectives / sharded work) come first, in the same order on every rank
    section("kpconv_stack", lambda: bench_kpconv_stack(dev, world, rank), all_ranks=True)
    section("c4", lambda: bench_c4(dev, world, rank), all_ranks=True)
    if world > 1:
        dist.barrier()
    if rank != 0:
        dog.cancel()
        if world > 1:
            dist.destroy_process_group()
        return

    line["roofline"] = bench_roofline(dev, chain, args)
    section("kpconv", lambda: bench_kpconv(dev, args))
    section("operators", lambda: bench_operators(dev))
    section("c1", lambda: bench_c1(dev))
    section("c2_model", lambda: bench_c2_model(dev, line["ms_per_step"]))
    section("rsconv_model", lambda: bench_rsconv_model(dev))   # written after the last GPU session: first run at round end
    line["cpu_baseline"] = None
    section("cpu_baseline", lambda: cpu_baseline(max_seconds=12.0))
    dog.cancel()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_roofline(dev, chain, args):
    """roofline of the dominant kernel, measured live on the launch that is IN the step: the fused xyz+feature gather
    of SA2 scale 2 (p3d_group_xyz_feat_f32, [16,320,512] + positions -> [16,323,128,128], 339 MB of stores); the plain
    320-channel gather (p3d_group_fwd_f32) is timed beside it.  Back-to-back launches alternate between two outputs
    (2 x 339 MB > 126 MB L2)."""
    import torch
    from torch_points3d_b200 import _cabi, ops
    peaks, peaks_src = _peaks()
    x1 = chain.d["x1"].detach()
    pos1, pos2 = chain.d["pos1"], chain.d["pos2"]
    ridx = ops.ball_query(0.8, 128, pos1, pos2, return_dist=False)[0]
    outs = [torch.empty((B_CLOUDS, 323, 128, 128), device=dev) for _ in range(2)]
    st = torch.cuda.current_stream

    def fused(i):
        _cabi.call("p3d_group_xyz_feat_f32", pos1.data_ptr(), pos2.data_ptr(), x1.data_ptr(), ridx.data_ptr(), B_CLOUDS,
                   320, 512, 128, 128, 1.0, outs[i & 1].data_ptr(), st().cuda_stream)

    def plain(i):
        _cabi.call("p3d_group_fwd_f32", x1.data_ptr(), ridx.data_ptr(), B_CLOUDS, 320, 512, 128, 128,
                   outs[i & 1].data_ptr(), 323 * 128 * 128, 3, st().cuda_stream)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def avg_ms(fn):
        for i in range(5):
            fn(i)
        reps = max(args.steps, 20)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_f, ms_p = avg_ms(fused), avg_ms(plain)
    by_f, by_p = Chain.algorithmic_bytes_dominant(True), Chain.algorithmic_bytes_dominant(False)
    ach_f, ach_p = by_f / (ms_f * 1e-3) / 1e9, by_p / (ms_p * 1e-3) / 1e9
    return {"bound": "hbm",
            "kernel": "group_fwd_kernel as launched in the step: p3d_group_xyz_feat_f32, SA2 scale 2, "
                      "[16,320,512] + pos -> [16,323,128,128]",
            "achieved": round(ach_f, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": round(ach_f / peaks["hbm_gbs"], 4), "traffic": _ncu_traffic(),
            "algorithmic_bytes": by_f, "avg_launch_ms": round(ms_f, 5), "peak_source": peaks_src,
            "frac_of_nominal_8TBs": round(ach_f / 8000.0, 4),
            "plain_gather": {"kernel": "p3d_group_fwd_f32 (320 channels, same table)", "achieved": round(ach_p, 1),
                             "frac": round(ach_p / peaks["hbm_gbs"], 4), "algorithmic_bytes": by_p,
                             "avg_launch_ms": round(ms_p, 5)}}


def _ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None"""
    path = os.path.join(ROOT, "profiles", "dominant_traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            return None
    return None


def bench_kpconv(dev, args):
    """KPConv layer points/s on the C3 layer-0b shape: ~50k points, M = 34, Cin = Cout = 32."""
    import torch
    from torch_points3d_b200 import kpconv, ops, synthetic
    pos_h, batch_h = synthetic.indoor_spheres()
    pos, batch = torch.from_numpy(pos_h).to(dev), torch.from_numpy(batch_h).to(dev)
    n = pos.shape[0]
    nb = ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch, return_dist=False)[0]
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn((n, 32), device=dev, generator=g, requires_grad=True)
    W = (torch.randn((15, 32, 32), device=dev, generator=g) * 0.1).requires_grad_(True)
    kp = torch.from_numpy(kpconv.load_kernels(0.06, random_rotation=False)).to(dev)
    dout = torch.randn((n, 32), device=dev, generator=g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 10)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def fwd():
        with torch.no_grad():
            return ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum", precision="tensor")

    def fwd_fp32():
        with torch.no_grad():
            return ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum", precision="fp32")

    def fwd_bwd():
        out = ops.KPConv_ops(pos, pos, nb, x, kp, W, 0.04, "linear", "sum")
        torch.autograd.grad(out, (x, W), dout)

    def query():
        ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch, return_dist=False)

    def timed_graph(fn):
        """device time of fn: captured once in a CUDA graph and replayed (an eager forward+backward of one layer is
        bound by ~10 Python-side launches, not by the GPU)"""
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                keep = fn()  # noqa: F841
            return timed(graph.replay)
        except Exception:  # noqa: BLE001
            return None

    t_f, t_f32, t_fb, t_q = timed(fwd), timed(fwd_fp32), timed(fwd_bwd), timed(query)
    t_fb_graph = timed_graph(fwd_bwd)
    t_fb_eager = t_fb
    if t_fb_graph is not None and t_fb_graph < t_fb:
        t_fb = t_fb_graph
    with torch.no_grad():
        err = float((fwd() - fwd_fp32()).abs().max() / fwd_fp32().abs().max())
    # flops = n (2 K M Cin + 2 K Cin Cout + 12 K M); bytes = 12 (n + N0) + 8 n M + 4 N0 Cin + 4 K Cin Cout + 4 n Cout
    K_, M_, Ci, Co = 15, 34, 32, 32
    flops = n * (2 * K_ * M_ * Ci + 2 * K_ * Ci * Co + 12 * K_ * M_)
    nbytes = 12 * (n + n) + 8 * n * M_ + 4 * n * Ci + 4 * K_ * Ci * Co + 4 * n * Co
    peaks, peaks_src = _peaks()
    tf, gbs = flops / (t_f * 1e-3) / 1e12, nbytes / (t_f * 1e-3) / 1e9
    res = {"workload": "C3 layer 0b: KPConv 32->32, %d points in 8 spheres, M=34, K=15, linear/sum" % n,
           "fwd_points_per_s": round(n / (t_f * 1e-3), 1), "fwd_ms": round(t_f, 4),
           "fwd_path": "tcgen05 (bf16 hi+lo operands, fp32 accumulate in TMEM)",
           "fwd_fp32_simt_points_per_s": round(n / (t_f32 * 1e-3), 1), "fwd_fp32_simt_ms": round(t_f32, 4),
           "tensor_vs_fp32_max_rel_diff": float("%.3e" % err),
           "fwd_bwd_points_per_s": round(n / (t_fb * 1e-3), 1), "fwd_bwd_ms": round(t_fb, 4),
           "fwd_bwd_eager_ms": round(t_fb_eager, 4),
           "fwd_bwd_timing": "CUDA graph replay" if t_fb is not t_fb_eager and t_fb != t_fb_eager else "eager launches",
           "neighbour_search_ms": round(t_q, 4), "unit": "points/s",
           # live roofline of the forward kernel (this build, this run): both ceilings, since the operator is a gather
           # (HBM / L2) feeding a small dense contraction (tensor pipe); the SIMT neighbour aggregation bounds it
           "roofline": {"flops": flops, "algorithmic_bytes": nbytes, "avg_launch_ms": round(t_f, 5),
                        "tensor": {"achieved": round(tf, 3), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                   "frac": round(tf / peaks["bf16_tflops_sustained"], 5)},
                        "hbm": {"achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": round(gbs / peaks["hbm_gbs"], 4)},
                        "peak_source": peaks_src,
                        "ncu_capture": "profiles/r2_kpconv_fwd.json (tensor-pipe / issue counters of the same kernel)"}}
    # CPU baseline beside it (north_star: "the torch KPConv ops" on the host cores): the oracle's restatement of
    # convolution_ops.py:19-107 (pinned to the reference's own outputs by tests/golden/kpconv_*.npz), on a bounded sample
    try:
        import oracle
        oracle.set_num_threads(os.cpu_count() or 1)
        sub = min(n, 6000)
        nb_h = nb[:sub].cpu().numpy()
        pos_n, x_n, W_n, kp_n = pos_h, x.detach().cpu().numpy(), W.detach().cpu().numpy(), kp.cpu().numpy()
        oracle.kpconv_fwd(pos_n[:64], pos_n, nb_h[:64], x_n, kp_n, W_n, 0.04, "linear", "sum")
        t0, done = time.perf_counter(), 0
        while time.perf_counter() - t0 < 4.0:
            oracle.kpconv_fwd(pos_n[:sub], pos_n, nb_h, x_n, kp_n, W_n, 0.04, "linear", "sum")
            done += sub
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": round(done / dt, 1), "unit": "points/s", "cores": oracle.num_threads(),
                               "kind": "port", "sample": "%d query points of the same layer through oracle.kpconv_fwd "
                               "(forward only, %.1f s)" % (done, dt)}
    except Exception as e:  # noqa: BLE001
        res["cpu_baseline"] = {"error": str(e)[:160]}
    return res


def bench_operators(dev):
    """Per-operator device times on the C2 / C3 / C4 shapes with the figures SURVEY.md section 8d asks for: dependent
    rounds/s for FPS (latency bound), pair tests/s for the neighbour searches (FP32 ALU bound; the FP32 peak of
    148 SMs x 128 lanes x 1.965 GHz = 37.2 T tests/s would need one instruction per test, a test is ~8), achieved
    GB/s on algorithmic bytes for the gather / scatter kernels (HBM bound), points/s for voxel clustering and the
    deformable KPConv.  Every call is captured in a CUDA graph and replayed (device time, no launch overhead)."""
    import torch
    from torch_points3d_b200 import _cabi, kpconv, ops, synthetic
    peaks, _ = _peaks()
    res = {}

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        run = fn
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                keep = fn()  # noqa: F841
            run = graph.replay
            run()
        except Exception:  # noqa: BLE001
            run = fn
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3  # us

    def put(name, us, **kw):
        d = {"us": round(us, 2)}
        for k, v in kw.items():
            d[k] = float("%.4g" % v)
        if "gb_s" in d:
            d["hbm_frac"] = round(d["gb_s"] / peaks["hbm_gbs"], 3)
        res[name] = d

    B = B_CLOUDS
    g = torch.Generator(device=dev).manual_seed(0)
    xyz = torch.from_numpy(synthetic.surface_clouds(B, N_PTS, 1002)).to(dev)
    us = timed(lambda: ops.furthest_point_sample(xyz, 512))
    put("fps 16x2048->512", us, rounds_per_s=511 / (us * 1e-6), us_per_round=us / 511)
    idx1 = ops.furthest_point_sample(xyz, 512)
    pos1 = xyz.gather(1, idx1.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx2 = ops.furthest_point_sample(pos1, 128)
    pos2 = pos1.gather(1, idx2.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    x1 = torch.randn((B, 320, 512), device=dev, generator=g)
    for (p, q, r, k) in [(xyz, pos1, 0.2, 64), (pos1, pos2, 0.8, 128)]:
        n, s_ = p.shape[1], q.shape[1]
        us = timed(lambda: ops.ball_query(r, k, p, q, return_dist=False))
        put("ball_query N=%d S=%d r=%.1f K=%d" % (n, s_, r, k), us, pair_tests_per_s=B * n * s_ / (us * 1e-6),
            gb_s=B * (12 * (n + s_) + 8 * s_ * k) / us / 1e3)
    ridx = ops.ball_query(0.8, 128, pos1, pos2, return_dist=False)[0]
    C, S, K, n = 320, 128, 128, 512
    us = timed(lambda: ops.group_features(x1, pos1, pos2, ridx))
    put("group_features fwd SA2 K=128 C=320", us, gb_s=B * (4 * C * n + 8 * S * K + 4 * (C + 3) * S * K) / us / 1e3)
    gout = torch.randn((B, C + 3, S, K), device=dev, generator=g)
    ws_b = _cabi.lib().p3d_group_bwd_ws_bytes(B, C, n, S, K)
    ws = torch.empty(max(ws_b, 1), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream
    _cabi.call("p3d_group_bwd_prepare", ridx.data_ptr(), B, C, n, S, K, ws.data_ptr(), ws_b, st().cuda_stream)
    dfe = torch.empty((B, C, n), device=dev)
    us = timed(lambda: _cabi.call("p3d_group_bwd_f32", gout.data_ptr(), ridx.data_ptr(), B, C, n, S, K, dfe.data_ptr(),
                                  (C + 3) * S * K, 3, ws.data_ptr(), ws_b, 1, st().cuda_stream))
    put("grouping bwd SA2 K=128 C=320 (inverse index prepared)", us, gb_s=B * (4 * C * S * K + 8 * S * K + 4 * C * n) / us / 1e3)
    us = timed(lambda: ops.three_nn(xyz, pos1))
    put("three_nn n=2048 m=512", us, pair_tests_per_s=B * 2048 * 512 / (us * 1e-6))
    dist, i3 = ops.three_nn(xyz, pos1)
    w3 = torch.rand_like(dist)
    f3 = torch.randn((B, 128, 512), device=dev, generator=g)
    us = timed(lambda: ops.three_interpolate(f3, i3, w3))
    put("three_interpolate fwd C=128 512->2048", us, gb_s=B * (4 * 128 * 512 + 36 * 2048 + 4 * 128 * 2048) / us / 1e3)
    # C3: partial-dense search, voxel clustering, deformable KPConv
    pos_h, batch_h = synthetic.indoor_spheres()
    pos, batch = torch.from_numpy(pos_h).to(dev), torch.from_numpy(batch_h).to(dev)
    npts = pos.shape[0]
    us = timed(lambda: ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch,
                                      return_dist=False))
    pairs = float((torch.bincount(batch).double() ** 2).sum())
    put("ball_query partial n=%d r=0.1 M=34" % npts, us, pair_tests_per_s=pairs / (us * 1e-6), points_per_s=npts / (us * 1e-6))
    nb = ops.ball_query(0.1, 34, pos, pos, mode="partial_dense", batch_x=batch, batch_y=batch, return_dist=False)[0]
    N_ = pos.shape[0]
    cl = torch.empty(N_, dtype=torch.int64, device=dev)
    la = torch.empty(N_, dtype=torch.int64, device=dev)
    od = torch.empty(N_, dtype=torch.int32, device=dev)
    sg = torch.empty(N_ + 1, dtype=torch.int32, device=dev)
    mp = torch.empty(N_, 3, device=dev)
    ob = torch.empty(N_, dtype=torch.int64, device=dev)
    nd = torch.empty(1, dtype=torch.int64, device=dev)
    gb = _cabi.lib().p3d_grid_cluster_ws_bytes(N_)
    gws = torch.empty(gb, dtype=torch.uint8, device=dev)
    us = timed(lambda: _cabi.call("p3d_grid_cluster_f32", pos.data_ptr(), batch.data_ptr(), N_, 0.08, cl.data_ptr(),
                                  la.data_ptr(), od.data_ptr(), sg.data_ptr(), mp.data_ptr(), ob.data_ptr(), None,
                                  nd.data_ptr(), gws.data_ptr(), gb, st().cuda_stream))
    put("grid_cluster n=%d size=0.08 (GridSampling3D)" % npts, us, points_per_s=npts / (us * 1e-6), voxels=float(nd.item()))
    kp = torch.from_numpy(kpconv.load_kernels(0.06, random_rotation=False)).to(dev)
    xk = torch.randn((npts, 32), device=dev, generator=g, requires_grad=True)
    W = (torch.randn((15, 32, 32), device=dev, generator=g) * 0.1).requires_grad_(True)
    off = (torch.randn((npts, 15, 3), device=dev, generator=g) * 0.012).requires_grad_(True)
    dout = torch.randn((npts, 32), device=dev, generator=g)

    def dfwd():
        with torch.no_grad():
            return ops.KPConv_deform_fitting(pos, pos, nb, xk, kp, off, None, W, 0.04, "linear", "sum")[0]

    def dfb():
        out, kpmin, _ = ops.KPConv_deform_fitting(pos, pos, nb, xk, kp, off, None, W, 0.04, "linear", "sum")
        return torch.autograd.grad([out, kpmin], (xk, W, off), [dout, torch.ones_like(kpmin)])

    us = timed(dfwd, reps=5)
    put("kpconv_deform fwd 32->32 n=%d M=34 (fp32)" % npts, us, points_per_s=npts / (us * 1e-6))
    us = timed(dfb, reps=5)
    put("kpconv_deform fwd+bwd 32->32 n=%d M=34 (fp32)" % npts, us, points_per_s=npts / (us * 1e-6))
    return res


def bench_kpconv_stack(dev, world, rank):
    """C3 / C5: the KPConvPaper segmentation network (encoder + FPModule_PD decoder + per-point head, NLL loss,
    SGD momentum 0.98 -- models/segmentation/kpconv.py:106-140, conf/training/kpconv.yaml:8-16) on S3DIS-shaped spheres,
    ~50k points per rank, max_num_neighbors clipped to 34.  With more than one rank the model is wrapped in
    DistributedDataParallel (one NCCL gradient all-reduce per step).  Numbers: the step with the pyramid built on the
    device inside it; with a precomputed pyramid (the reference's DataLoader configuration); the precomputed step as
    ONE CUDA graph (forward + backward + all-reduce + SGD); and the all-reduce time that does not hide behind the
    backward pass (step with vs without gradient synchronisation)."""
    import torch
    import torch.distributed as dist
    from torch_points3d_b200 import segmentation
    torch.manual_seed(5)
    tr = segmentation.KPConvSegmentationTrainer(feat_dim=4, num_classes=13, device=dev, ddp=world > 1,
                                                seed=1003 + 97 * rank)
    n = int(tr.pos.shape[0])

    def timed(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms_ = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
        return float(ms_.item())

    loss0 = tr.step()
    ms_full = timed(lambda: tr.step_async(rebuild_geometry=True))
    ms_geo = timed(tr.build_pyramid)
    ms_pre = timed(lambda: tr.step_async(rebuild_geometry=False))
    loss1 = float(tr.step_async(rebuild_geometry=False))
    npts = torch.tensor([float(n)], device=dev)
    if world > 1:
        dist.all_reduce(npts)
    total_pts = float(npts.item())
    res = {"workload": "C3/C5: KPConvPaper segmentation network (5 KPDualBlocks down, 4 FPModule_PD up, %.1f M params), "
                       "~%d points/rank in 8 spheres, fwd+bwd+SGD%s" % (tr.num_params / 1e6, n,
                                                                       ", DDP over NCCL" if world > 1 else ""),
           # headline of this section: the training step in the reference's configuration -- the multi-scale pyramid comes
           # precomputed (the reference builds it in DataLoader workers: MultiScaleTransform); replaced below by the
           # CUDA-graph replay of the same step when the capture succeeds on every rank
           "ms_per_step": round(ms_pre, 3), "points_per_s": round(total_pts / (ms_pre * 1e-3), 1),
           "execution": "eager launches, precomputed pyramid",
           "eager_ms_per_step": round(ms_pre, 3),
           "eager_ms_per_step_with_on_device_geometry": round(ms_full, 3),
           "geometry_ms": round(ms_geo, 3), "precomputed_geometry_ms_per_step": round(ms_pre, 3),
           "loss_first": round(loss0, 4), "loss_after": round(loss1, 4), "unit": "points/s"}
    if world > 1:
        with_sync, no_sync = segmentation.exposed_allreduce_ms(tr)
        t = torch.tensor([with_sync, no_sync], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["allreduce"] = {"collective": "DistributedDataParallel gradient all-reduce (NCCL ring/NVLS over NVSwitch), "
                                          "%.1f MB fp32 in %d MB buckets" % (tr.num_params * 4 / 1e6, segmentation.DDP_BUCKET_MB),
                            "step_ms_with_sync": round(float(t[0]), 3), "step_ms_without_sync": round(float(t[1]), 3),
                            "exposed_ms": round(float(t[0] - t[1]), 3)}
        # no_sync steps let the ranks' parameters drift apart: put them back in step before the graph section
        for p_ in tr.model.parameters():
            dist.broadcast(p_.data, 0)
    ok = torch.ones(1, device=dev)
    try:
        tr.capture()
    except Exception as e:  # noqa: BLE001
        ok.zero_()
        res["cuda_graph_error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # replay only if every rank captured
    if float(ok.item()) > 0:
        ms_graph = timed(tr.replay)
        res["precomputed_geometry_cuda_graph_ms_per_step"] = round(ms_graph, 3)
        res["points_per_s_graph"] = round(total_pts / ((ms_graph + ms_geo) * 1e-3), 1)  # (pyramid build included)
        res["ms_per_step"] = round(ms_graph, 3)
        res["points_per_s"] = round(total_pts / (ms_graph * 1e-3), 1)
        res["execution"] = ("one CUDA graph per step: forward + NLL loss + backward%s + SGD; precomputed pyramid"
                            % (" + DDP bucketed all-reduce (NCCL)" if world > 1 else ""))
        res["graph_loss"] = round(float(tr.graph_loss.item()), 4)
    return res


def bench_c4(dev, world, rank):
    """C4 (BASELINE.json configs[3]): SemanticKITTI-shaped scans, FPS 4096 of 120 000 + knn k = 16, batch of 8 scans
    sharded by cloud over the ranks.  Timed twice: the two operators alone, and the RandLA-Net down-convolution chain
    the config names -- sampler -> KNNNeighbourFinder(16) -> RandlaKernel (BaseConvolutionDown.forward,
    core/base_conv/message_passing.py:44-58; modules/RandLANet/modules.py:9-67) with the chain's stock MLPs."""
    import torch
    import torch.distributed as dist
    from torch_points3d_b200 import message_passing as mp, ops, partial_dense as pd, spatial_ops, synthetic
    from torch_points3d_b200.sharding import cloud_range
    lo, hi = cloud_range(8, world, rank)
    scans_h = synthetic.lidar_scans(B=8)[lo:hi]
    scans = torch.from_numpy(scans_h).to(dev)
    b = hi - lo
    x = scans.reshape(-1, 3)
    bx = torch.arange(b, device=dev).repeat_interleave(120000)
    by = torch.arange(b, device=dev).repeat_interleave(4096)

    def step_ops():
        qi = ops.furthest_point_sample(scans, 4096)
        q = scans.gather(1, qi.unsqueeze(-1).expand(-1, -1, 3)).reshape(-1, 3)
        return ops.knn_table(x, q, 16, bx, by, return_dist=False)[0]

    torch.manual_seed(9)
    conv = mp.RandlaConv(ratio=4096.0 / 120000.0, k=16, sampler=spatial_ops.FPSSampler(num_to_sample=4096 * b, random_start=False),
                         point_pos_nn=[10, 8, 3], attention_nn=[6, 8, 6], down_conv_nn=[6, 8, 16]).to(dev).eval()
    data = pd._Data(x=x, pos=x, batch=bx)

    def step_chain():
        with torch.no_grad():
            return conv(data).x

    def timed(fn, warm=3, reps=10):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_fps = timed(lambda: ops.furthest_point_sample(scans, 4096))
    ms_ops = timed(step_ops)
    ms_chain = timed(step_chain)
    res = {"workload": "C4: FPS 4096 of 120000 + knn k=16, batch of 8 scans sharded over %d GPU(s)" % world,
           "ms_per_batch": round(ms_ops, 3), "clouds_per_s": round(8 / (ms_ops * 1e-3), 2),
           "fps_ms": round(ms_fps, 3), "knn_ms": round(ms_ops - ms_fps, 3),
           "us_per_fps_round": round(ms_fps * 1e3 / 4095, 3),
           "randla_chain_ms_per_batch": round(ms_chain, 3), "randla_chain_clouds_per_s": round(8 / (ms_chain * 1e-3), 2),
           "timing": "3 warm-ups, 10 repetitions, CUDA events, max over ranks"}
    if rank == 0:
        try:
            res["cpu_baseline"] = cpu_c4(scans_h[:1])
        except Exception as e:  # noqa: BLE001
            res["cpu_baseline"] = {"error": str(e)[:160]}
    return res


def cpu_c4(scan):
    """host-core baseline of C4 on ONE scan: FPS through the oracle (the ATen-loop restatement of torch_cluster's fps is
    ~6 s per scan) and knn through scipy's cKDTree (stand-in for the nanoflann kd-tree inside
    torch_cluster / torch-points-kernels, which is not installable here), all host threads"""
    import oracle
    from scipy.spatial import cKDTree
    oracle.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    idx = oracle.fps_dense(scan, 4096)
    t_fps = time.perf_counter() - t0
    q = np.take_along_axis(scan, idx[..., None], 1)[0]
    t0 = time.perf_counter()
    tree = cKDTree(scan[0])
    tree.query(q, k=16, workers=-1)
    t_kd = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle.knn_csr(scan[0], q, 16, np.array([0, scan.shape[1]]), np.array([0, 4096]))
    t_bf = time.perf_counter() - t0
    total = t_fps + t_kd
    return {"value": round(1.0 / total, 3), "unit": "clouds/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "1 of the 8 scans: oracle FPS %.3f s (one thread per cloud) + cKDTree build+query %.3f s "
                      "(workers=-1; kd-tree stand-in for nanoflann); oracle brute-force knn %.3f s" % (t_fps, t_kd, t_bf)}


def bench_c1(dev):
    """C1 (BASELINE.json configs[0], the reference's own CPU-runnable case): FPS 512 of 1024 + ball_query r = 0.2 k = 32
    + grouping of 64 feature channels, batch of 16 clouds -- the dense SA-layer chain of core/base_conv/dense.py:60-87"""
    import torch
    from torch_points3d_b200 import ops, synthetic
    xyz_h = synthetic.surface_clouds(16, 1024, 1001)
    xyz = torch.from_numpy(xyz_h).to(dev)
    feat = torch.randn((16, 64, 1024), device=dev, generator=torch.Generator(device=dev).manual_seed(1001))

    def step():
        idx, new_xyz = ops.furthest_point_sample(xyz, 512, return_xyz=True)
        ridx = ops.ball_query(0.2, 32, xyz, new_xyz, return_dist=False)[0]
        return ops.group_features(feat, xyz, new_xyz, ridx)

    for _ in range(3):
        step()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = step()  # noqa: F841
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"workload": C1_WORKLOAD, "ms_per_step": round(ms, 4), "clouds_per_s": round(16 / (ms * 1e-3), 1),
            "unit": "clouds/s", "cpu_baseline": cpu_c1(4.0)}


C1_WORKLOAD = "C1: PointNet++ SA-layer op chain forward: FPS 512 of 1024 + ball_query r=0.2 k=32 + grouping C=64+3, batch 16"


def cpu_c1(max_seconds):
    import oracle
    from torch_points3d_b200 import synthetic
    oracle.set_num_threads(os.cpu_count() or 1)
    xyz = synthetic.surface_clouds(16, 1024, 1001)
    feat = np.random.default_rng(1001).standard_normal((16, 64, 1024)).astype(np.float32)
    pos_t = np.ascontiguousarray(xyz.transpose(0, 2, 1))

    def once():
        idx = oracle.fps_dense(xyz, 512)
        new = np.take_along_axis(xyz, idx[..., None], 1)
        ridx, _ = oracle.ball_query_dense(0.2, 32, xyz, new)
        gp = oracle.group_fwd(pos_t, ridx)
        gp -= new.transpose(0, 2, 1)[..., None]
        oracle.group_fwd(feat, ridx)

    once()
    t0, done = time.perf_counter(), 0
    while time.perf_counter() - t0 < max_seconds:
        once()
        done += 16
    dt = time.perf_counter() - t0
    return {"value": round(done / dt, 2), "unit": "clouds/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d clouds of the C1 chain through oracle/p3d_oracle.c (%.1f s)" % (done, dt)}


def bench_c2_model(dev, chain_ms):
    """BASELINE.json configs[1] as the reference runs it: the whole `pointnet2_charlesmsg` network
    (conf/models/segmentation/pointnet2.yaml:95-130; models/segmentation/pointnet2.py:90-113), forward + cross-entropy
    on 50 classes + backward + Adam step, batch 16 x 2048 points with 3 normals.  Operators from libp3d_b200.so, MLPs /
    BatchNorm / max-pool stock cuDNN as in the reference.  Places the operator chain against the only published
    figure: `iteration=0.231 s` per batch of 16 (benchmark/shapenet/pointnet2_charlesmsg.md:211, hardware not stated)."""
    import torch
    from torch_points3d_b200 import pointnet2, synthetic
    torch.manual_seed(12)
    pos_h = synthetic.surface_clouds(B_CLOUDS, N_PTS, 1002)
    pos = torch.from_numpy(pos_h).to(dev)
    nrm = pos / (pos.norm(dim=-1, keepdim=True) + 1e-12)
    labels = torch.randint(0, 50, (B_CLOUDS, N_PTS), device=dev)
    model = pointnet2.PointNet2_D(feat_dim=3, num_classes=50).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model.loss(model(pos, nrm), labels)
        loss.backward()
        opt.step()
        return loss.detach()

    losses = [float(step()) for _ in range(5)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"workload": "C2 full model: pointnet2_charlesmsg fwd + cross-entropy + bwd + Adam, batch 16 x 2048 (eager)",
            "ms_per_step": round(ms, 3), "clouds_per_s": round(B_CLOUDS / (ms * 1e-3), 1),
            "loss_first": round(losses[0], 4), "loss_fifth": round(losses[-1], 4),
            "operator_chain_ms": chain_ms, "operator_share_of_step": round(chain_ms / ms, 3),
            "published_reference": {"iteration_s": 0.231, "clouds_per_s": round(16 / 0.231, 1), "hardware": "not stated",
                                    "source": "benchmark/shapenet/pointnet2_charlesmsg.md:211"},
            "speedup_vs_published_iteration": round(231.0 / ms, 1)}


def bench_rsconv_model(dev):
    """The RSConv segmentation network the reference benchmarks (`RSConv_MSN`, conf/models/segmentation/rsconv.yaml:57-108;
    models/segmentation/rsconv.py:14-124): forward + cross-entropy on 50 classes + backward + Adam, batch 12 x 2048
    points with 3 normals -- the batch size of the published run (`iteration=0.133 s`, benchmark/shapenet/rscnn_original.md,
    hardware not stated).  Same operators as the C2 chain (FPS, multi-radius ball query, fused grouping, three_nn /
    three_interpolate); 1x1 convolutions / BatchNorm stock cuDNN as in the reference."""
    import torch
    from torch_points3d_b200 import rsconv, synthetic
    torch.manual_seed(13)
    clouds = 12
    pos = torch.from_numpy(synthetic.surface_clouds(clouds, N_PTS, 1003)).to(dev)
    nrm = pos / (pos.norm(dim=-1, keepdim=True) + 1e-12)
    labels = torch.randint(0, 50, (clouds, N_PTS), device=dev)
    model = rsconv.RSConv_MSN(feat_dim=3, num_classes=50).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model.loss(model(pos, nrm), labels)
        loss.backward()
        opt.step()
        return loss.detach()

    losses = [float(step()) for _ in range(5)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"workload": "RSConv_MSN full model: fwd + cross-entropy + bwd + Adam, batch 12 x 2048 (eager)",
            "ms_per_step": round(ms, 3), "clouds_per_s": round(clouds / (ms * 1e-3), 1),
            "parameters": int(sum(p.numel() for p in model.parameters())),
            "loss_first": round(losses[0], 4), "loss_fifth": round(losses[-1], 4),
            "published_reference": {"iteration_s": 0.133, "clouds_per_s": round(12 / 0.133, 1), "hardware": "not stated",
                                    "source": "benchmark/shapenet/rscnn_original.md"},
            "speedup_vs_published_iteration": round(133.0 / ms, 1)}


# ------------------------------------------------------------ CPU arm (oracle port) -----------
def cpu_chain_once(h, clouds):
    """the same chain on host cores with the oracle (oracle/p3d_oracle.c, OpenMP over clouds/queries)"""
    import oracle
    from torch_points3d_b200 import synthetic
    pos = h["pos"][:clouds]
    feats = [h["x0"][:clouds], h["x1"][:clouds]]
    levels = [pos]
    for li, l in enumerate(synthetic.C2_SA):
        idx = oracle.fps_dense(pos, l["npoint"])
        new_pos = np.take_along_axis(pos, idx[..., None], 1)
        pos_t = np.ascontiguousarray(pos.transpose(0, 2, 1))
        for r, k in zip(l["radii"], l["nsample"]):
            ridx, _ = oracle.ball_query_dense(r, k, pos, new_pos)
            gp = oracle.group_fwd(pos_t, ridx)
            gp -= new_pos.transpose(0, 2, 1)[..., None]
            gf = oracle.group_fwd(feats[li], ridx)
            oracle.group_bwd(gf, ridx, pos.shape[1])      # gradient of the gather (dout := gf, same bytes)
        pos = new_pos
        levels.append(pos)
    for unknown, known, x in [(levels[1], levels[2], h["x2"][:clouds]), (levels[0], levels[1], h["x3"][:clouds])]:
        dist, idx = oracle.three_nn(unknown, known)
        recip = 1.0 / (dist + 1e-8)
        w = (recip / recip.sum(2, keepdims=True)).astype(np.float32)
        out = oracle.three_interpolate_fwd(x, idx, w)
        oracle.three_interpolate_bwd(out, idx, w, known.shape[1])
    return clouds


def cpu_baseline(max_seconds=12.0, clouds=B_CLOUDS):
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use all host cores anyway
    h = make_host_inputs(1002)
    cpu_chain_once(h, 1)  # warm (builds / loads the oracle)
    t0 = time.perf_counter()
    done = 0
    while True:
        done += cpu_chain_once(h, clouds)
        dt = time.perf_counter() - t0
        if dt > max_seconds:
            break
    return {"value": round(done / dt, 3), "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d clouds of the same C2 chain through oracle/p3d_oracle.c (OpenMP, %d threads, %.1f s)"
                      % (done, oracle.num_threads(), dt)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    reference's own arithmetic lives in un-vendored wheels (torch-points-kernels, torch-cluster)
    that cannot be installed offline, so this times the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle
    from oracle import build as oracle_build
    oracle_build.build()
    oracle.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use all host cores anyway
    h = make_host_inputs(1002)
    clouds = B_CLOUDS  # the whole 16-cloud batch per step; the number of steps is bounded below
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_chain_once(h, clouds)
    steps = min(args.steps, 50)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_chain_once(h, clouds)
    dt = time.perf_counter() - t0
    value = round(clouds * steps / dt, 3)
    sample = "%d of the 16 clouds per step, %d steps, oracle/p3d_oracle.c with %d OpenMP threads" % (
        clouds, steps, oracle.num_threads())
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clouds_per_gpu": B_CLOUDS, "points_per_cloud": N_PTS,
                       "parallelism": "host cores (the same batch of 16 clouds per step, %d steps)" % steps},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    try:  # the reference's own CPU-runnable case (BASELINE.json configs[0]) beside the headline
        line["c1"] = {"workload": C1_WORKLOAD, **cpu_c1(4.0)}
    except Exception as e:  # noqa: BLE001
        line["c1"] = {"error": str(e)[:160]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph replay")
    ap.add_argument("--skip-extras", action="store_true",
                    help="profiling runs: skip the KPConv section and the CPU baseline")
    ap.add_argument("--extras-timeout", type=int, default=420,
                    help="seconds the secondary sections may take before the measured line is printed without them")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
