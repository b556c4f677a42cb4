#!/usr/bin/env python
"""Benchmark of the X-GGM graph-generative block (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload block|iteration|visn|eval] [--branch node|relation|mixed] [--precision fp32|bf16]

Default workload (`block`, the headline).  One "step" = one pass of the hot path over one batch: the GGM
node-generation branch the shipped VQA-CP v2 recipe always takes (--delta 0, script/vqacpv2.sh:22 of the reference):
strip_diag(adj_true) -> node_fc(x) (+36-fold broadcast) -> Gaussian feature noise -> GCNGenerator(L=2) ->
symmetric-KL + score-matching losses -> fusion_fc read-out, forward AND backward down to every parameter gradient
and the gradients of the LXMERT outputs that feed the block, then the gradient exchange (N > 1), clip_grad_norm_(5.)
and the BertAdam update of the block's parameters.  The LXMERT encoder / answer head that surround the block are
outside the hot path; their place is taken by resident inputs and a fixed cotangent on x_gen.

Per-GPU batch B=256 (BASELINE configs[1]), N=36, H=768, fp32.  N>1 ranks = data parallel (weak scaling): each rank
runs its own B=256 shard; the gradient averaging, the clip and the update run as one fused sequence over NVLink peer
memory (xggm_dp_bertadam_step; XGGM_DP_FUSED=0: one NCCL all-reduce + separate clip / update), inside the captured step.

Prints ONE JSON line (rank 0) with `roofline`, `cpu_baseline`, `eager_gpu_baseline` (the reference's own eager PyTorch
path on the same GPU), `e2e`, `clocks`, `gpu_launches`.  `--impl reference` times the reference's CPU path on the host
cores instead (the unmodified reference modules vendored in oracle/_ref, else the oracle port; rank 0 only).
Other workloads: `iteration` = the full VQA-CP v2 trainer iteration around a stock-PyTorch LXMERT (BASELINE
configs[1] / [2]); `visn` = the VisualFeatEncoder leg; `eval` = BASELINE configs[4].
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_NODES, HID, N_LAYERS, SIGMA, NUM_ANS, GNN = 256, 36, 768, 2, 1.0, 2274, "GCN"
CPU_SAMPLE_B = 32
METRIC = "xggm_graph_block_train_samples_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel -- the grouped forward
# projection {P = h Wc^T (operand planes out) ; z = h W0^T + b (fp32 out)}, two [9216,768]x[768,768] members, CTA-pair
# kernel, prepared weight planes -- from the committed `ncu --set full` capture
# profiles/r02m_ncu_gemm_group_fwd_and_adj_apply.txt: 34.88 MB read + 15.29 MB written back.  Algorithmic bytes of that
# launch: 28.3 MB A planes (shared by both members) + 2 x 2.4 MB W planes + 28.3 MB P planes + 28.3 MB z = 89.6 MB; the
# DRAM traffic is BELOW it because the A planes arrive in L2 from the producing kernel and most of the output is still in
# the 126 MB L2 when the kernel ends.
TRAFFIC_NCU = 34881536 + 15287296


def algorithmic_flops_per_sample(N=N_NODES, H=HID, L=N_LAYERS):
    """F_fwd = L(5*2NH^2 + 3*2N^2H); fwd+bwd = 3x."""
    return 3 * L * (5 * 2 * N * H * H + 3 * 2 * N * N * H)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs.  Primary: an in-process NVML thread
    (nvidia_ml_py) polling every 2 ms -- the timed region of the default run is ~80 ms, a `nvidia-smi -lms` stream
    only gets a handful of lines out in that time (`clocks.samples: 1`).  Fallback: one streaming
    `nvidia-smi -lms 25` process.  Only rank 0 samples (eight pollers on one host cost 1.3 ms/step in round 1)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None
        self.nvml, self.stop, self.src = None, False, None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self.nvml is not None:
            nv, h = self.nvml
            names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
            try:
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                while not self.stop:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    self.rows.append((time.time(), [str(sm), str(mx)] + ["Active" if mask & bit else "Not Active" for _, bit in names]))
                    time.sleep(0.002)
            except Exception:
                pass
            return
        try:
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.strip().split(",")]))
        except Exception:
            pass

    def start(self):
        """Start sampling (call a little before the region so that the first samples are not lost)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self.nvml = (nv, nv.nvmlDeviceGetHandleByIndex(phys))
            self.src = "nvml thread, 2 ms period"
            self.t.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.src = "nvidia-smi -lms 25"
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.t1 = time.time()
        time.sleep(0.05)   # let the line of the last in-region sample arrive
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.t.is_alive():
            self.t.join(timeout=6)

    def summary(self):
        inside = [r for t, r in self.rows if self.t0 <= t <= self.t1 + 0.03]
        rows = inside if inside else [r for _, r in self.rows[-1:]]   # (never empty-handed: fall back to the nearest sample)
        sm = sorted(int(r[0]) for r in rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.src}


# ------------------------------------------------------------------------------- CPU / eager baselines
def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_steps(steps, warmup, threads, B=CPU_SAMPLE_B, branch="node", gnn=GNN):
    """The reference's PyTorch CPU path for the same step: the UNMODIFIED reference modules vendored into
    oracle/_ref by oracle/vendor_ref.py (kind "reference"), else the oracle port (kind "port").
    Returns (seconds per step (median), kind)."""
    from oracle.ref_step import ReferenceStep
    torch.set_num_threads(threads)
    ref = ReferenceStep("cpu", B, branch, gnn, HID, N_LAYERS, N_NODES, SIGMA, NUM_ANS, lr=4e-6)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ref.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], ref.kind


def eager_gpu_steps(dev, steps, warmup, B, branch="node", gnn=GNN):
    """The reference's eager PyTorch path ON THE B200 (SURVEY 2a / BASELINE.md section 1, bar (i)): same modules, same
    step, device cuda, true fp32 (allow_tf32 off -- the torch default the reference runs with), CUDA-event timed.
    Returns (ms per step, kind)."""
    from oracle.ref_step import ReferenceStep
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = ReferenceStep(dev, B, branch, gnn, HID, N_LAYERS, N_NODES, SIGMA, NUM_ANS, lr=4e-6)
        for _ in range(warmup):
            ref.step()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s_, e_ in evs:
            s_.record()
            ref.step()
            e_.record()
        torch.cuda.synchronize()
        return sum(s_.elapsed_time(e_) for s_, e_ in evs) / steps, ref.kind
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sec, kind = cpu_reference_steps(args.steps, args.warmup, threads, branch=args.branch, gnn=args.gnn)
    val = CPU_SAMPLE_B / sec
    other = "relation" if args.branch == "node" else "node"
    sec_other, _ = cpu_reference_steps(3, 1, threads, branch=other, gnn=args.gnn)
    sec_1t, _ = cpu_reference_steps(2, 1, 1, branch=args.branch, gnn=args.gnn)
    sample = (f"B={CPU_SAMPLE_B} graphs per step (BASELINE configs[0]; a bounded sample of the B={B_PER_GPU}/GPU workload), "
              f"{args.steps} timed steps after {args.warmup} warm-up, median; eager PyTorch on the host, {threads} threads")
    cfg = workload_config(args.gpus, gnn=args.gnn, branch=args.branch)
    cfg["workload"] = cfg["workload"].replace("fp32 (tensor cores, 3 split-bf16 passes)", "fp32")
    cfg.update({"arm": "reference CPU path: eager PyTorch on the host cores, no CUDA graph, fp32",
                "per_step_batch": CPU_SAMPLE_B, "device": "cpu", "threads": threads, "cpu_model": _cpu_model(),
                "l2": "n/a (CPU)", "launch": "eager PyTorch ops from Python"})
    line = {"impl": "reference", "kind": kind, "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample,
                             "branch": args.branch, "other_branch": {"branch": other, "value": CPU_SAMPLE_B / sec_other,
                                                                     "ms_per_step": sec_other * 1e3},
                             "one_thread": {"value": CPU_SAMPLE_B / sec_1t, "ms_per_step": sec_1t * 1e3}},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, B=B_PER_GPU, gnn=GNN, precision="fp32", branch="node"):
    arith = {"fp32": "fp32 (tensor cores, 3 split-bf16 passes)", "bf16": "bf16 tensor cores, fp32 storage",
             "fp32_simt": "fp32 FMA"}[precision]
    which = {"node": "VQA-CP v2 GGM node branch (delta=0)", "relation": "GGM relation branch (GQA-OOD weights)",
             "mixed": "GQA-OOD recipe: relation / node branch drawn per step by a rank-synchronous BranchSchedule (delta=5)"}[branch]
    return {"workload": f"cfg2 graph block: {which}, {gnn}Generator L={N_LAYERS}, fwd+bwd + clip_grad_norm_(5) + BertAdam, "
                        f"B={B}/GPU, N={N_NODES}, H={HID}, sigma={SIGMA}, A={NUM_ANS}, {arith}",
            "global_batch": B * n_gpus, "per_gpu_batch": B, "parallelism": f"dp{n_gpus}",
            "l2": "flushed between timed steps (256 MiB write, outside the per-step CUDA-event pairs)",
            "launch": "one CUDA-graph replay per step (xggm_b200.GraphedStep)"}



def synthetic_inputs(seed, B, n_nodes=N_NODES, hidden=HID):
    """Synthetic block inputs in the shapes of SURVEY.md section 8d (the same recipe as the oracle's input factory,
    restated here so that the measured arm does not touch oracle/): visn = layer_norm(N(0,1)) -- LXMERT's visual
    output is a LayerNorm output --, pooled = tanh(N(0,1)), adj_true = obj36_adj look-alike (symmetrised
    U(-0.2, 1), max-normalised, non-zero diagonal)."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, n_nodes, hidden, generator=g, dtype=torch.float64)
    v = (v - v.mean(-1, keepdim=True)) / v.std(-1, unbiased=False, keepdim=True)
    xp = torch.tanh(torch.randn(B, hidden, generator=g, dtype=torch.float64))
    c = torch.rand(B, n_nodes, n_nodes, generator=g, dtype=torch.float64) * 1.2 - 0.2
    c = c + c.transpose(1, 2)
    c = c / c.amax(dim=(1, 2), keepdim=True)
    return v.float(), xp.float(), c.float()


def pin_to_gpu_numa(local):
    """Bind this rank to the cores of the NUMA node its GPU hangs off (pinned host buffers are then allocated there and
    the per-step H2D copies of eight ranks do not all cross the inter-socket link).  Best effort: silently does nothing
    where sysfs does not expose the topology.  XGGM_BENCH_NO_NUMA=1 disables it."""
    if os.environ.get("XGGM_BENCH_NO_NUMA") == "1":
        return None
    try:
        pr = torch.cuda.get_device_properties(local)
        dom = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dom}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    import xggm_b200 as X
    from xggm_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(9595)  # reference default seed, src/param.py:49 (same on every rank: same branch, same init)

    B = args.batch
    global N_NODES
    N_NODES = args.nodes
    X.set_precision(args.precision)
    model = X.XGGMHeads(HID, args.gnn, N_LAYERS, N_NODES).to(dev).train()
    from xggm_b200.ddp import FlatGrads
    # .grad are views of one flat buffer (no staging copy); N > 1: one NCCL all-reduce (AVG) at the end of the
    # step, inside the captured graph.  XGGM_DDP_OVERLAP=1 instead ships the bucket of the modules whose backward
    # runs first (fusion_fc, last generator layer) on a side stream during the first layer's backward -- measured
    # equal at N=2 (2.013 vs 2.015 ms/step): the persistent one-CTA-per-SM GEMMs leave NCCL no SMs to overlap on.
    early = []
    if os.environ.get("XGGM_DDP_OVERLAP") == "1":
        early = list(model.fusion_fc.parameters()) + list(model.generator.gnn_layers[-1].parameters())
    # N > 1: the gradient bucket (and the parameter buffer) live in NVLink-addressable symmetric memory and the
    # all-reduce + clip + BertAdam run as the fused peer-memory sequence (xggm_dp_bertadam_step); XGGM_DP_FUSED=0 keeps
    # the NCCL all-reduce + separate clip / update kernels for A/B runs
    want_fused = world > 1 and os.environ.get("XGGM_DP_FUSED", "1") != "0" and not early
    grads = FlatGrads(model.parameters(), early=early, symmetric=want_fused)
    flat_grad = grads.flat
    # BertAdam over the block's parameters as the trainer configures it for the down-task group (4 * --lr with
    # --lr 1e-6, script/vqacpv2.sh:24, src/vqa/vqacpv2.py:125-128; constant schedule so the captured step stays
    # valid), preceded by clip_grad_norm_(., 5.) (src/vqa/vqacpv2.py:252)
    optim = X.BertAdam(model.parameters(), lr=4e-6, flat_grads=grads)
    fused_dp = want_fused and optim.fused_allreduce_available()

    visn_h, xp_h, adj_h = (t.pin_memory() for t in synthetic_inputs(9596 + rank, B, N_NODES, HID))
    cot_h = torch.randn(B, HID, generator=torch.Generator().manual_seed(2 + rank)).pin_memory()
    visn_d, xp_d, adj_d, cot_d = (t.to(dev) for t in (visn_h, xp_h, adj_h, cot_h))
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    loss_host = torch.zeros(1).pin_memory()

    w_rel, w_node = torch.tensor(6.0, device=dev), torch.tensor(1.1, device=dev)   # vqacpv2.py:221,250

    def compute(visn, xp, adj):  # forward + backward of the block; gradients land in the flat bucket
        grads.zero_()
        x = xp.requires_grad_(True)
        feat = visn.requires_grad_(True)
        with grads.overlap(average=True):
            if args.branch == "relation":   # GQA-OOD weights, src/gqa/gqa_ood.py:197 ("mixed" captures both)
                x_gen, loss_sm, _, _ = model.relation_step(x, feat, adj, SIGMA, NUM_ANS, kl_weight=12.0)
                w_sm = w_rel
            else:
                x_gen, loss_sm, _, _ = model.node_step(x, feat, adj, SIGMA, NUM_ANS)
                w_sm = w_node
            # loss = <x_gen, cot> + w * loss_sm, where cot stands for d BCE(logit_fc(x_gen)) / d x_gen of the answer head
            # that follows the block (outside the hot path): its backward pass is seeded directly with (cot, w)
            torch.autograd.backward([x_gen, loss_sm], [cot_d, w_sm])
        if fused_dp:
            optim.step_allreduce(5.0)       # reduce-scatter + clip + BertAdam (1/N of it) + parameter all-gather over NVLink
        else:
            grads.all_reduce(average=True)  # NCCL gradient all-reduce (no-op at world size 1)
            optim.step(X.clip_grad_norm_(grads, 5.0))   # one norm reduction + one fused update kernel
        return loss_sm.detach()

    # the public entry point for a captured step: one CUDA-graph launch per step (xggm_b200.GraphedStep)
    graphed = None
    launches_per_step = None
    mixed = None
    if not args.no_graph:
        for _ in range(2):
            compute(visn_d.detach(), xp_d.detach(), adj_d)
        l0 = _lib.kernel_launches()
        compute(visn_d.detach(), xp_d.detach(), adj_d)
        launches_per_step = _lib.kernel_launches() - l0
        graphed = X.GraphedStep(compute, [visn_d, xp_d, adj_d])
        if args.branch == "mixed":
            # GQA-OOD recipe (--delta 5, script/gqa_ood.sh:21): a rank-synchronous BranchSchedule draws the branch of every
            # step; one captured graph per branch, the schedule picks the graph to replay
            from xggm_b200.ddp import BranchSchedule
            sched = BranchSchedule(args.delta if args.delta > 0 else 5)
            args.branch = "relation"
            for _ in range(2):
                compute(visn_d.detach(), xp_d.detach(), adj_d)
            g_rel = X.GraphedStep(compute, [visn_d, xp_d, adj_d])
            args.branch = "mixed"
            mixed = (sched, {"node": graphed, "relation": g_rel})

    def resident_step():
        if mixed is not None:
            l = mixed[1][mixed[0].next()].replay()
        elif graphed is not None:
            l = graphed.replay()
        else:
            l = compute(visn_d.detach(), xp_d.detach(), adj_d)
        return l

    def e2e_step():
        if mixed is not None:
            g_ = mixed[1][mixed[0].next()]
            l = g_(visn_h.to(dev, non_blocking=True), xp_h.to(dev, non_blocking=True), adj_h.to(dev, non_blocking=True))
        elif graphed is not None:
            # every step: this step's inputs were put on the wire (pinned host -> device staging, side
            # stream) while the previous step computed; move them in, replay, start the next step's H2D
            l = graphed.run_prefetched()
            graphed.prefetch(visn_h, xp_h, adj_h)
        else:
            l = compute(visn_h.to(dev, non_blocking=True), xp_h.to(dev, non_blocking=True),
                        adj_h.to(dev, non_blocking=True))
        loss_host.copy_(l.reshape(1), non_blocking=True)

    def timed(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for s, e in evs:
            flush.fill_(1.0)  # L2 flush, outside the event pair
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            dist.barrier()
        return ms

    if graphed is not None:
        graphed.prefetch(visn_h, xp_h, adj_h)
    for _ in range(max(args.warmup, 3)):
        resident_step()
        e2e_step()
    torch.cuda.synchronize()

    # only rank 0 polls (its line is the one printed): eight 50 Hz NVML pollers on one host slow every rank down
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("XGGM_BENCH_NO_CLOCKS") != "1":
        sampler.start()
    for _ in range(3):       # keep the GPU busy while the sampler process starts
        resident_step()
    torch.cuda.synchronize()
    with sampler as clk:
        l0 = _lib.kernel_launches()
        ms_res = timed(resident_step, args.steps)
        launches = _lib.kernel_launches() - l0
        if graphed is not None:  # replayed kernels are not launched from the host: count them from the capture
            launches = launches_per_step * args.steps
        ms_e2e = timed(e2e_step, args.steps)
    clocks = clk.summary()

    # per-launch timing of the dominant kernel (projection GEMM), CUDA events on its own stream
    _lib.gemm_profile(True)
    prof_steps = min(args.steps, 3)
    for _ in range(prof_steps):
        flush.fill_(1.0)
        compute(visn_d.detach(), xp_d.detach(), adj_d)   # eager launches: the per-kernel events need them
    torch.cuda.synchronize()
    g_ms, g_n, g_flops = _lib.gemm_profile()
    _lib.gemm_profile(False)

    # secondary leg: the same step on the single-pass bf16 engine (BASELINE cfg 3 arithmetic, 2e-2 budget)
    alt = None
    if args.precision == "fp32" and graphed is not None and not args.quick:
        X.set_precision("bf16")
        for _ in range(2):
            compute(visn_d.detach(), xp_d.detach(), adj_d)
        graphed_bf16 = X.GraphedStep(compute, [visn_d, xp_d, adj_d])

        def bf16_step():
            graphed_bf16.replay()

        for _ in range(3):
            bf16_step()
        ms_alt = timed(bf16_step, args.steps)
        X.set_precision("fp32")
        alt = {"dtype": "bf16", "ms_per_step": ms_alt / args.steps, "value": B * world / (ms_alt / args.steps * 1e-3),
               "unit": "samples/s", "note": "same step, projection engine set to single-pass bf16 tensor cores"}

    def shutdown():
        # the captured steps contain NCCL kernels: tearing the communicator down under them can hang, and
        # nothing is left to clean up in a benchmark process -> flush and leave
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        shutdown()
        return
    peaks = load_peaks()
    ms_step = ms_res / args.steps
    value = B * world / (ms_step * 1e-3)
    e2e_val = B * world / (ms_e2e / args.steps * 1e-3)
    flops_step = algorithmic_flops_per_sample(N_NODES) * B
    if args.gnn == "GIN":   # L(3*2NH^2 + 2*2N^2H) fwd, x3
        flops_step = 3 * N_LAYERS * (3 * 2 * N_NODES * HID * HID + 2 * 2 * N_NODES * N_NODES * HID) * B
    passes = 1 if args.precision == "bf16" else 3
    gemm_tflops = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    peak = peaks["bf16_tflops"]
    h2d = sum(t.numel() * 4 for t in (visn_h, xp_h, adj_h))
    threads = os.cpu_count() or 1
    if args.quick or world > 1:      # sweep lines (tools/sweep.py) and multi-GPU runs: the GPU arm only (baselines: N = 1)
        cpu_sec, cpu_kind, eager_ms, eager_kind = float("inf"), "skipped", float("inf"), "skipped"
    else:
        cpu_sec, cpu_kind = cpu_reference_steps(5, 2, threads, branch=args.branch, gnn=args.gnn)
        eager_ms, eager_kind = eager_gpu_steps(dev, 10, 3, B, branch=args.branch, gnn=args.gnn)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": dict(workload_config(world, B, args.gnn, args.precision, args.branch),
                       collective=("none (one GPU)" if world == 1 else
                                   ("fused reduce-scatter + clip + BertAdam + parameter all-gather over NVLink peer memory "
                                    "(xggm_dp_bertadam_step; "
                                    + ("NVSwitch multicast: multimem.ld_reduce / multimem.st" if getattr(flat_grad, "_xggm_multicast_ptr", 0)
                                       else "unicast peer loads / stores") + ")"
                                    if fused_dp else "NCCL all-reduce (AVG) of the flat gradient bucket"))),
        "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "numa_node_of_rank0": numa},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor",
                     "kernel": "gemm_tc_kernel (768x768 node projections fwd/dgrad/wgrad; tcgen05 kind::f16, "
                               "fp32 parity via 3 split-bf16 passes => 3x the algorithmic FLOPs are executed)",
                     "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s", "frac": gemm_tflops / peak,
                     "executed_tflops": passes * gemm_tflops, "executed_frac": passes * gemm_tflops / peak,
                     "traffic": TRAFFIC_NCU,
                     "traffic_of": "one grouped forward launch (2 projections: 2 x 10.87 GFLOP algorithmic; 89.6 MB algorithmic "
                                   "bytes), ncu --set full, profiles/r02m_ncu_gemm_group_fwd_and_adj_apply.txt",
                     "peak_source": f"{peaks['source']} bf16 burst (MEASURED_PEAKS.json)",
                     "launches_timed": g_n, "avg_launch_us": (g_ms / g_n * 1e3) if g_n else None,
                     "gemm_share_of_step": (g_ms / prof_steps) / ms_step if ms_step > 0 else None,
                     "note": "launches of the dominant shape [B*N,768]x[768,768] (fwd, dgrad, wgrad), timed eagerly "
                             "with CUDA events on the launching stream inside the library"},
        "block_roofline": {"algorithmic_gflop_per_step": flops_step / 1e9,
                           "achieved_tflops": flops_step / (ms_step * 1e-3) / 1e12,
                           "frac_of_bf16_peak": flops_step / (ms_step * 1e-3) / 1e12 / peak,
                           "t_roof_us": flops_step / (peak * 1e12) * 1e6},
        "bf16_engine": alt,
        "cpu_baseline": None if cpu_kind == "skipped" else
                        {"value": CPU_SAMPLE_B / cpu_sec, "unit": "samples/s", "cores": threads, "kind": cpu_kind,
                         "sample": f"B={CPU_SAMPLE_B} graphs/step, 5 timed steps after 2 warm-up, median"},
        # the reference's own eager PyTorch path on this B200 (same step, same B, true fp32, no CUDA graph)
        "eager_gpu_baseline": None if eager_kind == "skipped" else
                              {"value": B / (eager_ms * 1e-3), "unit": "samples/s", "ms_per_step": eager_ms,
                               "kind": eager_kind, "dtype": "f32 (allow_tf32=False)", "n_gpus": 1,
                               "sample": f"B={B} graphs/step on rank 0's GPU, 10 timed steps after 3 warm-up, CUDA events"},
    }
    print(json.dumps(line), flush=True)
    shutdown()


# ---------------------------------------------------------------------------- full training iteration
ITER_METRIC = "xggm_train_iteration_samples_per_sec"


def run_iteration(args):
    """BASELINE configs[1] (fp32) / configs[2] (--precision bf16): one full trainer iteration -- stock-PyTorch LXMERT
    (tools/lxmert_torch.py, shared by every arm, outside the hot path) + the library's graph block, answer head,
    BCE, clip and BertAdam; two optimiser steps, two all-reduces of the 220.8 M-parameter buckets (2 x 0.88 GB) per
    iteration at N > 1.  See tools/iteration.py."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import iteration as IT
    from xggm_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    autocast = args.precision == "bf16"
    torch.backends.cuda.matmul.allow_tf32 = False      # the reference's (torch default) true-fp32 matmuls
    torch.backends.cudnn.allow_tf32 = False
    it = IT.XGGMIteration(dev, B, autocast=autocast, delta=args.delta)
    batches = [IT.synthetic_batch(9596 + 17 * rank + i, B) for i in range(2)]
    h2d = it.h2d_bytes(batches[0])
    loss_host = torch.zeros(1).pin_memory()
    it.prefetch(batches[0])
    resident = [t.to(dev) for t in batches[0]]
    import xggm_b200 as X
    graphs = None
    if not args.no_graph:
        # one captured graph per GGM branch the schedule can draw (delta = 0: node only); X.GraphedStep keeps the
        # library's dropout fresh across replays (device epoch) and torch's own RNG is graph-aware
        branches = ["node"] if args.delta <= 0 else (["relation"] if args.delta >= 10 else ["node", "relation"])
        graphs = {b: X.GraphedStep(lambda *t, b=b: it.iteration(list(t), b), resident, warmup=2) for b in branches}

    def run_graph(src):
        g = graphs[it.branch.next()]
        return g(*src)

    def resident_iter():
        if graphs is not None:
            return run_graph(resident)
        return it.iteration(resident)

    k = [0]

    def e2e_iter():
        cur = it.take()                      # this iteration's batch (copied while the previous one computed)
        k[0] += 1
        it.prefetch(batches[k[0] % 2])       # next iteration's H2D on the side stream
        l = run_graph(cur) if graphs is not None else it.iteration(cur)
        loss_host.copy_(l.reshape(1), non_blocking=True)

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            dist.barrier()
        return ms

    for _ in range(max(args.warmup, 3)):
        resident_iter()
        e2e_iter()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("XGGM_BENCH_NO_CLOCKS") != "1":
        sampler.start()
    resident_iter()
    torch.cuda.synchronize()
    with sampler as clk:
        l0 = _lib.kernel_launches()
        ms_res = timed(resident_iter, args.steps)
        launches = _lib.kernel_launches() - l0
        ms_e2e = timed(e2e_iter, args.steps)
    clocks = clk.summary()
    # the library's share: the GGM part of step B alone on fixed encoder outputs
    visn, pooled = it.block_only(resident)
    adj = resident[5]

    def block():
        it.fg.zero_()
        x = pooled.detach().requires_grad_(True)
        f = visn.detach().requires_grad_(True)
        x_gen, loss_sm, _, _ = it.heads.node_step(x, f, adj, 1.0, it.A)
        torch.autograd.backward([x_gen, loss_sm], [torch.ones_like(x_gen), torch.tensor(1.1, device=dev)])

    for _ in range(3):
        block()
    ms_block = timed(block, 10) / 10
    if rank != 0:
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)
        return
    eager = None
    if not args.no_eager:
        try:
            from oracle.ref_step import ReferenceIteration, reference_available
            if reference_available():
                import contextlib
                with contextlib.redirect_stdout(sys.stderr):     # the reference prints its layer counts
                    ref = ReferenceIteration(dev, autocast=autocast)
                for _ in range(2):
                    ref.iteration(batches[0])
                torch.cuda.synchronize()
                n = max(2, min(args.steps, 5))
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for i in range(n):
                    ref.iteration(batches[i % 2])
                e.record()
                torch.cuda.synchronize()
                ms_ref = s.elapsed_time(e) / n
                eager = {"value": B / (ms_ref * 1e-3), "unit": "samples/s", "ms_per_iteration": ms_ref, "kind": "reference",
                         "n_gpus": 1, "sample": f"the unmodified reference modules (LXMERT, heads, GCNGenerator, BertAdam) "
                                                f"from oracle/_ref on rank 0's GPU, eager PyTorch, B={B}, {n} iterations"}
        except Exception as ex:   # the baseline must never take the measured arm down
            eager = {"unavailable": repr(ex)[:200]}
    ms_it = ms_res / args.steps
    n_par = sum(p.numel() for p in it.fg.params)
    line = {
        "metric": ITER_METRIC, "value": B * world / (ms_it * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_it, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if autocast else "f32", "data": "synthetic",
        "config": {"workload": f"cfg{3 if autocast else 2} full VQA-CP v2 training iteration (step A plain VQA + step B GGM "
                               f"node branch, delta={args.delta}): stock-PyTorch LXMERT 9/5/5 ({'bf16 autocast' if autocast else 'true fp32'}) "
                               f"+ xggm_b200 graph block / answer head / BCE / clip / BertAdam, B={B}/GPU, 36 objects x 2048-d, "
                               f"20 tokens, A={NUM_ANS}",
                   "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "parameters": n_par, "allreduce_bytes_per_iteration": 2 * 4 * n_par if world > 1 else 0,
                   "collective": ("none (one GPU)" if world == 1 else
                                  ("fused reduce-scatter + joint clip + BertAdam + parameter all-gather over NVLink peer memory"
                                   if it.fused_dp else "NCCL all-reduce (AVG)")),
                   "l2": "inputs and activations far exceed the 126 MB L2 (no flush needed)",
                   "launch": "one CUDA-graph replay per iteration (xggm_b200.GraphedStep, one graph per GGM branch)"
                             if graphs is not None else "eager launches"},
        "e2e": {"value": B * world / (ms_e2e / args.steps * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                "note": "one H2D of the batch per iteration (the reference copies it in both steps), overlapped with "
                        "the previous iteration"},
        "gpu_launches": launches, "clocks": clocks,
        "graph_block": {"ms": ms_block, "share_of_iteration": ms_block / ms_it,
                        "note": "the library's GGM part of step B alone (node_step fwd+bwd) on fixed encoder outputs"},
        "eager_gpu_baseline": eager,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


# ---------------------------------------------------------------------------- VisualFeatEncoder leg
def run_visn(args):
    """The 2048-d region-feature projection that produces the block's node features
    ((LN(W_f feats) + LN(W_b boxes)) / 2 + dropout, src/lxrt/modeling.py:530-556) on the library kernels: forward and
    forward+backward at B graphs x 36 objects.  Roofline: the K = 2048 projection is tensor-bound in fp32-parity mode
    (3 passes) and close to the HBM line in bf16 mode; both figures are reported."""
    import xggm_b200 as X
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, H, F = args.batch, HID, 2048
    X.set_precision(args.precision)
    torch.manual_seed(9595)
    enc = X.VisualFeatEncoder(hidden_size=H, hidden_dropout_prob=0.1, feat_dim=F, pos_dim=4).to(dev).train()
    from xggm_b200.ddp import FlatGrads
    import xggm_b200.functional as XF
    fg = FlatGrads(enc.parameters())
    XF.cache_weight_planes([p for p in enc.parameters() if p.dim() == 2])
    g = torch.Generator().manual_seed(1)
    feats_h = torch.relu(torch.randn(B, 36, F, generator=g)).pin_memory()
    xy = torch.rand(B, 36, 2, 2, generator=g).sort(dim=-1)[0]
    boxes_h = torch.stack([xy[..., 0, 0], xy[..., 1, 0], xy[..., 0, 1], xy[..., 1, 1]], dim=-1).contiguous().pin_memory()
    feats, boxes = feats_h.to(dev), boxes_h.to(dev)
    cot = torch.randn(B, 36, H, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def fwd():
        with torch.no_grad():
            return enc((feats, boxes))

    def fwd_bwd():
        fg.zero_()
        out = enc((feats, boxes))
        torch.autograd.backward([out], [cot])
        return out

    def e2e():
        fg.zero_()
        out = enc((feats_h.to(dev, non_blocking=True), boxes_h.to(dev, non_blocking=True)))
        torch.autograd.backward([out], [cot])
        return out

    def timed(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        torch.cuda.synchronize()
        for s_, e_ in evs:
            flush.fill_(1.0)
            s_.record()
            fn()
            e_.record()
        torch.cuda.synchronize()
        return sum(s_.elapsed_time(e_) for s_, e_ in evs) / k

    from xggm_b200 import _lib
    for fn in (fwd, fwd_bwd, e2e):
        for _ in range(max(args.warmup, 3)):
            fn()
    l0 = _lib.kernel_launches()
    ms_f = timed(fwd, args.steps)
    l1 = _lib.kernel_launches()
    ms_fb = timed(fwd_bwd, args.steps)
    l2 = _lib.kernel_launches()
    ms_e2e = timed(e2e, args.steps)
    peaks = load_peaks()
    M = B * 36
    flops_f = 2.0 * M * F * H + 2.0 * M * 4 * H
    bytes_f = 4.0 * (M * F + M * 4 + M * H + F * H)                 # read feats, boxes, weights; write out
    bytes_fb = bytes_f + 4.0 * (M * H + M * F + 2 * F * H)          # + read gout, re-read feats (wgrad), weight-grad RMW
    passes = 1 if args.precision == "bf16" else 3
    line = {
        "metric": "xggm_visual_feat_encoder_samples_per_sec", "value": B / (ms_fb * 1e-3), "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_fb, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"SURVEY 8 f-1 VisualFeatEncoder fwd+bwd (weight gradients; the inputs need none), B={B} x 36 objects x "
                               f"2048-d features + 4-d boxes -> 768, dropout 0.1, {args.precision} engine",
                   "l2": "flushed between timed steps", "launch": "eager launches"},
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(feats_h.numel() * 4 + boxes_h.numel() * 4), "d2h_bytes_per_step": 0},
        "gpu_launches": l2 - l1, "forward": {"ms": ms_f, "launches_per_call": (l1 - l0) // args.steps},
        "roofline": {"bound": "tensor" if passes == 3 else "hbm", "kernel": "visn_fc projection [B*36,2048]x[2048,768] + fused row tail",
                     "forward_tflops_algorithmic": flops_f / (ms_f * 1e-3) / 1e12,
                     "forward_tflops_executed": passes * flops_f / (ms_f * 1e-3) / 1e12,
                     "forward_frac_of_bf16_peak_executed": passes * flops_f / (ms_f * 1e-3) / 1e12 / peaks["bf16_tflops"],
                     "forward_gbs_algorithmic": bytes_f / (ms_f * 1e-3) / 1e9,
                     "forward_frac_of_hbm_peak": bytes_f / (ms_f * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "fwd_bwd_gbs_algorithmic": bytes_fb / (ms_fb * 1e-3) / 1e9,
                     "algorithmic_bytes_per_sample_forward": bytes_f / B, "peak_source": peaks["source"]},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- BASELINE configs[4]: evaluation
def run_eval(args):
    """predict() / evaluate() of the reference (src/vqa/vqacpv2.py:315-344): LXMERT forward + logit_fc + argmax.  The
    generator is NOT on this path: the graph block contributes zero work, the line exists so that every
    BASELINE config has a measurement.  Stock-PyTorch LXMERT (tools/lxmert_torch.py) + the library's VisualFeatEncoder and
    answer head, no gradients; B questions per GPU (default 1024), data parallel without any collective."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import iteration as IT
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch if args.batch != B_PER_GPU else 1024
    autocast = args.precision == "bf16"
    torch.backends.cuda.matmul.allow_tf32 = False
    it = IT.XGGMIteration(dev, B, autocast=autocast)
    it.lxmert.eval(); it.answer.eval(); it.heads.eval()
    host = IT.synthetic_batch(9596 + rank, B)
    res = [t.to(dev) for t in host]
    pred_host = torch.zeros(B, dtype=torch.long).pin_memory()

    def step(batch):
        with torch.no_grad():
            _, pooled = it._encode(batch[0], batch[1], batch[2], batch[3])
            return it.answer(pooled).max(1)[1]                      # vqacpv2.py:331-332

    def e2e():
        b = [t.to(dev, non_blocking=True) for t in host[:4]]
        pred_host.copy_(step(b), non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        step(res); e2e()

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(n):
            fn()
        e_.record()
        torch.cuda.synchronize()
        ms = s_.elapsed_time(e_)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / n

    ms = timed(lambda: step(res), args.steps)
    ms_e2e = timed(e2e, args.steps)
    if rank == 0:
        print(json.dumps({
            "metric": "xggm_eval_samples_per_sec", "value": B * world / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if autocast else "f32", "data": "synthetic",
            "config": {"workload": f"cfg5 evaluation: stock-PyTorch LXMERT forward + logit_fc + argmax, B={B}/GPU; the graph "
                                   "generator is not on the evaluation path (src/vqa/vqacpv2.py:315-344)",
                       "global_batch": B * world, "parallelism": f"dp{world} (no collective)"},
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[:4]), "d2h_bytes_per_step": B * 8}}), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xggm_b200", choices=["xggm_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"],
                    help="projection engine (default fp32 = BASELINE cfg 2; bf16 = cfg 3 arithmetic)")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="graphs per GPU (default 256, the BASELINE config)")
    ap.add_argument("--gnn", default=GNN, choices=["GCN", "GIN"])
    ap.add_argument("--branch", default="node", choices=["node", "relation", "mixed"],
                    help="GGM branch of the step: node generation (the --delta 0 recipe of script/vqacpv2.sh, default) or "
                         "relation generation (taken with probability delta/10; GQA-OOD uses delta 5)")
    ap.add_argument("--nodes", type=int, default=N_NODES, help="nodes per graph (36 = obj36; 64/100 = BASELINE cfg 4 sweep)")
    ap.add_argument("--quick", action="store_true", help="skip the CPU / eager-GPU baselines and the bf16 leg (sweeps)")
    ap.add_argument("--workload", default="block", choices=["block", "iteration", "visn", "eval"],
                    help="block: the graph block's training step (default, the headline); iteration: the full trainer "
                         "iteration with a stock-PyTorch LXMERT around the block (BASELINE configs[1]/[2])")
    ap.add_argument("--delta", type=int, default=0, help="--workload iteration: GGM branch threshold out of 10 (0 = VQA-CP recipe)")
    ap.add_argument("--no-eager", action="store_true", help="--workload iteration: skip the eager reference baseline")
    args = ap.parse_args()
    if args.workload == "iteration" and args.impl != "reference":
        run_iteration(args)
        return
    if args.workload == "visn" and args.impl != "reference":
        run_visn(args)
        return
    if args.workload == "eval" and args.impl != "reference":
        run_eval(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
