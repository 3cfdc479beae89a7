#!/usr/bin/env python
"""bench.py -- train images/sec of one AugmentedCycleGAN.train_instance step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port) on host cores

Workload (config.workload): BASELINE.json configs[1] -- Augmented CycleGAN 64x64 edges2shoes-shaped, batch 80
PER GPU (weak scaling), bf16 activations / fp32 master weights, synthetic data, random-init weights.
A "step" = one full train_instance: 15 network forwards, D and G/E backward, 6 clips, 4 Adam steps.

JSON line keys: see the task contract; in short
  value   : global images / s, CUDA-event timed over exactly K CUDA-graph replays, inputs resident in HBM
  e2e     : same through the public API with HOST (pinned) inputs: H2D of real_A/real_B/prior_z and the D2H of
            the packed loss vector inside the timed region, every step
  roofline: tensor-core kernels (conv fwd/dgrad = igemm_kernel + pconv_kernel, weight gradient = wgrad_kernel),
            algorithmic FLOPs / CUDA-event time per launch, measured live in an instrumented pass on the launching
            stream, vs MEASURED_PEAKS.json (sustained bf16); `traffic` = DRAM bytes per launch of the dominant kernel's
            heaviest layer from the committed ncu --set full capture (profiles/), null if absent
  cpu_baseline: oracle port (restatement of the reference's train_instance, torch CPU fp32) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH = 80
METRIC = "train_images_per_sec"
UNIT = "img/s"
WORKLOAD = ("Augmented CycleGAN 64x64 edges2shoes-shaped, batch %d per GPU, bf16, one train_instance step "
            "(G_A_B, G_B_A, E_B, D_A, D_B, D_z_B fwd+bwd, clip, Adam)" % BATCH)
FLOP_PER_IMAGE = 36.003e9          # useful conv+linear FLOPs per image per step at 64x64
# --workload: the default is the BASELINE.json metric's configuration (configs[1]); the others time the parity-test
# configurations 3 and 4 through the same harness (StochCycleGAN is the reference-native step above 64x64)
WORKLOADS = {
    "aug64": dict(model="aug", size=64, output_nc=3, kind="edges2shoes", batch=BATCH, desc=WORKLOAD),
    "stoch128": dict(model="stoch", size=128, output_nc=1, kind="climate", batch=40,
                     desc="StochCycleGAN 128x128 Livneh-style climate fields (3 -> 1 channels), batch %d per GPU, bf16, "
                          "one train_instance step (G_A_B, G_B_A, D_A, D_B fwd+bwd, clip, Adam)"),
    "aug128": dict(model="aug", size=128, output_nc=1, kind="climate", batch=40, enc_grid=128,
                   desc="Augmented CycleGAN 128x128 Livneh-style climate fields (3 -> 1 channels) with the N3 encoder "
                        "extension (one extra stride-2 stage; the reference's E_B cannot run above 64x64), batch %d per GPU, "
                        "bf16, one train_instance step (G_A_B, G_B_A, E_B, D_A, D_B, D_z_B fwd+bwd, clip, Adam)"),
    "stoch256": dict(model="stoch", size=256, output_nc=3, kind="edges2shoes", batch=10,
                     desc="StochCycleGAN 256x256 edges2shoes-shaped (the reference's 3-block generators), batch %d per GPU, "
                          "bf16, one train_instance step (G_A_B, G_B_A, D_A, D_B fwd+bwd, clip, Adam)"),
}


def _oracle_for(wl, device="cpu"):
    """(oracle model, batch maker) of a workload: the CPU / cuDNN restatement of the same step"""
    from oracle import nets as onets, step as ostep
    opt = ostep.default_opt(output_nc=wl["output_nc"])
    state = onets.init_model_state(seed=1234, output_nc=wl["output_nc"], img_size=wl.get("enc_grid", 64))
    om = (ostep.OracleModel if wl["model"] == "aug" else ostep.OracleStochModel)(opt, state, device=device)
    mk = lambda n, seed=4321: ostep.synthetic_batch(n, size=wl["size"], seed=seed, output_nc=wl["output_nc"], kind=wl["kind"])
    return om, mk


def _reference_for(wl, device="cpu"):
    """(model, kind): the UNMODIFIED reference model.py (staged under baseline/_ref/ by oracle/stage_ref.py, or
    /root/reference in the build container; `.data[0]` -> `.item()` reporting shim only) -> kind "reference"; the oracle
    port when the sources are not available -> kind "port"."""
    import copy
    from oracle import live_reference as lr, nets as onets, step as ostep
    if not lr.available() or wl.get("enc_grid", 64) != 64:      # the reference itself has no encoder above 64x64
        return _oracle_for(wl, device)[0], "port"
    opt = ostep.default_opt(output_nc=wl["output_nc"])
    opt.gpu_ids = [torch.cuda.current_device()] if device != "cpu" else []
    opt.expr_dir = "/tmp"
    state = onets.init_model_state(seed=1234, output_nc=wl["output_nc"])
    return lr.build_reference_model(copy.deepcopy(opt), state, stoch=wl["model"] != "aug"), "reference"


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.rows)}
        sm = []
        reasons = set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        return out


def _ncu_traffic(workload, kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` summaries (profiles/*ncu_summary.json: {capture: {kernel, workload, dram_bytes_read, ...}}); None
    when no capture of THIS workload and kernel is in the tree."""
    import glob
    best = None
    for fn in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_summary.json"))):
        try:
            for key, k in json.load(open(fn)).items():
                name = str(k.get("kernel", key)).replace("void ", "").split("<")[0].split("(")[0].split("::")[-1].strip()
                if k.get("workload", "aug64") == workload and name and name in kernel:
                    best = k["dram_bytes_read"] + k["dram_bytes_write"]
        except Exception:
            continue
    return best


def kernel_profile(m, dev, _lib, ops, replays=3):
    """per-kernel device time of ONE serial step (see the call site); returns {"serial_us", "kernels": {name: {n, us,
    flops, bytes}}, "ops": {op: {n, us}}} (per step)"""
    from torch.profiler import ProfilerActivity, profile
    snap = m._snapshot()
    prev = _lib.set_option("pdl", 0)
    m.lanes.enabled = False
    try:
        for _ in range(2):
            m._step_device(*dev)
        torch.cuda.synchronize()
        ops.PROFILE = ops.OpLog()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m._step_device(*dev)
        log = list(ops.PROFILE.records)
        ops.PROFILE = None
        g.replay()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(replays):
                g.replay()
            torch.cuda.synchronize()
        del g
    finally:
        ops.PROFILE = None
        _lib.set_option("pdl", prev)
        m.lanes.enabled = True
        m._restore(snap)
    ev = []
    for k in prof.profiler.kineto_results.events():
        if "CUDA" in str(k.device_type()) and k.duration_ns() > 0:
            ev.append((k.start_ns(), k.duration_ns(), k.name()))
    ev.sort()
    span_us = (max(a + d for a, d, _ in ev) - ev[0][0]) / 1e3 / replays
    own = [(d, n) for a, d, n in ev if "dtg::" in n]
    per_step = sum(r[1] for r in log)
    if per_step == 0 or len(own) != per_step * replays:
        raise RuntimeError("kernel records (%d) do not match the op log (%d launches x %d replays)" % (len(own), per_step, replays))
    kernels, opsum = {}, {}
    for r in range(replays):
        i = r * per_step
        for name, nl, fl, by in log:
            recs = own[i:i + nl]
            i += nl
            us = sum(d for d, _ in recs) / 1e3
            o = opsum.setdefault(name, {"n": 0, "us": 0.0})
            o["n"] += 1
            o["us"] += us
            # the op's algorithmic work belongs to its heaviest kernel (conv: the tcgen05 kernel; wgrad: wgrad_kernel,
            # its split-K reduction is listed separately with no FLOPs of its own)
            heavy = max(range(len(recs)), key=lambda j: recs[j][0]) if recs else -1
            for j, (d, n) in enumerate(recs):
                kn = n.split("(")[0].replace("void ", "").replace("dtg::", "")
                kk = kernels.setdefault(kn, {"n": 0, "us": 0.0, "flops": 0.0, "bytes": 0.0})
                kk["n"] += 1
                kk["us"] += d / 1e3
                if j == heavy:
                    kk["flops"] += fl
                    kk["bytes"] += by
    for d in list(kernels.values()) + list(opsum.values()):
        for f in d:
            d[f] = d[f] / replays
        d["n"] = int(round(d["n"]))
    return {"serial_us": span_us, "kernels": kernels, "ops": opsum}


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port; the reference is pure Python over
    PyTorch, nothing to compile), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_n = 8 if wl["size"] == 64 else 2
    mk = _oracle_for(wl)[1]
    om, kind = _reference_for(wl)
    a, b, z = mk(sample_n)
    for _ in range(max(1, min(args.warmup, 2))):
        om.train_instance(a, b, z)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        om.train_instance(a, b, z)
    dt = time.perf_counter() - t0
    v = sample_n * steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": _desc(wl, args), "sample": "batch %d of the batch-%d workload per step" % (sample_n, args.batch or wl["batch"])},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d train_instance steps at batch %d, %s, torch %s CPU fp32, %d threads"
                                       % (steps, sample_n, "unmodified reference model.py" if kind == "reference" else "oracle port",
                                          torch.__version__, cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _desc(wl, args):
    n = args.batch or wl["batch"]
    return wl["desc"] % n if "%d" in wl["desc"] else wl["desc"].replace("batch %d" % BATCH, "batch %d" % n)


def cpu_baseline(wl):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 8 if wl["size"] == 64 else 2
    mk = _oracle_for(wl)[1]
    om, kind = _reference_for(wl)
    a, b, z = mk(n)
    om.train_instance(a, b, z)
    steps, t0 = 0, time.perf_counter()
    while steps < 3 or (time.perf_counter() - t0 < 10.0 and steps < 20):
        om.train_instance(a, b, z)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d train_instance steps at batch %d%s, %s, torch %s CPU fp32, %d threads"
                      % (steps, n, " (config 1)" if wl["model"] == "aug" else " at %dx%d" % (wl["size"], wl["size"]),
                         "unmodified reference model.py" if kind == "reference" else "oracle port", torch.__version__, cores)}


def torch_gpu_baseline(n, wl):
    """The reference's own PyTorch path on this B200 -- the denominator of north_star's >= 15x target (BASELINE.md 3,
    B2): the unmodified model.py (kind "reference"; the oracle port if the staged sources are missing) with fp32 / TF32
    convolutions and under torch.autocast(bf16), each in NCHW and channels_last; `best` is the fastest of the four."""
    res = {}
    a, b, z = [t.cuda() for t in _oracle_for(wl)[1](n)]
    kind = "port"
    for name in ("tf32", "bf16_autocast", "tf32_channels_last", "bf16_autocast_channels_last"):
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        try:
            om, kind = _reference_for(wl, "cuda")
            xa, xb = a, b
            if name.endswith("channels_last"):
                if kind != "reference":
                    continue
                for k in ("netG_A_B", "netG_B_A", "netE_B", "netD_A", "netD_B", "netD_z_B"):
                    if hasattr(om, k):
                        getattr(om, k).to(memory_format=torch.channels_last)
                xa, xb = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)

            def one():
                if name.startswith("tf32"):
                    om.train_instance(xa, xb, z)
                else:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        om.train_instance(xa, xb, z)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(3):
                    one()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                k = 8
                for _ in range(k):
                    one()
                torch.cuda.synchronize()
            res[name] = n * k / (time.perf_counter() - t0)
        except Exception as e:
            res[name + "_error"] = str(e).splitlines()[0][:160]
        finally:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
    vals = [v for k, v in res.items() if not k.endswith("_error")]
    res["best"] = max(vals) if vals else None
    res["unit"] = UNIT
    res["kind"] = kind
    res["note"] = ("%s step on cuda (cuDNN / ATen), batch %d, includes its >= 23 host syncs per step"
                   % ("unmodified reference model.py" if kind == "reference" else "oracle port of the reference", n))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--workload", default="aug64", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-baselines", action="store_true", help="skip cpu_baseline / torch_gpu_baseline legs")
    ap.add_argument("--no-sync-bn", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    verbose = os.environ.get("DTG_BENCH_VERBOSE") is not None

    def say(msg):
        if verbose:
            sys.stderr.write("[bench rank %d %.1fs] %s\n" % (rank, time.perf_counter() - T0, msg))
            sys.stderr.flush()

    T0 = time.perf_counter()
    if verbose:
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("DTG_BENCH_VERBOSE") or 60), exit=True)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    json_out = sys.stdout
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner to fd 1 when the communicator is created, so
        # fd 1 points at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    say("process group up")

    import dtg  # noqa: F401
    from dtg_b200 import _lib, engine, model as dmodel, ops, parallel
    from oracle import step as ostep          # synthetic data generator + default opt only

    engine.set_precision(args.precision)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    wl = WORKLOADS[args.workload]
    n = args.batch or wl["batch"]
    opt = argparse.Namespace(**vars(ostep.default_opt(output_nc=wl["output_nc"])), expr_dir="/tmp", niter_decay=25,
                             encoder_grid_size=wl.get("enc_grid", 64))
    torch.manual_seed(1234)
    m = (dmodel.AugmentedCycleGAN if wl["model"] == "aug" else dmodel.StochCycleGAN)(opt, testing=True)
    m.prepare()
    say("model built")
    if world > 1:
        m.dp = parallel.DataParallelPlan(sync_bn=not args.no_sync_bn)
        m.dp.broadcast_model(m)
    say("replicas broadcast")
    a, b, z = ostep.synthetic_batch(n, size=wl["size"], seed=4321 + rank, output_nc=wl["output_nc"], kind=wl["kind"])
    host = [t.pin_memory() for t in (a, b, z)]
    dev = [t.cuda() for t in host]
    use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- warm-up (captures the CUDA graph) -----------------------------------------------------------
    graph_ok = use_graph
    try:
        for _ in range(W):
            m.train_instance(*dev, use_graph=use_graph, report=False)
    except Exception as e:          # graph capture with NCCL can be refused: fall back to eager launches
        if not use_graph:
            raise
        sys.stderr.write("bench: CUDA-graph capture failed (%s); running eager\n" % str(e).splitlines()[0])
        graph_ok = False
        m._graphs = {}
        for _ in range(W):
            m.train_instance(*dev, use_graph=False, report=False)
    say("warm-up done (graph=%s)" % graph_ok)
    # kernels per step: count one eager pass (a graph replay launches exactly the captured set)
    snap = m._snapshot()
    lc0 = _lib.lib().dtg_launch_count()
    m._step_device(*dev)
    launches_per_step = _lib.lib().dtg_launch_count() - lc0
    torch.cuda.synchronize()
    m._restore(snap)

    # ---- timed region 1: device-resident inputs ------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        m.train_instance(*dev, use_graph=graph_ok, report=False)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * n * K / (ms * 1e-3)
    say("timed region 1 done: %.2f ms/step" % (ms / K))

    # ---- timed region 2: end to end through the public API, host inputs ------------------------------
    # dtg_b200.trainer.StagedBatches: each step's real_A / real_B travel pinned host -> device inside the timed region
    # (double-buffered on a copy stream), prior_z_B is drawn on the device, and the packed loss vector is read back
    # every step (report=True)
    from dtg_b200 import trainer

    def host_batches(k):
        for _ in range(k):
            yield {"A": a, "B": b}

    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    last = None
    for batch in trainer.StagedBatches(host_batches(2), opt.nlatent):      # allocate the staging buffers untimed
        m.train_instance(*batch, use_graph=graph_ok, report="defer")[0].get()
    barrier()
    e2.record()
    staged = trainer.StagedBatches(host_batches(K), opt.nlatent)
    import collections
    pending = collections.deque()
    for batch in staged:
        # every step's packed loss vector is read back (D2H) inside the timed region; the read of step k is resolved
        # after step k+2 has been issued (PendingReport allows four), so neither the host wait nor the host's issue
        # latency leaves the GPU idle between steps
        pending.append(m.train_instance(*batch, use_graph=graph_ok, report="defer")[0])
        if len(pending) > 2:
            last = pending.popleft().get()
    while pending:
        last = pending.popleft().get()
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    clocks = sampler.stop()
    e2e = {"value": world * n * K / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(staged.h2d_bytes // K), "d2h_bytes_per_step": int(m.scalars.numel() * 4)}
    if last is not None and not all(v == v for v in last[0].values()):
        raise SystemExit("bench: NaN in losses")

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": _desc(wl, args),
                       "global_batch": world * n, "parallelism": "dp%d" % world, "cuda_graph": bool(graph_ok),
                       "sync_bn": bool(world > 1 and not args.no_sync_bn),
                       "sync_bn_exchange": ("symmetric-memory one-shot" if (m.dp is not None and m.dp.sync_bn is not None and m.dp.sync_bn.symm)
                                            else ("nccl" if (m.dp is not None and m.dp.sync_bn is not None) else None)),
                       "l2": "no flush needed: each step streams > 4 GB of activations, far above the 126 MB L2"},
            "e2e": e2e, "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "conv_roofline_frac_of_step": value / world * FLOP_PER_IMAGE / ((_peaks() or {}).get("bf16_tflops_sustained", 1418.6) * 1e12)}

    # ---- per-kernel roofline: ONE serial step (single stream, programmatic dependent launch off) is captured in a CUDA
    # graph and replayed under CUPTI (torch.profiler); every dtg:: kernel record is attributed to the ops.* call whose
    # launch range covers it, which gives exact device durations per kernel (no host launch gaps, no time spent waiting
    # on a predecessor) next to the algorithmic FLOPs / bytes of the call.  Numbers taken under the profiler explain the
    # step; `value` / `e2e` above are never taken under it.  EVERY rank runs the pass (it contains the gradient and
    # batch-norm collectives); rank 0 reports its own records.
    kp = None
    try:
        if os.environ.get("DTG_BENCH_NO_KERNEL_PROFILE"):      # under ncu (launch lists): CUPTI belongs to the outer profiler
            raise RuntimeError("DTG_BENCH_NO_KERNEL_PROFILE is set")
        kp = kernel_profile(m, dev, _lib, ops)
    except Exception as e:      # CUPTI unavailable etc.: the line still carries value / e2e
        sys.stderr.write("bench: kernel profile unavailable (%s)\n" % str(e).splitlines()[0][:200])
    say("instrumented pass done")
    if rank == 0:
        peaks = _peaks()
        peak = (peaks or {}).get("bf16_tflops_sustained", 1590.0 * 0.88)
        hbm = (peaks or {}).get("hbm_gbs", 6650.0)
        if args.precision == "tf32":
            peak = peak / 2.0
        if kp is not None:
            tens = {k: v for k, v in kp["kernels"].items() if v["flops"] > 0}
            tot_us = sum(v["us"] for v in tens.values())
            tot_fl = sum(v["flops"] for v in tens.values())
            if wl["model"] != "aug":     # algorithmic conv FLOPs of this workload = what the step's launches add up to
                line["conv_roofline_frac_of_step"] = tot_fl / (ms / K * 1e-3) / (peak * 1e12)
            dom = max(tens, key=lambda k: tens[k]["us"])
            ach = tens[dom]["flops"] / (tens[dom]["us"] * 1e-6) / 1e12
            per_kernel = {}
            for k, v in sorted(kp["kernels"].items(), key=lambda kv: -kv[1]["us"]):
                d = {"launches_per_step": v["n"], "us_per_step": round(v["us"], 1)}
                if v["flops"] > 0:
                    d["tflops"] = v["flops"] / (v["us"] * 1e-6) / 1e12
                    d["frac_of_peak"] = d["tflops"] / peak
                if v["bytes"] > 0:
                    d["gbs"] = v["bytes"] / (v["us"] * 1e-6) / 1e9
                    d["frac_of_hbm"] = d["gbs"] / hbm
                per_kernel[k] = d
            line["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                                "frac": ach / peak, "traffic": _ncu_traffic(args.workload, dom),
                                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.59 PF x 0.88",
                                "hbm_peak_gbs": hbm,
                                "how": "serial CUDA graph of one step, PDL off, CUPTI kernel durations (3 replays, mean)",
                                "per_kernel": per_kernel,
                                "serial_step_us": round(kp["serial_us"], 1),
                                "tensor_kernels_us_per_step": round(tot_us, 1),
                                "tensor_kernels_share_of_serial_step": tot_us / kp["serial_us"],
                                "tensor_kernels_time_over_step_time": tot_us * 1e-3 / (ms / K),
                                "all_tensor_kernels_tflops": tot_fl / (tot_us * 1e-6) / 1e12}
        if world == 1 and not args.no_baselines:
            try:
                line["torch_gpu_baseline"] = torch_gpu_baseline(n, wl)
            except Exception as e:
                line["torch_gpu_baseline"] = {"error": str(e).splitlines()[0][:200]}
            line["cpu_baseline"] = cpu_baseline(wl)
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        # release the captured graph (it holds NCCL kernels) before tearing the communicator down; the teardown of a
        # communicator that was used under graph capture can block for minutes, so leave without it
        m._graphs = {}
        torch.cuda.synchronize()
        dist.barrier()
        say("leaving")
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
