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
    "stoch256": dict(model="stoch", size=256, output_nc=3, kind="edges2shoes", batch=10,
                     desc="StochCycleGAN 256x256 edges2shoes-shaped (the reference's 3-block generators), batch %d per GPU, "
                          "bf16, one train_instance step (G_A_B, G_B_A, D_A, D_B fwd+bwd, clip, Adam)"),
}


def _oracle_for(wl, device="cpu"):
    """(oracle model, batch maker) of a workload: the CPU / cuDNN restatement of the same step"""
    from oracle import nets as onets, step as ostep
    opt = ostep.default_opt(output_nc=wl["output_nc"])
    state = onets.init_model_state(seed=1234, output_nc=wl["output_nc"])
    om = (ostep.OracleModel if wl["model"] == "aug" else ostep.OracleStochModel)(opt, state, device=device)
    mk = lambda n, seed=4321: ostep.synthetic_batch(n, size=wl["size"], seed=seed, output_nc=wl["output_nc"], kind=wl["kind"])
    return om, mk


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


def _ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the residual-stack conv, from the committed
    ncu --set full summary; None when the profile is not in the tree."""
    for fn, key in (("r1b_ncu_summary.json", "pconv2_kernel_res_conv"), ("r1_ncu_summary.json", "igemm_kernel_res_conv")):
        try:
            k = json.load(open(os.path.join(ROOT, "profiles", fn)))[key]
            return k["dram_bytes_read"] + k["dram_bytes_write"]
        except Exception:
            continue
    return None


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
    om, mk = _oracle_for(wl)
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
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d train_instance steps at batch %d, torch %s CPU fp32, %d threads"
                                       % (steps, sample_n, torch.__version__, cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _desc(wl, args):
    n = args.batch or wl["batch"]
    return wl["desc"] % n if "%d" in wl["desc"] else wl["desc"].replace("batch %d" % BATCH, "batch %d" % n)


def cpu_baseline(wl):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 8 if wl["size"] == 64 else 2
    om, mk = _oracle_for(wl)
    a, b, z = mk(n)
    om.train_instance(a, b, z)
    steps, t0 = 0, time.perf_counter()
    while steps < 3 or (time.perf_counter() - t0 < 10.0 and steps < 20):
        om.train_instance(a, b, z)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d train_instance steps at batch %d%s, torch %s CPU fp32, %d threads"
                      % (steps, n, " (config 1)" if wl["model"] == "aug" else " at %dx%d" % (wl["size"], wl["size"]),
                         torch.__version__, cores)}


def torch_gpu_baseline(n, wl):
    """The reference's PyTorch path on this B200 (oracle port = same torch ops / cuDNN kernels), as the
    denominator of north_star's >=15x target: best of fp32(TF32 conv) and bf16 autocast."""
    res = {}
    a, b, z = [t.cuda() for t in _oracle_for(wl)[1](n)]
    for name in ("tf32", "bf16_autocast"):
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        om = _oracle_for(wl, "cuda")[0]

        def one():
            if name == "tf32":
                om.train_instance(a, b, z)
            else:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    om.train_instance(a, b, z)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 8
        for _ in range(k):
            one()
        torch.cuda.synchronize()
        res[name] = n * k / (time.perf_counter() - t0)
    res["unit"] = UNIT
    res["note"] = "oracle port of the reference step on cuda (cuDNN/ATen), batch %d, includes its 23 host syncs" % n
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
    opt = argparse.Namespace(**vars(ostep.default_opt(output_nc=wl["output_nc"])), expr_dir="/tmp", niter_decay=25)
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
    pending = None
    for batch in staged:
        # every step's packed loss vector is read back (D2H) inside the timed region; the read of step k is resolved
        # after step k+1 has been issued, so the host wait does not leave the GPU idle between steps
        nxt = m.train_instance(*batch, use_graph=graph_ok, report="defer")[0]
        if pending is not None:
            last = pending.get()
        pending = nxt
    last = pending.get()
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
                       "l2": "no flush needed: each step streams > 4 GB of activations, far above the 126 MB L2"},
            "e2e": e2e, "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "conv_roofline_frac_of_step": value / world * FLOP_PER_IMAGE / ((_peaks() or {}).get("bf16_tflops_sustained", 1418.6) * 1e12)}

    # ---- roofline of the dominant (tensor-core) kernels: instrumented eager pass.  EVERY rank runs the two steps
    # (they contain the gradient / batch-norm collectives); rank 0 reports its own timings.
    snap = m._snapshot()
    ops.PROFILE = ops.KernelProfile()
    m.lanes.enabled = False           # one stream: every kernel is timed alone, not against a concurrent branch
    for _ in range(2):
        m._step_device(*dev)
    torch.cuda.synchronize()
    m.lanes.enabled = True
    summ = ops.PROFILE.summary()
    ops.PROFILE = None
    m._restore(snap)
    say("instrumented pass done")
    if rank == 0:
        peaks = _peaks()
        peak = (peaks or {}).get("bf16_tflops_sustained", 1590.0 * 0.88)
        if args.precision == "tf32":
            peak = peak / 2.0
        tot_ms = sum(v["ms"] for v in summ.values())
        tot_fl = sum(v["flops"] for v in summ.values())
        if wl["model"] != "aug":     # algorithmic conv FLOPs of this workload = what the step's launches add up to
            line["conv_roofline_frac_of_step"] = (tot_fl / 2) / (ms / K * 1e-3) / (peak * 1e12)
        dom = max(summ, key=lambda k: summ[k]["ms"])
        ach = summ[dom]["flops"] / (summ[dom]["ms"] * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": dom + " (conv fwd/dgrad: igemm_kernel + pconv_kernel + pconv2_kernel)" if dom == "igemm_kernel" else dom,
                            "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                            "frac": ach / peak, "traffic": _ncu_traffic(),
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.59 PF x 0.88",
                            "per_kernel": {k: {"tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12, "ms_per_step": v["ms"] / 2,
                                               "launches_per_step": v["launches"] // 2} for k, v in summ.items()},
                            "tensor_kernels_share_of_step": (tot_ms / 2) / (ms / K),
                            "all_tensor_kernels_tflops": tot_fl / (tot_ms * 1e-3) / 1e12}
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
