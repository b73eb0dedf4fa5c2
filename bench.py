"""bench.py -- the driver's benchmark contract for the RaCFormer sampling hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one sample of synthetic input at the racformer_r50_nuimg_704x256_f8
shapes (see bench_workloads.py for the workloads). One process per GPU; the path shards by sample with no
data-path collective, so N GPUs run N independent samples per step ("scaling": "weak").

Prints ONE JSON line on rank 0 (keys: see the task contract): value = whole-job samples/s with inputs resident in
HBM; e2e = the same through the public API with pinned-host inputs copied in and the result read back every step;
roofline = the step's dominant kernel (timed by event nodes inside the timed region's CUDA graph) against the measured
peak that bounds it; roofline_sampling_in_step = MSMV / MSDA forward inside the same step; roofline_ops = config 1, the
four sampling kernels forward + backward at the op shapes with a DRAM-level fraction beside the algorithmic one;
train / full_inference = configs 4 / 3 as secondary legs; kernels = per-family launch table of the step;
cpu_baseline = the reference's PyTorch grid_sample path (oracle port) on the host cores.

`--impl reference` times only that CPU path (rank 0 only) on the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference_arm(args, rank):
    """The reference's own CPU implementation of the path (PyTorch grid_sample; oracle port) on the host cores."""
    if rank != 0:
        return None
    import bench_workloads as workloads
    wl = workloads.build(args.workload, device="cpu", seed=0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(args.warmup):
        wl.reference_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.reference_step()
    dt = time.perf_counter() - t0
    value = args.steps * wl.reference_samples_per_step / dt
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": value, "unit": wl.unit, "cores": cores, "kind": "port",
                         "sample": wl.reference_sample_description},
        "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


class QuietStdout:
    """Route everything written to fd 1 (e.g. NCCL's version banner) to stderr until the JSON line is printed, so
    that stdout carries exactly one line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def main():
    with QuietStdout():
        line = _main()
    if line is not None:
        print(json.dumps(line), flush=True)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("RACF_BENCH_WORKLOAD", "decoder_forward_f8"))
    ap.add_argument("--cpu-baseline-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default=os.environ.get("RACF_BENCH_LEGS", "ops,train,train_3cam,full_inference"),
                    help="secondary measurements added to the JSON line (comma list; '' = none): ops = config 1 op "
                         "microbench (N=1 only), train = config 4 decoder training step, train_3cam = config 5, "
                         "full_inference = config 3")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import torch.distributed as dist
    try:   # make sure the in-tree library matches the sources (no-op when it is current; needs nvcc otherwise)
        from racformer_b200 import build as _lib_build
        if env_int("LOCAL_RANK", 0) == 0:
            _lib_build.build()
    except Exception as exc:  # the prebuilt .so travels with the snapshot; loading fails loudly below if it is absent
        print(f"bench.py: library rebuild skipped ({exc})", file=sys.stderr)
    import bench_workloads as workloads
    from racformer_b200 import parallel
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the ops)"
    rank, world, dev = parallel.init_distributed("nccl")

    def barrier():
        parallel.barrier(dev)

    wl = workloads.build(args.workload, device=dev, seed=rank)
    hbm_peak, peak_src = load_peaks()

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        wl.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    wl.reset_kernel_timers()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        wl.step()
    stop.record()
    barrier()
    total_ms = parallel.max_over_ranks(start.elapsed_time(stop), dev)   # device time, max over ranks
    value = world * args.steps * wl.samples_per_step / (total_ms * 1e-3)
    # per-kernel times: the dominant kernel from event nodes inside the timed region's own graph (they now hold its last
    # replay); every other kernel family from an instrumented copy of the graph replayed right after (workloads that do
    # not run as a graph: an eager pass with events around each launch)
    ms_per_step = total_ms / args.steps
    if getattr(wl, "use_graph", False):
        wl.probe_kernels()
    else:
        for _ in range(min(args.steps, 20)):
            wl.step(time_kernels=True)
    barrier()
    roof = wl.roofline(hbm_peak, peak_src, step_ms=ms_per_step)
    kernels = wl.kernel_report(hbm_peak)
    roof_sampling = wl.sampling_rooflines(hbm_peak)
    launches = wl.launches_per_step * args.steps

    # ---- end to end: pinned host inputs -> H2D -> ops -> D2H result, every step --------------------------------
    wl.prepare_host_inputs()
    flush = getattr(wl, "e2e_flush", lambda: None)
    for _ in range(3):
        wl.e2e_step()
    flush()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s2.record()
    for _ in range(e2e_steps):
        wl.e2e_step()
    flush()          # the last sample's result has been read back before the clock stops
    e2.record()
    barrier()
    ms2 = parallel.max_over_ranks(s2.elapsed_time(e2), dev)
    e2e_value = world * e2e_steps * wl.samples_per_step / (ms2 * 1e-3)
    clocks = sampler.stop()   # sampled across both timed regions (device-resident and end-to-end)

    # ---- CPU baseline (rank 0, N=1 only): the reference's PyTorch path on a bounded sample ----------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and "train" not in args.workload:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_wl = workloads.build(args.workload, device="cpu", seed=0)
        cpu_wl.reference_step()
        t0 = time.perf_counter()
        for _ in range(args.cpu_baseline_steps):
            cpu_wl.reference_step()
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": args.cpu_baseline_steps * cpu_wl.reference_samples_per_step / dt, "unit": wl.unit,
                        "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_baseline_steps} steps of: {cpu_wl.reference_sample_description}"}

    wl_metric, wl_unit, wl_config = wl.metric, wl.unit, wl.config()
    h2d_bytes, d2h_bytes = wl.h2d_bytes_per_step, wl.d2h_bytes_per_step

    # ---- secondary legs (never part of `value`): each one reports {"error": ...} instead of failing the run ------
    legs = {}
    want = [x for x in args.legs.split(",") if x]
    if args.workload != "decoder_forward_f8":
        want = []
    del wl
    torch.cuda.empty_cache()
    for leg in want:
        try:
            if leg == "ops":
                if world == 1:
                    legs["roofline_ops"] = workloads.OpMicrobench(dev).run(hbm_peak)
            elif leg in workloads.LEGS:
                legs[leg] = workloads.LEGS[leg](dev, rank, world, parallel)
        except Exception as exc:  # noqa: BLE001
            import traceback
            legs[leg if leg != "ops" else "roofline_ops"] = {"error": f"{type(exc).__name__}: {exc}",
                                                            "trace": traceback.format_exc()[-600:]}
        barrier()
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": wl_metric, "value": value, "unit": wl_unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl_config,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": wl_unit, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "h2d_gbs_total": e2e_value * h2d_bytes / 1e9, "h2d_gbs_per_gpu": e2e_value * h2d_bytes / 1e9 / world,
                    "note": "bounded by the host->device link: compare h2d_gbs with the measured concurrent pinned-copy "
                            "ceiling of the box (profiles/r02_h2d_ceiling_*.json, tools/h2d_ceiling.py); the full_inference "
                            "leg is the same path fed with its real input (95 MB of images / radar maps per sample)"},
            "gpu_launches": launches,
            "roofline": roof,
            "roofline_sampling_in_step": roof_sampling,
            "cpu_baseline": cpu_baseline,
            "kernels": kernels,
        }
        line.update(legs)
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
