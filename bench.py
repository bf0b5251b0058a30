#!/usr/bin/env python3
"""bench.py -- AIS channels demodulated per second on N B200s (BASELINE.json metric).

A "step" is one pass of the demod hot path over one batch of synthetic IQ:
`--channels` independent 48 ksps channels per GPU (default 16384: between BASELINE.json
configs[1]'s 4096 and configs[3]'s 32768 per GPU; the two per-channel recurrences cost the
same ~6 ms for any batch up to ~19k channels, so small batches under-use the GPU;
`--channels 4096` runs configs[1]'s size), each `--seconds` long (default 1 s = 48 000
complex samples), through
  workload "chain"    : freq sync -> AGC -> corr_est_cc -> msk_timing_recovery_cc -> bits
                        (every row of SURVEY.md section 8a; the default)
  workload "corr_msk" : corr_est_cc -> msk_timing_recovery_cc -> bits only
                        (the literal stage list of configs[1])
value = channel-seconds of IQ demodulated per second, whole job, inputs resident in HBM.
e2e   = the same metric through the public host-buffer call (ais_demod.work: pinned host
        IQ in, host bits out, H2D/D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the CPU restatement of the reference path (oracle/, OpenMP over
all host cores) on a bounded sample of the same workload: the reference itself needs
GNU Radio 3.8 + VOLK and cannot be built here.
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

FS = 48000
METRIC = "AIS channels demodulated/sec (48 ksps IQ)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chain", choices=["chain", "corr_msk"])
    ap.add_argument("--channels", type=int, default=16384, help="channels per GPU")
    ap.add_argument("--seconds", type=float, default=1.0, help="record length per channel")
    ap.add_argument("--template", default="north_star", choices=["north_star", "intended", "reference"])
    ap.add_argument("--snr-db", type=float, default=20.0)
    ap.add_argument("--cpu-channels", type=int, default=0,
                    help="channels in the CPU sample (0 = about 12 s of CPU work, probed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", type=int, default=1,
                    help="channel groups forked over internal streams inside one call (1 = none)")
    ap.add_argument("--strict", action="store_true",
                    help="time strictly ordered work_dev calls instead of the pipelined enqueue_dev")
    return ap.parse_args()


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_host_batch(args, pinned_cls):
    """configs[1]-style fan-out: one seeded synthetic record (AWGN + 4 AIS bursts), each
    channel its own copy rotated by 16*c samples so channels are not phase-locked."""
    from gr_ais_b200 import synth
    n = int(round(args.seconds * FS))
    base, _ = synth.make_record(0, n=n, nbursts=max(1, int(4 * args.seconds)), snr_db=args.snr_db)
    pin = pinned_cls((args.channels, n), np.complex64)
    for c in range(args.channels):
        pin.array[c] = np.roll(base, 16 * c)
    return pin, n


def stages_for(workload, B):
    return (B.STAGE_FREQSYNC | B.STAGE_AGC) if workload == "chain" else 0


def workload_name(args, n):
    st = ("freqest->mix->agc->corr_est->msk_timing->quad_demod->slicer->diff->invert"
          if args.workload == "chain" else "corr_est->msk_timing->quad_demod->slicer->diff->invert")
    return "%d channels/GPU x %d samples (%.2f s @ 48 ksps), %s" % (args.channels, n, n / FS, st)


def cpu_sample(args, threads, sample_channels, n, steps, warmup):
    """Time the oracle (CPU restatement of the reference path) on `sample_channels` channels."""
    from gr_ais_b200 import synth
    from gr_ais_b200.ais_demod import preamble_template
    from oracle import oracle as O
    base, _ = synth.make_record(0, n=n, nbursts=max(1, int(4 * args.seconds)), snr_db=args.snr_db)
    x = np.stack([np.roll(base, 16 * c) for c in range(sample_channels)])
    tmpl = preamble_template(args.template)
    stages = (O.STAGE_FREQSYNC | O.STAGE_AGC) if args.workload == "chain" else 0
    cfg = O.chain_cfg(stages=stages)
    for _ in range(warmup):
        O.demod_chain_batch(x[:threads], tmpl, cfg, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.demod_chain_batch(x, tmpl, cfg, nthreads=threads)
    dt = (time.perf_counter() - t0) / steps
    return sample_channels * (n / FS) / dt, dt


def cpu_sample_size(args, threads, n, seconds=12.0):
    """Channels that give about `seconds` of CPU work per pass (probed with 2 per thread)."""
    if args.cpu_channels:
        return args.cpu_channels
    probe = 2 * threads
    v, _ = cpu_sample(args, threads, probe, n, 1, 1)
    want = int(v * seconds / (n / FS))
    return max(probe, min(want // threads * threads, 16384))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = int(round(args.seconds * FS))
    sample = cpu_sample_size(args, threads, n)
    warm = 1 if args.warmup > 0 else 0
    value, dt = cpu_sample(args, threads, sample, n, max(1, args.steps), warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "channels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, n), "template_taps": {"north_star": 120, "intended": 140, "reference": 1120}[args.template],
                   "note": "CPU arm: oracle port of the reference path (reference needs GNU Radio 3.8+VOLK, unbuildable here)"},
        "cpu_baseline": {"value": value, "unit": "channels/s", "cores": threads, "kind": "port",
                         "sample": "%d channels x %d samples per step, OpenMP over %d host threads" % (sample, n, threads)},
        "e2e": {"value": value, "unit": "channels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    from gr_ais_b200 import binding as B
    from gr_ais_b200.ais_demod import ais_demod, preamble_template

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    B.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from gr_ais_b200 import sharding
    numa_cpus = sharding.bind_to_gpu_numa_node(local_rank) if world > 1 else None

    # rank 0 owns the preamble template; the other ranks receive it over NCCL (the only collective)
    from gr_ais_b200 import sharding
    tmpl = sharding.broadcast_template(preamble_template(args.template) if rank == 0 else None,
                                       src=0, device=dev)

    pin_in, n = make_host_batch(args, B.PinnedArray)
    C = args.channels
    d = ais_demod(channels=C, max_samples=n, template=tmpl, stages=stages_for(args.workload, B))
    mb = d.max_bits(n)
    x_dev = torch.empty((C, n, 2), dtype=torch.float32, device=dev)
    x_dev.copy_(torch.from_numpy(pin_in.array.view(np.float32).reshape(C, n, 2)), non_blocking=False)
    bits_dev = torch.zeros((C, mb), dtype=torch.uint8, device=dev)
    nbits_dev = torch.zeros(C, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream

    def step_dev():
        d.work_dev(x_dev.data_ptr(), n, bits_dev.data_ptr(), mb, nbits_dev.data_ptr(), None, None, sp)

    # the timed region submits the K records with enqueue_dev (the timing loop of record k runs
    # on a high-priority side stream under the front half of record k+1) and joins before the
    # closing event, so all K results are complete inside the timed region
    def step_timed():
        if args.strict:
            step_dev()
        else:
            d.enqueue_dev(x_dev.data_ptr(), n, bits_dev.data_ptr(), mb, nbits_dev.data_ptr(), None, None, sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 1)):
        step_dev()
    torch.cuda.synchronize(dev)
    d.status()

    # ---- device-resident timed region (CUDA events on the launching stream) ----
    d.set_overlap(args.overlap)
    for _ in range(max(args.warmup, 3)):
        step_timed()
    d.join(sp)
    torch.cuda.synchronize(dev)
    d.status()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = B.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step_timed()
        d.join(sp)
        e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = B.launch_count() - launches0
    d.status()
    ms_step = sharding.max_over_ranks(ms_total, dev) / args.steps
    value = world * C * (n / FS) / (ms_step * 1e-3)
    nbits_host = nbits_dev.cpu().numpy()

    # ---- per-kernel durations: the same K steps again with every kernel serialised on the
    # launching stream and CUDA events around each launch (concurrent channel groups would
    # make a single kernel's duration meaningless) ----
    d.profile(True)
    d.stage_ms()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        p0.record(stream)
        for _ in range(args.steps):
            step_dev()
        p1.record(stream)
    torch.cuda.synchronize(dev)
    serial_ms_step = p0.elapsed_time(p1) / args.steps
    stage_ms, calls = d.stage_ms()
    d.profile(False)
    d.status()
    clocks = sampler.stop()

    # ---- end to end through the public host-buffer call ----
    e2e = None
    if not args.no_e2e:
        pin_bits = B.PinnedArray((C, mb), np.uint8)
        pin_nbits = B.PinnedArray((C,), np.int32)
        pin_tags = B.PinnedArray((C, d.max_tags), B.TAG_DTYPE)
        pin_ntags = B.PinnedArray((C,), np.int32)

        def step_host():
            d.work(pin_in.array, pin_bits.array, pin_nbits.array, pin_tags.array, pin_ntags.array)

        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        dt_step = sharding.max_over_ranks(dt, dev) / args.steps
        if not np.array_equal(pin_nbits.array, nbits_host):
            raise RuntimeError("host-buffer and device-resident runs disagree on symbol counts")
        e2e = {"value": world * C * (n / FS) / dt_step, "unit": "channels/s",
               "h2d_bytes_per_step": int(C * n * 8),
               "d2h_bytes_per_step": int(C * mb + C * 4 + C * d.max_tags * B.TAG_DTYPE.itemsize + C * 4),
               "ms_per_step": dt_step * 1e3, "host_memory": "pinned (b200ais_host_alloc)" +
               (", rank bound to the GPU's %d local CPUs" % len(numa_cpus) if numa_cpus else "")}

    # ---- roofline of the dominant kernel (corr_est correlator) ----
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    corr_ms = stage_ms["corr"] / max(calls, 1)
    taps = len(tmpl)
    fft = 2
    while fft < 2 * taps:
        fft *= 2                      # kernel::fft_filter_ccc: 2 * 2^ceil(log2 taps)
    ns = fft - taps + 1
    n1 = (n // 1024) * 1024 if args.workload == "chain" else n
    n2 = (n1 // ns) * ns              # corr_est processes whole filter blocks
    alg_bytes = 8.0 * C * n2          # 8 B per complex sample read
    achieved = alg_bytes / (corr_ms * 1e-3) / 1e9 if corr_ms > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "corr_traffic.json")) as fh:
            t = json.load(fh)
            traffic = t["dram_bytes_per_launch"] * (C * n2) / float(t["samples_per_launch"])
    except Exception:
        pass
    lg = fft.bit_length() - 1
    flop_per_block = 2 * (fft // 2) * lg * 10 + 6 * fft + 2 * (taps - 1) + 3 * ns
    roofline = {"bound": "hbm", "kernel": "k_corr_fft", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_moved_per_launch": 16.0 * C * n2 + C * n2 / 8.0,
                "ms_per_launch": corr_ms,
                "timing": "CUDA events around each k_corr_fft launch, K steps re-run with all kernels serialised on the launching stream, right after the timed region",
                "fp32_tflops": (flop_per_block * (C * n2 / ns) / (corr_ms * 1e-3) / 1e12) if corr_ms > 0 else None,
                "note": "GNU Radio's fft_filter (FFT overlap-add, fftsize %d, %d items per block): ~%d flop per 8-byte sample, "
                        "compute-bound; achieved counts the 8 B/sample read, bytes_moved adds the correlator stream "
                        "(8 B/sample) and the bitmask it writes" % (fft, ns, flop_per_block // ns)}

    line = {
        "metric": METRIC, "value": value, "unit": "channels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n), "template_taps": taps,
                   "sharding": "%d channels on each of %d GPUs, no data-path collective; NCCL broadcast of the template" % (C, world),
                   "l2_policy": "inputs larger than L2: %.2f GB of IQ per GPU per step vs 126 MB L2" % (C * n * 8 / 1e9),
                   "snr_db": args.snr_db},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "stage_ms_per_step": {k: v / max(calls, 1) for k, v in stage_ms.items()},
        "serialized_ms_per_step": serial_ms_step, "overlap_groups": args.overlap,
        "submission": ("strictly ordered work_dev calls" if args.strict else
                       "enqueue_dev x K + join: msk_timing + bit tail of record k on a high-priority side "
                       "stream under the front half of record k+1; all K results complete before the closing event"),
        "symbols_per_channel": int(nbits_host[0]),
    }
    if e2e:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = cpu_sample_size(args, threads, n)
        v, dt = cpu_sample(args, threads, sample, n, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": "channels/s", "cores": threads, "kind": "port",
                                "sample": "%d channels x %d samples, one pass, OpenMP over %d host threads (%.1f s)" % (sample, n, threads, dt)}
    if rank == 0:
        print(json.dumps(line))
    d.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
