#!/usr/bin/env python3
"""bench.py -- AIS channels demodulated per second on N B200s (BASELINE.json metric).

A "step" is one pass of the demod hot path over one batch of synthetic IQ: `--channels`
independent 48 ksps channels per GPU, each `--seconds` long (default 1 s = 48 000 complex
samples).  Defaults are BASELINE.json's own configurations:
  1 GPU            : configs[2], 65 536 channels through the full chain
  N GPUs (torchrun): configs[3], 32 768 channels per GPU (262 144 over 8)
  --channels 4096 --workload corr_msk : configs[1] literally
workload "chain"    : freq sync -> AGC -> corr_est_cc -> msk_timing_recovery_cc -> bits
                      (every row of SURVEY.md section 8a; the default)
workload "corr_msk" : corr_est_cc -> msk_timing_recovery_cc -> bits only
Records (`--records`):
  independent (default): 64 different seeded records (own burst times, payloads, CFO, phase,
      fractional delay), channel c = record c % 64 rotated by its own random offset, so the
      lanes of a warp meet their bursts at unrelated times;
  coherent: ONE record rotated by 16*c samples (round 1's input; kept as a second line under
      "coherent_records": the data-dependent kernels see neighbouring channels in lock step).
value = channel-seconds of IQ demodulated per second, whole job, inputs resident in HBM.
e2e   = the same metric through the public host-buffer call (ais_demod.work: pinned host
        IQ in, host bits out, H2D/D2H inside the timed region); "e2e_sc16" the same with the
        IQ delivered as interleaved int16 (the SDR wire format), half the bytes over PCIe.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the reference's CPU implementation of the path on the host cores:
the gr-ais block sources compiled unmodified (oracle/_ref, GNU Radio kernels underneath
restated -- GNU Radio 3.8 + VOLK are not installed) inside the oracle's schedule, OpenMP over
all host threads, on a bounded sample of the same workload (extrapolated linearly: channels
are independent).
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
    ap.add_argument("--channels", type=int, default=0,
                    help="channels per GPU (0 = BASELINE configs[2]'s 65536 on one GPU, configs[3]'s 32768 per GPU on several)")
    ap.add_argument("--records", default="independent", choices=["independent", "coherent"])
    ap.add_argument("--no-coherent", action="store_true", help="skip the second (coherent-records) device line")
    ap.add_argument("--no-sc16", action="store_true", help="skip the int16-IQ end-to-end line")
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


POOL = 64  # distinct seeded records behind the "independent" batch
# warp-instructions k_corr_fft executes per input sample, by template length (ncu, see bench note)
CORR_WIPS = {120: 3.885}


def record_plan(args, n, channels, first_channel=0):
    """(pool of records [P][n] complex64, record index per channel, rotation per channel).
    Channel c of the job is np.roll(pool[pid[c]], off[c]); the plan is a pure function of the
    global channel number, so every rank and the CPU arm build the same channels."""
    from gr_ais_b200 import synth
    nb = max(1, int(4 * args.seconds))
    cg = np.arange(first_channel, first_channel + channels, dtype=np.int64)
    if args.records == "coherent":
        pool = [synth.make_record(0, n=n, nbursts=nb, snr_db=args.snr_db)[0]]
        return np.stack(pool), np.zeros(channels, np.int64), (16 * cg) % n
    pool = [synth.make_record(p, n=n, nbursts=nb, snr_db=args.snr_db, random_impairments=True)[0]
            for p in range(POOL)]
    # per-channel rotation: splitmix64 of the global channel number (counter-based, no stream state)
    z = (cg.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(31)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(29)
    return np.stack(pool), cg % POOL, (z % np.uint64(n)).astype(np.int64)


def host_rows(plan, lo, hi):
    pool, pid, off = plan
    return np.stack([np.roll(pool[pid[c]], int(off[c])) for c in range(lo, hi)])


def fill_device_batch(x_dev, plan, dev):
    """x_dev[c] = roll(pool[pid[c]], off[c]) built on the GPU (a 65536 x 48000 batch is 25 GB)."""
    import torch
    pool, pid, off = plan
    n = pool.shape[1]
    pool_d = torch.view_as_complex(torch.from_numpy(pool.view(np.float32).reshape(len(pool), n, 2)).to(dev))
    xc = torch.view_as_complex(x_dev)
    ar = torch.arange(n, device=dev)
    pid_d = torch.from_numpy(pid).to(dev)
    off_d = torch.from_numpy(off).to(dev)
    step = 1024
    for c0 in range(0, xc.shape[0], step):
        c1 = min(c0 + step, xc.shape[0])
        idx = (ar[None, :] - off_d[c0:c1, None]) % n  # np.roll(base, k)[i] = base[(i - k) % n]
        xc[c0:c1] = pool_d[pid_d[c0:c1]].gather(1, idx)
    torch.cuda.synchronize(dev)


def stages_for(workload, B):
    return (B.STAGE_FREQSYNC | B.STAGE_AGC) if workload == "chain" else 0


def default_channels(args, world):
    if args.channels:
        return args.channels
    return 65536 if world == 1 else 32768


def workload_name(args, n, channels, world):
    st = ("freqest->mix->agc->corr_est->msk_timing->quad_demod->slicer->diff->invert"
          if args.workload == "chain" else "corr_est->msk_timing->quad_demod->slicer->diff->invert")
    if channels == 65536 and world == 1 and args.workload == "chain":
        cfgname = "BASELINE configs[2]: "
    elif channels == 32768 and world > 1 and args.workload == "chain":
        cfgname = "BASELINE configs[3] (%d of 8 GPUs, 32768 channels each): " % world
    elif channels == 4096 and world == 1 and args.workload == "corr_msk":
        cfgname = "BASELINE configs[1]: "
    else:
        cfgname = ""
    return "%s%d channels/GPU x %d samples (%.2f s @ 48 ksps), %s" % (cfgname, channels, n, n / FS, st)


def ref_blocks_or_none():
    """oracle/_ref (the reference's own block sources) when it was built, else None (oracle port)."""
    try:
        from oracle import ref as R
        if R.available():
            return R.blocks(), "reference+shim"
    except Exception:
        pass
    return None, "port"


def cpu_sample(args, threads, sample_channels, n, steps, warmup):
    """Time the reference's CPU path on `sample_channels` channels of the same batch."""
    from gr_ais_b200.ais_demod import preamble_template
    from oracle import oracle as O
    blocks, kind = ref_blocks_or_none()
    x = host_rows(record_plan(args, n, sample_channels), 0, sample_channels)
    tmpl = preamble_template(args.template)
    stages = (O.STAGE_FREQSYNC | O.STAGE_AGC) if args.workload == "chain" else 0
    cfg = O.chain_cfg(stages=stages)
    for _ in range(warmup):
        O.demod_chain_batch(x[:threads], tmpl, cfg, nthreads=threads, blocks=blocks)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.demod_chain_batch(x, tmpl, cfg, nthreads=threads, blocks=blocks)
    dt = (time.perf_counter() - t0) / steps
    return sample_channels * (n / FS) / dt, dt, kind


def cpu_sample_size(args, threads, n, seconds=12.0):
    """Channels that give about `seconds` of CPU work per pass (probed with 2 per thread)."""
    if args.cpu_channels:
        return args.cpu_channels
    probe = 2 * threads
    v, _, _ = cpu_sample(args, threads, probe, n, 1, 1)
    want = int(v * seconds / (n / FS))
    return max(probe, min(want // threads * threads, 16384))


CPU_KIND_NOTE = {
    "reference+shim": "the reference's own block sources (lib/*_impl.cc compiled unmodified, oracle/_ref) inside "
                      "the oracle's schedule; the GNU Radio / VOLK kernels under them and the stock blocks between "
                      "them are the oracle's restatements (GNU Radio 3.8 is not installed)",
    "port": "oracle port of the reference path (oracle/_ref was not built on this box)",
}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = int(round(args.seconds * FS))
    channels = default_channels(args, max(world, args.gpus))
    # a step = a bounded sample of the workload: ~6 s of CPU work, so K steps end within minutes
    sample = cpu_sample_size(args, threads, n, seconds=6.0)
    warm = 1 if args.warmup > 0 else 0
    value, dt, kind = cpu_sample(args, threads, sample, n, max(1, args.steps), warm)
    sample_txt = ("%d of the workload's %d channels x %d samples per step, OpenMP over %d host threads; "
                  "channels are independent, so channels/s extrapolates linearly to the full batch"
                  % (sample, channels, n, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "channels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, n, channels, max(world, args.gpus)),
                   "template_taps": {"north_star": 120, "intended": 140, "reference": 1120}[args.template],
                   "records": args.records, "cpu_sample": sample_txt,
                   "note": "CPU arm: " + CPU_KIND_NOTE[kind]},
        "cpu_baseline": {"value": value, "unit": "channels/s", "cores": threads, "kind": kind,
                         "sample": sample_txt},
        "e2e": {"value": value, "unit": "channels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def host_memory_available():
    try:
        with open("/proc/meminfo") as fh:
            for ln in fh:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return 0


def run_b200(args):
    import torch
    from gr_ais_b200 import binding as B
    from gr_ais_b200.ais_demod import ais_demod, preamble_template

    rank, local_rank, world = dist_env()
    # stdout carries the one JSON line: while the communicator comes up (NCCL prints its
    # "NCCL version ..." banner on file descriptor 1) it points at stderr
    saved_stdout = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    B.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from gr_ais_b200 import sharding
    numa_cpus = sharding.bind_to_gpu_numa_node(local_rank) if world > 1 else None

    # rank 0 owns the preamble template; the other ranks receive it over NCCL (the only collective)
    tmpl = sharding.broadcast_template(preamble_template(args.template) if rank == 0 else None,
                                       src=0, device=dev)
    if saved_stdout is not None:
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    C = default_channels(args, world)
    n = int(round(args.seconds * FS))
    d = ais_demod(channels=C, max_samples=n, template=tmpl, stages=stages_for(args.workload, B))
    mb = d.max_bits(n)
    x_dev = torch.empty((C, n, 2), dtype=torch.float32, device=dev)
    plan = record_plan(args, n, C, first_channel=rank * C)
    fill_device_batch(x_dev, plan, dev)
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

    def timed_region(sample_clocks):
        """W warm-up steps, then K steps between CUDA events on the launching stream."""
        for _ in range(max(args.warmup, 3)):
            step_timed()
        d.join(sp)
        torch.cuda.synchronize(dev)
        d.status()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
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
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        launches = B.launch_count() - launches0
        d.status()
        return ms, launches, sampler

    for _ in range(max(args.warmup, 1)):
        step_dev()
    torch.cuda.synchronize(dev)
    d.status()

    # ---- device-resident timed region (CUDA events on the launching stream) ----
    d.set_overlap(args.overlap)
    ms_step, launches, sampler = timed_region(True)
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
    e2e, e2e_sc16 = None, None
    if not args.no_e2e:
        in_bytes = C * n * 8
        # the whole batch in pinned host memory when the host has room for it (25 GB at 65536
        # channels), else the first `Ce` channels sent C / Ce times per step (same bytes over PCIe)
        avail = host_memory_available()
        Ce = C
        while Ce > 1024 and avail and Ce * n * 8 * 2.5 * max(world, 1) > avail:
            Ce //= 2
        reps = C // Ce
        de = d if Ce == C else ais_demod(channels=Ce, max_samples=n, template=tmpl,
                                          stages=stages_for(args.workload, B))
        torch.cuda.empty_cache()
        pin_in = B.PinnedArray((Ce, n), np.complex64)
        torch.from_numpy(pin_in.array.view(np.float32).reshape(Ce, n, 2)).copy_(x_dev[:Ce])
        pin_bits = B.PinnedArray((Ce, mb), np.uint8)
        pin_nbits = B.PinnedArray((Ce,), np.int32)
        pin_tags = B.PinnedArray((Ce, de.max_tags), B.TAG_DTYPE)
        pin_ntags = B.PinnedArray((Ce,), np.int32)

        def time_host(fn):
            for _ in range(max(1, min(args.warmup, 2))):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                for _ in range(reps):
                    fn()
            torch.cuda.synchronize(dev)
            return sharding.max_over_ranks(time.perf_counter() - t0, dev) / args.steps

        dt_step = time_host(lambda: de.work(pin_in.array, pin_bits.array, pin_nbits.array,
                                            pin_tags.array, pin_ntags.array))
        if not np.array_equal(pin_nbits.array, nbits_host[:Ce]):
            raise RuntimeError("host-buffer and device-resident runs disagree on symbol counts")
        d2h = int(C * mb + C * 4 + C * de.max_tags * B.TAG_DTYPE.itemsize + C * 4)
        host_note = ("pinned (b200ais_host_alloc)" +
                     ("" if Ce == C else ", %d-channel calls x %d per step (host memory)" % (Ce, reps)) +
                     (", rank bound to the GPU's %d local CPUs" % len(numa_cpus) if numa_cpus else ""))
        e2e = {"value": world * C * (n / FS) / dt_step, "unit": "channels/s",
               "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": d2h,
               "ms_per_step": dt_step * 1e3, "h2d_gbs_all_ranks": world * in_bytes / dt_step / 1e9,
               "input_format": "complex64 (gr_complex)", "host_memory": host_note}
        if not args.no_sc16 and hasattr(de, "work_sc16"):
            # the same channels as interleaved int16 IQ (what UHD / osmosdr put on the wire,
            # python/radio.py:151-203), converted on the device with the exact scale 2^-15
            pin_sc = B.PinnedArray((Ce, n, 2), np.int16)
            sc_host = torch.from_numpy(pin_sc.array)
            for c0 in range(0, Ce, 1024):  # in slices: the batch is 25 GB and torch makes temporaries
                c1 = min(c0 + 1024, Ce)
                sc_host[c0:c1].copy_((x_dev[c0:c1] * 8192.0).round().clamp(-32768, 32767).to(torch.int16))
            torch.cuda.empty_cache()
            dt16 = time_host(lambda: de.work_sc16(pin_sc.array, 1.0 / 8192.0, pin_bits.array, pin_nbits.array,
                                                  pin_tags.array, pin_ntags.array))
            e2e_sc16 = {"value": world * C * (n / FS) / dt16, "unit": "channels/s",
                        "h2d_bytes_per_step": int(C * n * 4), "d2h_bytes_per_step": d2h,
                        "ms_per_step": dt16 * 1e3, "h2d_gbs_all_ranks": world * C * n * 4 / dt16 / 1e9,
                        "input_format": "sc16 (interleaved int16 I/Q, scale 2^-13 applied on the device)",
                        "note": "quantised input: its bits are checked against the oracle on the same "
                                "quantised samples in tests/test_gpu_sc16.py, not against the fp32 run"}
        if de is not d:
            de.close()

    # ---- the other record layout (same step, same kernels) ----
    coherent = None
    if not args.no_coherent:
        other = argparse.Namespace(**vars(args))
        other.records = "coherent" if args.records == "independent" else "independent"
        fill_device_batch(x_dev, record_plan(other, n, C, first_channel=rank * C), dev)
        ms_o, _, _ = timed_region(False)
        coherent = {"records": other.records, "value": world * C * (n / FS) / (ms_o * 1e-3),
                    "unit": "channels/s", "ms_per_step": ms_o}

    # ---- rooflines: the correlator (north-star's kernel) and the longest kernel of the step ----
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    per = {k: v / max(calls, 1) for k, v in stage_ms.items()}
    corr_ms = per["corr"]
    taps = len(tmpl)
    fft = 2
    while fft < 2 * taps:
        fft *= 2                      # kernel::fft_filter_ccc: 2 * 2^ceil(log2 taps)
    ns = fft - taps + 1
    n1 = (n // 1024) * 1024 if args.workload == "chain" else n
    n2 = (n1 // ns) * ns              # corr_est processes whole filter blocks
    alg_bytes = 8.0 * C * n2          # 8 B per complex sample read
    achieved = alg_bytes / (corr_ms * 1e-3) / 1e9 if corr_ms > 0 else None
    lg = fft.bit_length() - 1
    # FP32 lane-operations per filter block (one FADD / FMUL / FFMA lane each; a packed FFMA2
    # is two): radix-2 butterflies of both transforms, the trivial twiddles 1 and -i as adds
    # only, the product with the taps spectrum, the tail add and |.|^2
    nontriv = (fft // 2) * lg - (fft - 1) - (fft // 2 - 1)
    lane_ops = 2 * (nontriv * 8 + ((fft - 1) + (fft // 2 - 1)) * 4) + fft * 4 + 2 * (taps - 1) + 3 * ns
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    fp32_lanes = 148 * 128 * sm_hz    # FP32 lane-operations per second of the chip
    fp32_floor_ms = lane_ops * (C * n2 / ns) / fp32_lanes * 1e3
    flop_per_block = 2 * (fft // 2) * lg * 10 + 6 * fft + 2 * (taps - 1) + 3 * ns
    roofline = {"bound": "hbm", "kernel": "k_corr_fft", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "ms_per_launch": corr_ms,
                "timing": "CUDA events around each k_corr_fft launch, K steps re-run with all kernels serialised on the launching stream, right after the timed region",
                "fp32_tflops": (flop_per_block * (C * n2 / ns) / (corr_ms * 1e-3) / 1e12) if corr_ms > 0 else None,
                "fp32_pipe_floor_ms": fp32_floor_ms,
                "fp32_pipe_frac": (fp32_floor_ms / corr_ms) if corr_ms > 0 else None,
                "hbm_frac_at_fp32_pipe_floor": alg_bytes / (fp32_floor_ms * 1e-3) / 1e9 / peak,
                # instruction-issue ceiling: the kernel executes WIPS warp-instructions per sample (a
                # property of its code, counted by ncu: smsp__inst_executed.sum / samples of the capture
                # in profiles/r02_ncu_corr.csv); 148 SMs x 4 schedulers issue one per cycle each
                "warp_instructions_per_sample": CORR_WIPS.get(taps),
                "issue_ceiling_frac": (148 * 4 * sm_hz / CORR_WIPS[taps] * 8.0 / 1e9 / peak) if taps in CORR_WIPS else None,
                "note": "GNU Radio's fft_filter (FFT overlap-add, fftsize %d, %d items per block) costs %d FP32 lane-operations "
                        "per 8-byte sample: at %.0f MHz the FP32 pipe alone holds the kernel to %.2f ms = the "
                        "hbm_frac_at_fp32_pipe_floor above, so the kernel is bound by the FP32 pipe / instruction issue, "
                        "not by HBM; `traffic` (dram bytes per launch) is not measurable in-run: the ncu capture is "
                        "profiles/r02_ncu_corr.csv" % (fft, ns, lane_ops // ns, sm_hz / 1e6, fp32_floor_ms)}
    longest = max(per, key=lambda k: per[k])
    alg_per_sample = {"sqfft_freqest": 8.0, "mix_agc": 16.0, "corr": 8.0, "msk": 9.6, "tail": 1.8,
                      "nco_phase": 0.0, "detect": 0.125}
    lb = alg_per_sample.get(longest, 8.0) * C * n1
    roofline_longest = {"bound": "hbm", "kernel": longest, "ms_per_launch": per[longest],
                        "algorithmic_bytes_per_launch": lb, "unit": "GB/s", "peak": peak,
                        "achieved": lb / (per[longest] * 1e-3) / 1e9 if per[longest] > 0 else None,
                        "frac": lb / (per[longest] * 1e-3) / 1e9 / peak if per[longest] > 0 else None,
                        "note": "algorithmic bytes per input sample: " + json.dumps(alg_per_sample)}

    line = {
        "metric": METRIC, "value": value, "unit": "channels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n, C, world), "template_taps": taps,
                   "records": ("%d seeded records with independent burst times / CFO / phase / delay, each channel its own random rotation"
                               % POOL) if args.records == "independent" else "one record rotated by 16*c samples",
                   "sharding": "%d channels on each of %d GPUs, no data-path collective; NCCL broadcast of the template" % (C, world),
                   "l2_policy": "inputs larger than L2: %.2f GB of IQ per GPU per step vs 126 MB L2" % (C * n * 8 / 1e9),
                   "snr_db": args.snr_db},
        "roofline": roofline, "roofline_longest": roofline_longest, "clocks": clocks,
        "gpu_launches": int(launches),
        "stage_ms_per_step": per,
        "serialized_ms_per_step": serial_ms_step, "overlap_groups": args.overlap,
        "submission": ("strictly ordered work_dev calls" if args.strict else
                       "enqueue_dev x K + join; all K results complete before the closing event.  Up to 53 248 "
                       "channels the timing loop of record k runs on a high-priority side stream under the front "
                       "half of record k+1; a larger batch fills the GPU by itself and stays in stream order"),
        "symbols_per_channel": int(nbits_host[0]),
    }
    if coherent:
        line["coherent_records" if coherent["records"] == "coherent" else "independent_records"] = coherent
    if e2e:
        line["e2e"] = e2e
    if e2e_sc16:
        line["e2e_sc16"] = e2e_sc16
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = cpu_sample_size(args, threads, n)
        v, dt, kind = cpu_sample(args, threads, sample, n, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": "channels/s", "cores": threads, "kind": kind,
                                "sample": "%d of %d channels x %d samples, one pass, OpenMP over %d host threads (%.1f s); %s"
                                          % (sample, C, n, threads, dt, CPU_KIND_NOTE[kind])}
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
