#!/usr/bin/env python
"""bench.py -- voice-samples/s and realtime factor of the generator back end.

Workload (BASELINE.json configs[2], "C3"): 4096 concurrent voices of 3-operator
PM/FM chains with envelope ramps at 96 kHz stereo.  A step is ONE generator
call (sauGenerator_run) of 24576 frames = 256 ms of audio for all voices, i.e.
4096 x 24576 voice-samples (the reference player's own call size,
saugns.c:471,526).  At N GPUs every rank renders an independent 4096-voice
script (weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W]        GPU arm
    python bench.py --impl reference ...                         reference CPU arm
    python bench.py --workload c4|c5 ...                         the other BASELINE configs
                                                                 (reported in DESIGN.md, not the headline)
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

SRATE = 96000
FRAMES = 24576          # 256 ms at 96 kHz, saugns.c:471
VOICES = 4096
SECS = 60
METRIC = "voice-samples/sec"
WORKLOAD = ("C3: 4096 concurrent voices x 3-operator PM/FM chains (alternating PM chain / "
            "range-FM+PM) with xpe/lin amp ramps, 60 s script at 96 kHz stereo; "
            "step = one 24576-frame generator call")


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


# ---------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an in-process NVML poller
    (5 ms period; nvidia_ml_py), or `nvidia-smi -lms` where NVML cannot be loaded.
    start() is called before the warm-up steps so that the poller is up when the timed
    region begins; mark() opens / closes the region and only samples inside it count (all
    samples when the region was shorter than a polling period)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.samples = []            # (time, sm_mhz, sm_max_mhz, set(reasons))
        self.marks = []
        self.stop_flag = False
        self.thread = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:                  # NVML counts physical devices
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    idx = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self.stop_flag:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.perf_counter(), sm, mx,
                                             {k for k, b in bits.items() if r & b}))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                self.samples.append((time.perf_counter(), float(p[1]), float(p[2]),
                                     {nm for k, nm in enumerate(names)
                                      if p[5 + k].lower().startswith("active")}))
            except ValueError:
                continue

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"],
                    "samples": 0}
        time.sleep(0.02)
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sel = self.samples
        if len(self.marks) >= 2:
            inside = [x for x in self.samples if self.marks[0] <= x[0] <= self.marks[-1]]
            if inside:
                sel = inside
            else:                    # region shorter than a period: the nearest samples
                mid = 0.5 * (self.marks[0] + self.marks[-1])
                sel = sorted(self.samples, key=lambda x: abs(x[0] - mid))[:3]
        sm = sorted(x[1] for x in sel)
        reasons = set()
        for x in sel:
            reasons |= x[3]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max((x[2] for x in sel), default=None), "reasons": sorted(reasons),
                "samples": len(sm), "source": self.how}


# ---------------------------------------------------------------------------
# reference CPU arm / cpu_baseline
# ---------------------------------------------------------------------------
def _ref_worker(conn, text):
    """One host process = one single-threaded reference generator (the reference has no
    threading of its own) over a slice of the workload's voices, kept alive across steps."""
    from oracle import pyref
    prg = pyref.Program(text)
    gen = None
    while True:
        cmd = conn.recv()
        if cmd[0] == "stop":
            break
        t0 = time.perf_counter()
        if gen is None:
            gen = pyref.RefGenerator(prg, SRATE)          # sau_create_Generator, timed
        frames = 0
        for _ in range(cmd[1]):
            more, _, n = gen.run(FRAMES)                  # sauGenerator_run, 24576 frames
            frames += n
        conn.send((frames, prg.vo_count, time.perf_counter() - t0))
    conn.close()


class ReferencePool:
    """The unmodified reference generator (oracle/_ref/libsauref.so) on `procs` host
    processes, each rendering a disjoint slice of the C3 voices call by call."""

    def __init__(self, voices, procs, seed=1):
        import multiprocessing as mp
        from saugns_b200 import workloads
        from oracle import pyref
        if not pyref.available():
            raise RuntimeError("oracle/_ref/libsauref.so missing")
        full = workloads.synth_c3(voices, SECS, seed=seed, fm="mix").splitlines()
        head, body = full[0], full[1:]
        per = (len(body) + procs - 1) // procs
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for p in range(procs):
            part = body[p * per:(p + 1) * per]
            if not part:
                continue
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_ref_worker, args=(b, "\n".join([head] + part) + "\n"),
                             daemon=True)
            pr.start()
            self.conns.append(a)
            self.procs.append(pr)
        self.used = len(self.procs)

    def step(self, calls=1):
        """All workers render `calls` more calls concurrently -> (voice_samples, wall s)."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("step", calls))
        res = [c.recv() for c in self.conns]
        wall = time.perf_counter() - t0
        return sum(f * v for f, v, _ in res), wall

    def close(self):
        for c in self.conns:
            c.send(("stop",))
        for p in self.procs:
            p.join(timeout=10)


def run_reference_arm(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    if args.warmup + args.steps > SECS * SRATE // FRAMES:
        raise SystemExit("steps+warmup exceed the workload's calls")
    pool = ReferencePool(VOICES, procs)
    # one step = the same 24576-frame call over all 4096 voices, on all host cores
    for _ in range(args.warmup):
        pool.step()
    t_tot, vs_tot = 0.0, 0
    for _ in range(args.steps):
        vs, wall = pool.step()
        t_tot += wall
        vs_tot += vs
    used = pool.used
    pool.close()
    value = vs_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voice-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_tot / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "voices": VOICES, "frames_per_step": FRAMES,
                   "srate": SRATE, "op_samples_per_step": 3 * VOICES * FRAMES,
                   "l2": "host arm: the reference's block buffers live in the CPU caches",
                   "parallelism": f"one 4096-voice script, voices split over {used} host processes "
                                  "(the reference is single-threaded)"},
        "realtime_factor": (FRAMES * args.steps / SRATE) / t_tot,
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": used,
                         "kind": "reference",
                         "sample": f"unmodified reference generator (oracle/_ref, -O3 -ffast-math), "
                                   f"{VOICES} voices split over {used} single-threaded processes, "
                                   f"{FRAMES} frames per step, wall clock around each step; "
                                   "script parsing excluded"},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import saugns_b200
    from saugns_b200 import workloads

    rank, local_rank, world = env_rank()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available() or saugns_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the B200 back end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()
    K, W = args.steps, args.warmup
    if W + K > SECS * SRATE // FRAMES - 1:
        raise SystemExit("steps+warmup exceed the workload's calls")

    def build_program():
        return workloads.build_c3(VOICES, SECS, seed=1 + rank, fm="mix")

    prg = build_program()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: PCM stays in HBM ----
    sched = int(os.environ.get("SAUGEN_BENCH_SCHED", "0"))   # developer knob; 0 = auto
    g = saugns_b200.Generator(prg, SRATE, device=local_rank, stream=stream.cuda_stream,
                              max_call_len=FRAMES, sched=sched)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        g.run_device(FRAMES)
    g.set_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    c0 = g.counters()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(K):
            more, _, n = g.run_device(FRAMES)
            assert more and n == FRAMES
        ev1.record(stream)
    barrier()
    sampler.mark()
    clocks = sampler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    c1 = g.counters()
    render_ms, mix_ms = g.kernel_ms()
    launches = (c1[0] - c0[0]) + (c1[1] - c0[1])
    g.close()

    # ---- end to end through the public call with HOST buffers ----
    # timed: sau_create_Generator (program flattening + H2D upload of events, op-data,
    # bytecode, tables) + W+K x sauGenerator_run with a host PCM buffer (D2H every call)
    # + destroy; the W warm-up calls are part of the same render, so they are timed and
    # counted too (a renderer cannot skip the start of its script).
    import ctypes
    prg_bytes = 0
    try:
        from saugns_b200 import program as P
        pp = P.Program.from_address(prg.ptr)
        prg_bytes = (ctypes.sizeof(P.Program) + pp.ev_count * ctypes.sizeof(P.Event) +
                     pp.op_count * (ctypes.sizeof(P.OpData) + 3 * ctypes.sizeof(P.Line)))
    except Exception:
        pass
    barrier()
    t0 = time.perf_counter()
    g2 = saugns_b200.Generator(prg, SRATE, device=local_rank, stream=stream.cuda_stream,
                               max_call_len=FRAMES)
    for _ in range(W + K):
        more, pcm, n = g2.run(FRAMES)
    g2.close()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_steps = W + K
    assert int(abs(pcm.astype("int32")).max()) > 0, "silent output"
    barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    vs_step = VOICES * FRAMES
    value = world * vs_step * K / (ms / 1000.0)
    e2e = world * vs_step * e2e_steps / e2e_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    # algorithmic bytes (8 B per voice-sample = one f32 carrier store by the
    # render kernel + one f32 load by the mix kernel; pans are constant in C3, so no r rows):
    # 4 B per voice-sample for EACH of the two kernels' launches
    alg_bytes = 4.0 * vs_step
    mix_s = (mix_ms / K) / 1000.0
    mix_alg = 4.0 * vs_step + 4.0 * FRAMES
    rk_s = (render_ms / K) / 1000.0
    achieved = alg_bytes / rk_s / 1e9
    prof = {}
    try:
        import glob
        latest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_render_kernel.json")))[-1]
        prof = json.load(open(latest))
        prof["file"] = os.path.relpath(latest, ROOT)
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "voice-samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "voices": VOICES, "frames_per_step": FRAMES,
                   "srate": SRATE, "op_samples_per_step": 3 * vs_step,
                   "l2": "per-step voice rows 403 MB > 126 MB L2 (working set larger than L2)",
                   "parallelism": f"independent 4096-voice scripts x{world}"},
        "realtime_factor": (FRAMES * K / SRATE) / (ms / 1000.0),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "voice-samples/s",
                "h2d_bytes_per_step": 40 + 12 + 6 * 12 + prg_bytes // max(e2e_steps, 1),
                "d2h_bytes_per_step": FRAMES * 2 * 2 + 8,
                "realtime_factor": (FRAMES * e2e_steps / SRATE) / e2e_s,
                "timed": f"create (program flatten + upload, {prg_bytes} B) + {e2e_steps} calls "
                         "with host PCM buffers + destroy, wall clock"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": prof.get("dram_bytes_per_launch"),
                     "kernel": "render_kernel", "kernel_ms_per_launch": render_ms / K,
                     "mix_kernel_ms_per_launch": mix_ms / K, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "mix_kernel": {"bound": "hbm", "achieved": mix_alg / mix_s / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": mix_alg / mix_s / 1e9 / peak,
                                    "algorithmic_bytes_per_launch": mix_alg},
                     "note": "path is issue-/FP64-pipe-bound, not HBM-bound (SURVEY.md 8d); "
                             "issue-slot figures from ncu in profiles/",
                     "issue_slot_frac_ncu": prof.get("issue_slot_frac"),
                     # the same launch's instruction count (ncu) over the issue slots of THIS run's
                     # measured kernel time: 4 schedulers x 148 SMs x SM clock (the ncu capture itself
                     # is a cold, serialised launch and runs ~40 % longer)
                     "issue_slot_frac_live": (prof["warp_insts"] / (4 * 148 * (clocks.get("sm_mhz") or 1965.0)
                                                                  * 1e6 * rk_s)
                                              if prof.get("warp_insts") else None),
                     "fp64_pipe_frac_ncu": prof.get("fp64_pipe_frac"),
                     "xu_pipe_frac_ncu": prof.get("xu_pipe_frac"),
                     "ncu_profile": prof.get("file")},
    }
    if world == 1 and not args.no_cpu:
        try:
            cores = min(os.cpu_count() or 1, 64)
            pool = ReferencePool(VOICES, cores)
            pool.step(1)                    # create + first call (warm-up)
            ncalls = 40                     # ~10 s of audio for all 4096 voices
            vs, wall = pool.step(ncalls)
            used = pool.used
            pool.close()
            line["cpu_baseline"] = {
                "value": vs / wall, "unit": "voice-samples/s", "cores": used, "kind": "reference",
                "sample": f"unmodified reference generator (oracle/_ref), same {VOICES}-voice "
                          f"script, calls 2..{ncalls + 1} of {FRAMES} frames, voices split over "
                          f"{used} single-threaded processes; {wall:.2f} s wall"}
        except Exception as e:   # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "voice-samples/s", "cores": 0,
                                    "kind": "reference", "sample": f"unavailable: {e}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_gpu_config(args):
    """The other BASELINE configs on one GPU (numbers for DESIGN.md section 8; the
    headline line stays C3).  Scripts go through the reference's own script front end
    on the host, exactly as in the drop-in (north star): oracle/_ref/libsauref.so is used
    here for PARSING only; every sample is rendered by the CUDA back end."""
    import numpy as np
    import torch
    import saugns_b200
    from saugns_b200 import workloads, batch
    from saugns_b200 import program as P
    from oracle import pyref, pyport
    import ctypes as C
    rank, local_rank, world = env_rank()
    torch.cuda.set_device(local_rank)
    t = pyport.ref_tables()
    tabs = saugns_b200.WaveTables.from_buffer_copy(bytes(t))
    tabs._keep = t
    K, W = args.steps, args.warmup
    if args.workload == "c3-sharded":
        # ONE 4096-voice script, its voices spread over the ranks (strong scaling): each rank
        # renders its voices' float mix planes, one NCCL sum-reduce per call, root converts
        import torch.distributed as dist
        from saugns_b200 import multigpu
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        prg = workloads.build_c3(VOICES, SECS, seed=1, fm="mix")
        vg = multigpu.VoiceShardedGenerator(prg, SRATE, device=local_rank, max_call_len=FRAMES)
        for _ in range(W):
            vg.run(FRAMES)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            more, pcm, n = vg.run(FRAMES)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        vg.close()
        if rank == 0:
            print(json.dumps({"metric": METRIC, "workload": "C3, ONE 4096-voice script voice-sharded over "
                              f"{world} GPU(s): one ncclReduce of the float L/R planes per call, PCM on the "
                              "root's host buffer", "value": VOICES * FRAMES * K / wall,
                              "unit": "voice-samples/s", "n_gpus": world, "steps": K,
                              "ms_per_step": 1000 * wall / K, "scaling": "strong",
                              "realtime_factor": (FRAMES * K / SRATE) / wall}))
        if world > 1:
            dist.destroy_process_group()
        return 0
    if args.workload == "c4":
        nv = 1024
        prg = pyref.Program(workloads.synth_c4(nv, SECS))
        g = saugns_b200.Generator(prg, SRATE, tables=tabs, device=local_rank, max_call_len=FRAMES)
        for _ in range(W):
            g.run_device(FRAMES)
        g.set_timing(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            g.run(FRAMES)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        rk, mk = g.kernel_ms()
        g.close()
        # latency model: every self-PM operator is one serial chain over the call
        sm_mhz = 1965.0
        line = {"metric": METRIC, "workload": "C4: 1024 voices, self-PM carriers (W and R) with "
                "range-AM / ring-mod, 96 kHz; step = one 24576-frame call, host PCM buffers",
                "value": nv * FRAMES * K / wall, "unit": "voice-samples/s", "steps": K,
                "ms_per_step": 1000 * wall / K, "render_kernel_ms": rk / K, "mix_kernel_ms": mk / K,
                "realtime_factor": (FRAMES * K / SRATE) / wall,
                "cycles_per_feedback_iteration_at_max_clock": (rk / K) * 1e-3 * sm_mhz * 1e6 / FRAMES}
        print(json.dumps(line))
        return 0
    # c5: independent scripts, this GPU's share of 10 000 (default 10000/8 = 1250)
    n = args.scripts
    dist = None
    if world > 1:              # script sharding: every rank its own scripts, no data-path collective
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    texts = [workloads.synth_c5_script(rank * n + i) for i in range(n)]
    t0 = time.perf_counter()
    prgs = [pyref.Program(x) for x in texts]
    parse_s = time.perf_counter() - t0
    vs = 0
    for p in prgs:
        d = P.dump(p.ptr)
        for ev in d["events"]:
            for od in ev["ops"]:
                if od["id"] == ev["carr_op_id"]:
                    vs += od["time"][0] * SRATE // 1000
    batch.render_batch(prgs[:16], srate=SRATE, device=local_rank, tables=tabs, group_size=16)  # warm-up
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    got = {}

    def sink(i, pcm):          # what a file writer would get: every script's PCM, once
        got[i] = (pcm.shape[0], int(pcm[::997].astype(np.int64).sum()))

    batch.render_batch(prgs, srate=SRATE, device=local_rank, tables=tabs,
                       group_size=args.group, threads=args.threads, sink=sink,
                       call_len=args.call_frames, pinned=args.pinned, depth=args.depth)
    wall = time.perf_counter() - t0
    assert len(got) == n
    frames = sum(v[0] for v in got.values())
    if dist is not None:       # whole job: all ranks' scripts over the slowest rank's time
        t = torch.tensor([wall, -float(vs), -float(frames)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0].item())
        tot = torch.tensor([float(vs), float(frames)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        vs, frames = int(tot[0].item()), int(tot[1].item())
        dist.destroy_process_group()
        if rank != 0:
            return 0
        n = n * world
    line = {"metric": METRIC, "n_gpus": world, "scaling": "weak",
            "workload": f"C5: {n} independent mixed scripts (4-16 voices, W/N/R, "
            f"1-10 s) on {world} GPU(s), batched saugen_run_many, every script's PCM delivered to a host "
            f"sink (arrays recycled), {args.threads} driver thread(s) x 2 alternating "
            f"live sets, {args.call_frames}-frame calls",
            "value": vs / wall, "unit": "voice-samples/s", "scripts": n, "group": args.group,
            "wall_s": wall, "scripts_per_s": n / wall, "audio_s": frames / SRATE,
            "realtime_factor": (frames / SRATE) / wall, "parse_s_reference_front_end": parse_s}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 / C5 / voice-sharded legs (`extra`)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5", "c3-sharded"])
    ap.add_argument("--scripts", type=int, default=1250, help="c5: scripts on this GPU")
    ap.add_argument("--group", type=int, default=128, help="c5: generators in flight per driver thread")
    ap.add_argument("--threads", type=int, default=1, help="c5: driver threads")
    ap.add_argument("--pinned", action="store_true", help="c5: page-locked (recycled) PCM arrays, no staging copy")
    ap.add_argument("--depth", type=int, default=2, help="c5: alternating live sets per driver thread")
    ap.add_argument("--call-frames", type=int, default=4 * FRAMES,
                    help="c5: frames per generator call (results do not depend on it)")
    args = ap.parse_args()
    # the script is 60 s (BASELINE config 3) unless more calls than that are asked for:
    # both arms then render the same, longer script (every step is a full 24576-frame call)
    global SECS, WORKLOAD
    need = ((args.steps + args.warmup + 3) * FRAMES + SRATE - 1) // SRATE
    if need > SECS:
        WORKLOAD = WORKLOAD.replace("60 s script", f"{need} s script")
        SECS = need
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload != "c3":
        return run_gpu_config(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
