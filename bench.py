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
# the workloads: script text for the reference's front end, the identical sauProgram for the
# product arm (saugns_b200.workloads; tests/test_program_builder.py pins the two field by field)
# ---------------------------------------------------------------------------
def c3_text(seed=1):
    from saugns_b200 import workloads
    return workloads.synth_c3(VOICES, SECS, seed=seed, fm="mix")


def c3_program(seed=1):
    from saugns_b200 import workloads
    return workloads.build_c3(VOICES, SECS, seed=seed, fm="mix")


def c3_config():
    """Identical in both arms (the driver compares the two lines' `config`)."""
    return {"workload": WORKLOAD, "voices": VOICES, "frames_per_step": FRAMES, "srate": SRATE,
            "op_samples_per_step": 3 * VOICES * FRAMES,
            "l2": "per-step voice rows 403 MB > 126 MB L2 (working set larger than L2)",
            "parallelism": "one independent 4096-voice script per GPU (per host-core group in the "
                           "reference arm: the reference is single-threaded, voices split over processes)"}


# ---------------------------------------------------------------------------
# reference CPU arm / cpu_baseline / parity (the only users of oracle/)
# ---------------------------------------------------------------------------
def _ref_worker(conn, text, frames):
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
        n_tot = 0
        for _ in range(cmd[1]):
            more, _, n = gen.run(frames)                  # sauGenerator_run
            n_tot += n
        conn.send((n_tot, prg.vo_count, time.perf_counter() - t0))
    conn.close()


class ReferencePool:
    """The unmodified reference generator (oracle/_ref/libsauref.so) on `procs` host
    processes, each rendering a disjoint slice of a many-voice script call by call."""

    def __init__(self, text, procs, frames=FRAMES):
        import multiprocessing as mp
        from oracle import pyref
        if not pyref.available():
            raise RuntimeError("oracle/_ref/libsauref.so missing")
        full = text.splitlines()
        head, body = full[0], full[1:]
        per = (len(body) + procs - 1) // procs
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for p in range(procs):
            part = body[p * per:(p + 1) * per]
            if not part:
                continue
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_ref_worker, args=(b, "\n".join([head] + part) + "\n", frames),
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


def _parity_worker(conn, text, ncalls, frames, table_file):
    """The unmodified reference generator on the WHOLE script (one process, in voice order:
    the mix's float summation order) for the first `ncalls` calls.  The wave tables are input
    data of the path: the GPU arm reads them from `table_file`, written from
    libsau on the build host; libm's last bits differ between host CPUs (tests/golden/
    make_golden.py), so when this host's libsau builds other bits the reference's table
    arrays are overwritten with the file's before it renders -- same input on both sides."""
    import ctypes as C
    import numpy as np
    from oracle import pyref
    L = pyref.lib()
    prg = pyref.Program(text)            # sau_global_init_Wave has run by now
    t = pyref.piluts()
    blob = open(table_file, "rb").read()
    want = np.frombuffer(blob, "<f4", 12 * 2048, 16).reshape(12, 2048)
    same = bool(np.array_equal(t, want))
    if not same:
        for w in range(12):
            C.memmove(L.refwb_pilut(w), want[w].ctypes.data, 2048 * 4)
    gen = pyref.RefGenerator(prg, SRATE)
    out = []
    t0 = time.perf_counter()
    for _ in range(ncalls):
        more, buf, n = gen.run(frames)
        out.append(buf.tobytes())
    conn.send((same, out, time.perf_counter() - t0))
    conn.close()


def start_parity(text, ncalls, frames=FRAMES):
    import multiprocessing as mp
    from saugns_b200 import generator as G
    ctx = mp.get_context("fork")
    a, b = ctx.Pipe()
    pr = ctx.Process(target=_parity_worker, args=(b, text, ncalls, frames, G.TABLES_PATH), daemon=True)
    pr.start()
    return a, pr


def finish_parity(handle, gpu_calls, what):
    """-> the `parity` object: max |GPU - reference| in LSB over the compared calls."""
    import hashlib
    import numpy as np
    conn, pr = handle
    same, ref_calls, secs = conn.recv()
    pr.join(timeout=10)
    mx, nz = 0, 0
    for g, r in zip(gpu_calls, ref_calls):
        d = np.abs(np.asarray(g, np.int32) - np.frombuffer(r, np.int16).astype(np.int32))
        mx = max(mx, int(d.max()) if d.size else 0)
        nz += int(np.count_nonzero(np.frombuffer(r, np.int16)))
    return {"checked": True, "max_lsb": mx, "calls": len(ref_calls), "what": what,
            "reference": "unmodified reference generator (oracle/_ref), whole script in one process",
            "nonzero_reference_samples": nz,
            "sha256_gpu": hashlib.sha256(b"".join(np.asarray(g, np.int16).tobytes() for g in gpu_calls)).hexdigest(),
            "sha256_reference": hashlib.sha256(b"".join(ref_calls)).hexdigest(),
            "host_tables_match_table_file": same, "reference_seconds": round(secs, 2)}


def run_reference_arm(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    if args.warmup + args.steps > SECS * SRATE // FRAMES:
        raise SystemExit("steps+warmup exceed the workload's calls")
    pool = ReferencePool(c3_text(1), procs)
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
        "config": c3_config(),
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
def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    return peak, ("measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650")


def program_bytes(prg):
    try:
        import ctypes
        from saugns_b200 import program as P
        pp = P.Program.from_address(prg.ptr)
        return (ctypes.sizeof(P.Program) + pp.ev_count * ctypes.sizeof(P.Event) +
                pp.op_count * (ctypes.sizeof(P.OpData) + 3 * ctypes.sizeof(P.Line)))
    except Exception:
        return 0


def latest_profile(pattern):
    try:
        import glob
        latest = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))[-1]
        prof = json.load(open(latest))
        prof["file"] = os.path.relpath(latest, ROOT)
        return prof
    except Exception:
        return {}


def run_gpu_arm(args):
    import numpy as np
    import torch
    import saugns_b200

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
    PARITY_CALLS = 2

    # the reference renders the same script's first calls on a host core meanwhile (rank 0)
    parity = None
    if rank == 0 and not args.no_cpu:
        try:
            parity = start_parity(c3_text(1), PARITY_CALLS)
        except Exception:
            parity = None

    prg = c3_program(1 + rank)

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
    g.set_timing(True)         # (undoes the call that ran ahead: the timed region is calls W+1 .. W+K for the events AND the kernel times)
    km0 = g.kernel_ms()
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
    km1 = g.kernel_ms()
    render_ms, mix_ms = km1[0] - km0[0], km1[1] - km0[1]
    launches = (c1[0] - c0[0]) + (c1[1] - c0[1])
    g.close()

    # ---- end to end through the public call with HOST buffers ----
    # timed: sau_create_Generator (program flattening + H2D upload of events, op-data,
    # bytecode, tables) + W+K x sauGenerator_run with a host PCM buffer (D2H every call)
    # + destroy; the W warm-up calls are part of the same render, so they are timed and
    # counted too (a renderer cannot skip the start of its script).
    prg_bytes = program_bytes(prg)
    barrier()
    first_calls = []
    t0 = time.perf_counter()
    g2 = saugns_b200.Generator(prg, SRATE, device=local_rank, stream=stream.cuda_stream,
                               max_call_len=FRAMES)
    for i in range(W + K):
        more, pcm, n = g2.run(FRAMES)
        if i < PARITY_CALLS:
            first_calls.append(pcm)
    g2.close()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_steps = W + K
    assert int(abs(pcm.astype("int32")).max()) > 0, "silent output"
    barrier()

    extra = {}
    if world > 1 and not args.no_extra:
        try:
            extra["c3_voice_sharded"] = leg_c3_sharded(args, dist, rank, local_rank, world)
        except Exception as e:      # an extra leg never takes the headline line down
            extra["c3_voice_sharded"] = {"error": repr(e)}
        try:
            extra["c5_script_sharded"] = leg_c5_sharded(args, dist, rank, local_rank, world)
        except Exception as e:
            extra["c5_script_sharded"] = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    vs_step = VOICES * FRAMES
    value = world * vs_step * K / (ms / 1000.0)
    e2e = world * vs_step * e2e_steps / e2e_s
    peak, peak_src = load_peaks()
    # algorithmic bytes (8 B per voice-sample = one f32 carrier store by the
    # render kernel + one f32 load by the mix kernel; pans are constant in C3, so no r rows):
    # 4 B per voice-sample for EACH of the two kernels' launches
    alg_bytes = 4.0 * vs_step
    mix_s = (mix_ms / K) / 1000.0
    mix_alg = 4.0 * vs_step + 4.0 * FRAMES
    rk_s = (render_ms / K) / 1000.0
    achieved = alg_bytes / rk_s / 1e9
    prof = latest_profile("r*_render_kernel.json")
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    line = {
        "metric": METRIC, "value": value, "unit": "voice-samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": c3_config(),
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
                     "note": "render_kernel is bound by instruction issue and the shared-memory data "
                             "pipe, not HBM (SURVEY.md 8d); the issue / pipe figures come from the "
                             "committed steady-state ncu capture named in ncu_profile, whose own "
                             "duration is ncu_duration_ms (this run's event timing: kernel_ms_per_launch)",
                     "ncu_profile": prof.get("file"),
                     "ncu_duration_ms": (prof.get("duration_ns") or 0) / 1e6 or None,
                     "issue_slot_frac_ncu": prof.get("issue_slot_frac"),
                     "smem_pipe_frac_ncu": (prof.get("lsu_wavefronts_pct") or 0) / 100.0 or None,
                     "fp64_pipe_frac_ncu": prof.get("fp64_pipe_frac"),
                     "xu_pipe_frac_ncu": prof.get("xu_pipe_frac"),
                     "warp_insts_per_32_op_samples_ncu": prof.get("warp_insts_per_32_op_samples"),
                     # the capture's instruction count over the issue slots of THIS run's kernel time
                     "issue_slot_frac_live": (prof["warp_insts"] / (4 * sms * sm_hz * rk_s)
                                              if prof.get("warp_insts") else None)},
    }
    if parity is not None:
        try:
            line["parity"] = finish_parity(
                parity, first_calls,
                f"first {PARITY_CALLS} calls ({PARITY_CALLS * FRAMES} frames x {VOICES} voices) of the timed "
                "end-to-end run's own PCM (program from ProgramBuilder) against the reference "
                "(script through its own parser)")
        except Exception as e:
            line["parity"] = {"checked": False, "error": repr(e)}
    if world == 1 and not args.no_cpu:
        try:
            cores = min(os.cpu_count() or 1, 64)
            pool = ReferencePool(c3_text(1), cores)
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
    if world == 1 and not args.no_extra:
        for name, leg in (("c2", leg_c2), ("c4", leg_c4), ("c5", leg_c5)):
            try:
                extra[name] = leg(args, local_rank, clocks)
            except Exception as e:
                extra[name] = {"error": repr(e)}
    if extra:
        line["extra"] = extra
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------
# the other BASELINE configs, as legs of the default run (`extra`)
# ---------------------------------------------------------------------------
C4_VOICES = 1024


def leg_c4(args, device, clocks, steps=20, warmup=3):
    """BASELINE config 4: 1024 voices with self-feedback PM carriers (W and R) and range-AM /
    ring modulation -- the sequential-per-sample path.  Every feedback operator is ONE serial
    chain over the render, so the figure of merit is cycles per feedback iteration."""
    import torch
    import saugns_b200
    from saugns_b200 import workloads
    text = workloads.synth_c4(C4_VOICES, SECS)
    par = None
    if not args.no_cpu:
        par = start_parity(text, 1)
    prg = workloads.build_c4(C4_VOICES, SECS)
    g = saugns_b200.Generator(prg, SRATE, device=device, max_call_len=FRAMES)
    for _ in range(warmup):
        g.run_device(FRAMES)
    g.set_timing(True)         # (undoes the call that ran ahead: events and kernel times cover the same calls)
    km0 = g.kernel_ms()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        g.run_device(FRAMES)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    rk, mk = g.kernel_ms()
    rk, mk = rk - km0[0], mk - km0[1]
    g.close()
    first = []
    t0 = time.perf_counter()
    g2 = saugns_b200.Generator(prg, SRATE, device=device, max_call_len=FRAMES)
    for i in range(warmup + steps):
        more, pcm, n = g2.run(FRAMES)
        if i < 1:
            first.append(pcm)
    g2.close()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    vs_step = C4_VOICES * FRAMES
    peak, peak_src = load_peaks()
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    prof = latest_profile("r*_c4_render_kernel.json")
    leg = {"metric": METRIC, "unit": "voice-samples/s", "value": vs_step * steps / (ms / 1000.0),
           "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "config": {"workload": "C4: 1024 voices, self-PM carriers (Wsin / Rlin / Wtri with p.a 0.3-1.0) "
                                  "with range-AM or ring modulation, 96 kHz stereo; step = one 24576-frame call",
                      "voices": C4_VOICES, "frames_per_step": FRAMES, "srate": SRATE},
           "realtime_factor": (FRAMES * steps / SRATE) / (ms / 1000.0),
           "e2e": {"value": vs_step * (warmup + steps) / e2e_s, "unit": "voice-samples/s",
                   "h2d_bytes_per_step": 40 + 12 + 6 * 12 + program_bytes(prg) // (warmup + steps),
                   "d2h_bytes_per_step": FRAMES * 2 * 2 + 8,
                   "timed": "create + calls with host PCM buffers + destroy, wall clock"},
           "roofline": {"bound": "latency", "kernel": "render_kernel",
                        "kernel_ms_per_launch": rk / steps, "mix_kernel_ms_per_launch": mk / steps,
                        "cycles_per_feedback_iteration": (rk / steps) * 1e-3 * sm_mhz * 1e6 / FRAMES,
                        "achieved": 4.0 * vs_step / (rk / steps / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": 4.0 * vs_step / (rk / steps / 1000.0) / 1e9 / peak, "peak_source": peak_src,
                        "traffic": prof.get("dram_bytes_per_launch"), "ncu_profile": prof.get("file"),
                        "note": "serial self-PM chains (SURVEY.md section 7): a call takes FRAMES dependent "
                                "feedback iterations whatever the voice count; the HBM fraction is reported "
                                "by contract, the bound is the iteration's dependent-issue latency"}}
    if par is not None:
        leg["parity"] = finish_parity(par, first, "first call of the end-to-end run against the reference")
        cores = min(os.cpu_count() or 1, 64)
        pool = ReferencePool(text, cores)
        pool.step(1)
        vs, wall = pool.step(8)
        used = pool.used
        pool.close()
        leg["cpu_baseline"] = {"value": vs / wall, "unit": "voice-samples/s", "cores": used, "kind": "reference",
                               "sample": f"unmodified reference generator, same script, calls 2..9, voices split "
                                         f"over {used} single-threaded processes; {wall:.2f} s wall"}
    return leg


def leg_c2(args, device, clocks):
    """BASELINE config 2: the reference's examples/misc1-4fm_pm.sau -- ONE voice alive at a time, range-FM nested up
    to three deep under a PM modulator (the few-voice case, where a GPU has nothing to spread over
    voices).  The whole 60 s render through the public call with host PCM buffers, against the unmodified reference
    on ONE host core (the reference is single-threaded and the script has one voice at a time)."""
    import numpy as np
    import torch
    import saugns_b200
    from saugns_b200 import workloads
    prg = workloads.build_c2()
    ncalls = (60 * SRATE + FRAMES - 1) // FRAMES
    par = None if args.no_cpu else start_parity(workloads.C2_TEXT, ncalls)

    def render_once():
        t0 = time.perf_counter()
        g = saugns_b200.Generator(prg, SRATE, device=device, max_call_len=FRAMES)
        calls, more = [], True
        while more:
            more, pcm, n = g.run(FRAMES)
            calls.append(pcm.copy())
        g.close()
        return time.perf_counter() - t0, calls

    first_s, _ = render_once()
    walls = []
    for _ in range(3):
        w, calls = render_once()
        walls.append(w)
    wall = sorted(walls)[1]
    # the kernels alone, PCM left on the device
    g = saugns_b200.Generator(prg, SRATE, device=device, max_call_len=FRAMES)
    torch.cuda.synchronize()
    g.set_timing(True)
    t0 = time.perf_counter()
    more, k = True, 0
    while more:
        more, _, n = g.run_device(FRAMES)
        k += 1
    torch.cuda.synchronize()
    dev_s = time.perf_counter() - t0
    rk, mk = g.kernel_ms()
    g.close()
    vs = 4 * 15 * SRATE                 # four voices of 15 s, one after the other
    leg = {"metric": METRIC, "unit": "voice-samples/s", "value": vs / dev_s, "ms_per_step": 1e3 * dev_s / k,
           "steps": k, "realtime_factor": 60.0 / dev_s,
           "config": {"workload": "C2: examples/misc1-4fm_pm.sau, 60 s, one voice slot (four 15 s voices in turn), "
                                  "7 / 4 / 5 / 5 operators with nested range-FM + PM; step = one 24576-frame call",
                      "frames_per_step": FRAMES, "srate": SRATE},
           "e2e": {"value": vs / wall, "unit": "voice-samples/s", "seconds_for_the_60_s_render": wall,
                   "first_render_of_the_process_s": first_s, "h2d_bytes_per_step": program_bytes(prg) // k,
                   "d2h_bytes_per_step": FRAMES * 2 * 2 + 8,
                   "timed": "create + every call with a host PCM buffer + destroy, wall clock (median of 3)"},
           "roofline": {"bound": "latency", "kernel": "render_kernel", "kernel_ms_per_launch": rk / k,
                        "mix_kernel_ms_per_launch": mk / k,
                        "note": "one voice alive: its team of warps spans several CTAs (DESIGN.md 3.1c); every member is "
                                "a latency-bound chain, the HBM roofline does not apply"}}
    if par is not None:
        leg["parity"] = finish_parity(par, calls, "every call of the 60 s render against the reference")
        secs = leg["parity"]["reference_seconds"]
        leg["cpu_baseline"] = {"value": vs / secs, "unit": "voice-samples/s", "cores": 1, "kind": "reference",
                               "seconds_for_the_60_s_render": secs,
                               "sample": "unmodified reference generator (oracle/_ref), the whole script in one "
                                         "process on one core (the reference is single-threaded)"}
        leg["speedup_vs_one_core_e2e"] = secs / wall
    return leg


def _cli_render(job):
    """One `saugns -m -d -o x.wav script.sau` of the stock reference CLI (the C5
    baseline is the reference's own program, one script per process, `xargs -P <cores>`)."""
    cli, src, out = job
    r = subprocess.run([cli, "-m", "-d", "-r", str(SRATE), "-o", out, src], capture_output=True)
    return r.returncode


def leg_c5(args, device, clocks):
    """BASELINE config 5: independent mixed scripts (4-16 voices of W+PM / N / R / swept W with
    range-AM, 1-10 s), this GPU's share of the 10 000 (1250 = 10 000 / 8), through the NATIVE
    batched driver (saugen_render_batch, csrc/batch_driver.cpp).  value: every script's PCM
    delivered to page-locked host arrays (no files); e2e: every script written as the
    reference's WAV file on a RAM disk by the driver's writer threads -- what the CPU baseline
    (the stock reference CLI, one process per script, all cores) does."""
    import hashlib
    import shutil
    import tempfile
    import numpy as np
    import torch
    import saugns_b200
    from saugns_b200 import workloads, batch
    from saugns_b200 import program as P
    n = args.scripts
    t0 = time.perf_counter()
    prgs = [workloads.build_c5_script(i) for i in range(n)]
    build_s = time.perf_counter() - t0

    def voice_samples(plist):
        vs = 0
        for p in plist:
            pp = P.Program.from_address(p.ptr)
            for e in range(pp.ev_count):
                ev = pp.events[e]
                for k in range(ev.op_data_count):
                    if ev.op_data[k].id == ev.carr_op_id:
                        vs += ev.op_data[k].time.v_ms * SRATE // 1000
        return vs

    vs = voice_samples(prgs)
    ramdisk = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="c5_", dir=ramdisk)
    try:
        batch.render_batch_native(prgs[:32], srate=SRATE, device=device, group_size=16)      # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch.render_batch_native(prgs, srate=SRATE, device=device, group_size=args.group, depth=args.depth,
                                  call_len=args.call_frames, discard=True)
        wall_cold = time.perf_counter() - t0       # the process's first batch: page-locked arrays are made here
        t0 = time.perf_counter()
        batch.render_batch_native(prgs, srate=SRATE, device=device, group_size=args.group, depth=args.depth,
                                  call_len=args.call_frames, discard=True)
        wall_dev = time.perf_counter() - t0
        paths = [os.path.join(tmp, f"g{i}.wav") for i in range(n)]
        t0 = time.perf_counter()
        batch.render_batch_native(prgs, srate=SRATE, device=device, group_size=args.group, depth=args.depth,
                                  call_len=args.call_frames, wav_paths=paths, io_threads=8)
        wall = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(p) for p in paths)
        frames = [(nbytes - 44 * n) // 4]
        leg = {"metric": METRIC, "unit": "voice-samples/s", "value": vs / wall_dev,
               "config": {"workload": f"C5: {n} independent mixed scripts (one GPU's share of 10 000 over 8), "
                                      f"4-16 voices each (W+PM / N / R / swept W + range-AM), 1-10 s, 96 kHz stereo",
                          "scripts": n, "srate": SRATE, "call_frames": args.call_frames, "group": args.group},
               "scripts_per_s": n / wall_dev, "first_batch_scripts_per_s": n / wall_cold, "audio_s": frames[0] / SRATE,
               "realtime_factor": (frames[0] / SRATE) / wall_dev,
               "timed": "saugen_render_batch: create + batched calls + destroy of every script, PCM into "
                        "page-locked host arrays and dropped there (value; first_batch_scripts_per_s: the same "
                        "batch when it was the process's first, the page-locked arrays still to be made); the same with the WAV files "
                        "written to a RAM disk by 8 writer threads (e2e); programs built beforehand "
                        "(program_build_s; the reference CLI's time includes its parser)",
               "e2e": {"value": vs / wall, "unit": "voice-samples/s", "scripts_per_s": n / wall, "wall_s": wall,
                       "h2d_bytes_per_step": sum(program_bytes(p) for p in prgs) // n,
                       "d2h_bytes_per_step": frames[0] * 4 // n, "wav_bytes": nbytes, "step": "one script"},
               "program_build_s": build_s}
        if not args.no_cpu:
            from oracle import pyref
            from concurrent.futures import ThreadPoolExecutor
            cores = min(os.cpu_count() or 1, 64)
            m = min(n, 20 * cores)             # bounded sample: ~20 scripts per core
            jobs = []
            for i in range(m):
                src = os.path.join(tmp, f"s{i}.sau")
                with open(src, "w") as f:
                    f.write(workloads.synth_c5_script(i))
                jobs.append((pyref.REF_EXE, src, os.path.join(tmp, f"s{i}.wav")))
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:          # one reference process per script, `cores` at a time
                rcs = list(ex.map(_cli_render, jobs))
            cpu_wall = time.perf_counter() - t0
            assert not any(rcs), "reference CLI failed"
            vs_m = voice_samples(prgs[:m])
            bad = 0
            for i in range(m):                             # byte-identical files
                with open(jobs[i][2], "rb") as fa, open(paths[i], "rb") as fb:
                    if hashlib.sha256(fa.read()).digest() != hashlib.sha256(fb.read()).digest():
                        bad += 1
            leg["cpu_baseline"] = {"value": vs_m / cpu_wall, "unit": "voice-samples/s", "cores": cores,
                                   "kind": "reference", "scripts_per_s": m / cpu_wall,
                                   "sample": f"the stock reference CLI (`saugns -m -d -o x.wav`, oracle/_ref), one "
                                             f"process per script, {cores} at a time, first {m} scripts, WAV files "
                                             f"on a RAM disk; {cpu_wall:.2f} s wall (parse + render + write)"}
            leg["parity"] = {"checked": True, "scripts_compared": m, "scripts_differing": bad,
                             "max_lsb": 0 if bad == 0 else None,
                             "what": "the WAV file the native batch driver wrote for each of the sample's scripts "
                                     "(program from ProgramBuilder) against the file the stock reference CLI wrote "
                                     "for the same script text: byte-identical files"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return leg


def leg_c5_sharded(args, dist, rank, local_rank, world):
    """BASELINE config 5 over the ranks: `--scripts` scripts PER GPU (1250 x 8 = the 10 000), script i
    on rank i % world, every rank through the native batched driver with its WAV files written to
    a RAM disk; no collective on the data path.  Wall clock from a barrier to the last rank's end."""
    import shutil
    import tempfile
    import torch
    from saugns_b200 import workloads, batch
    from saugns_b200 import program as P
    n_total = args.scripts * world
    mine = list(range(rank, n_total, world))
    prgs = [workloads.build_c5_script(i) for i in mine]
    vs = 0
    for p in prgs:
        pp = P.Program.from_address(p.ptr)
        for e in range(pp.ev_count):
            ev = pp.events[e]
            for k in range(ev.op_data_count):
                if ev.op_data[k].id == ev.carr_op_id:
                    vs += ev.op_data[k].time.v_ms * SRATE // 1000
    ramdisk = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix=f"c5r{rank}_", dir=ramdisk)
    try:
        batch.render_batch_native(prgs[:32], srate=SRATE, device=local_rank, group_size=16)      # warm-up
        torch.cuda.synchronize()
        paths = [os.path.join(tmp, f"g{i}.wav") for i in mine]
        dist.barrier()
        t0 = time.perf_counter()
        batch.render_batch_native(prgs, srate=SRATE, device=local_rank, group_size=args.group, depth=args.depth,
                                  call_len=args.call_frames, discard=True)
        wall_cold = time.perf_counter() - t0       # the process's first batch: page-locked arrays are made here
        dist.barrier()
        t0 = time.perf_counter()
        batch.render_batch_native(prgs, srate=SRATE, device=local_rank, group_size=args.group, depth=args.depth,
                                  call_len=args.call_frames, discard=True)
        wall_dev = time.perf_counter() - t0        # PCM to page-locked host arrays, dropped there
        dist.barrier()
        t0 = time.perf_counter()
        batch.render_batch_native(prgs, srate=SRATE, device=local_rank, group_size=args.group, depth=args.depth,
                                  call_len=args.call_frames, wav_paths=paths, io_threads=4)
        wall = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(p) for p in paths)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    t = torch.tensor([wall, float(vs), float(nbytes), wall_cold, wall_dev], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    wall_max, vs_all, bytes_all = float(tmax[0].item()), float(t[1].item()), float(t[2].item())
    cold_max, dev_max = float(tmax[3].item()), float(tmax[4].item())
    if rank != 0:
        return None
    return {"metric": METRIC, "unit": "voice-samples/s", "scaling": "weak", "n_gpus": world,
            "value": vs_all / dev_max, "scripts": n_total, "scripts_per_s": n_total / dev_max,
            "e2e": {"value": vs_all / wall_max, "unit": "voice-samples/s", "scripts_per_s": n_total / wall_max,
                    "wall_s": wall_max, "wav_bytes": bytes_all, "step": "one script"},
            "first_batch_scripts_per_s": n_total / cold_max,
            "config": {"workload": f"C5: {n_total} independent mixed scripts dealt to {world} GPUs ({args.scripts} each), "
                                   f"every WAV file written to a RAM disk", "srate": SRATE,
                       "call_frames": args.call_frames, "group": args.group},
            "timed": "per rank: saugen_render_batch over its scripts (create + batched calls + destroy), PCM to "
                     "page-locked host arrays and dropped there (value), and saugen_render_batch_wav with 4 writer "
                     "threads per rank, every WAV file on a RAM disk (e2e); wall clock, max over ranks; programs built beforehand; "
                     "first_batch_scripts_per_s: the same scripts as each process's first batch, PCM to host and "
                     "dropped (the page-locked arrays are made there)"}


def leg_c3_sharded(args, dist, rank, local_rank, world, steps=20, warmup=3):
    """ONE C3 script (seed 1), its voices spread over the ranks (strong scaling): each rank
    renders its voices' float mix planes, one NCCL sum-reduce per call, the root converts.
    Compared with the same script rendered unsharded on the root's GPU."""
    import numpy as np
    import torch
    import saugns_b200
    from saugns_b200 import multigpu
    prg = c3_program(1)
    vg = multigpu.VoiceShardedGenerator(prg, SRATE, device=local_rank, max_call_len=FRAMES)
    first = []
    gen = getattr(vg.shard, "gen", None)
    if gen is not None:
        gen.set_timing(True)       # (before the warm-up: switching it re-launches a call that ran ahead)
    for i in range(warmup):
        more, pcm, n = vg.run(FRAMES)
        if rank == 0 and i < 2:
            first.append(pcm.copy())
    dist.barrier()
    torch.cuda.synchronize()
    km0 = gen.kernel_ms() if gen is not None else (0.0, 0.0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        vg.run(FRAMES)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # each rank's own kernels (device-timed, per call): what the collective and the host add is ms/steps minus these
    km = gen.kernel_ms() if gen is not None else (0.0, 0.0)
    per_rank = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
    per_rank[2 * rank], per_rank[2 * rank + 1] = (km[0] - km0[0]) / steps, (km[1] - km0[1]) / steps
    dist.all_reduce(per_rank)
    per_rank = [round(float(x), 4) for x in per_rank.tolist()]
    ncoll = getattr(vg, "collectives", None)
    if ncoll is not None:
        ncoll = ncoll / float(warmup + steps)
    vg.close()
    leg = None
    if rank == 0:
        g = saugns_b200.Generator(prg, SRATE, device=local_rank, max_call_len=FRAMES)
        mx = 0
        for i in range(len(first)):
            more, pcm, n = g.run(FRAMES)
            mx = max(mx, int(np.abs(pcm.astype(np.int32) - first[i].astype(np.int32)).max()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.run(FRAMES)
        e1.record()
        torch.cuda.synchronize()
        one_ms = e0.elapsed_time(e1)
        g.close()
        leg = {"metric": METRIC, "unit": "voice-samples/s", "scaling": "strong", "n_gpus": world,
               "value": VOICES * FRAMES * steps / (ms / 1000.0), "ms_per_step": ms / steps,
               "one_gpu_ms_per_step": one_ms / steps, "speedup_vs_one_gpu": one_ms / ms,
               "config": {"workload": "C3, ONE 4096-voice script voice-sharded over the ranks: one reduce of the "
                                      "float L/R planes per call, PCM on the root's host buffer",
                          "voices": VOICES, "frames_per_step": FRAMES},
               "collectives_per_call": ncoll,
               "shard_render_ms_per_call": per_rank[0::2], "shard_mix_ms_per_call": per_rank[1::2],
               "parity": {"checked": True, "max_lsb": mx, "calls": len(first),
                          "what": "sharded PCM against the same script rendered unsharded on one GPU "
                                  "(the float summation order across ranks differs: <= 1 LSB allowed)"}}
    return leg


def run_gpu_config(args):
    """One of the `extra` legs on its own (developer use): --workload c4 | c5 | c3-sharded."""
    import torch
    rank, local_rank, world = env_rank()
    torch.cuda.set_device(local_rank)
    clocks = {"sm_mhz": None}
    if args.workload == "c3-sharded":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        leg = leg_c3_sharded(args, dist, rank, local_rank, world, steps=args.steps)
        if rank == 0:
            print(json.dumps(leg))
        dist.destroy_process_group()
        return 0
    leg = leg_c4(args, local_rank, clocks, steps=args.steps) if args.workload == "c4" else \
        leg_c5(args, local_rank, clocks)
    print(json.dumps(leg))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs")
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
