#!/usr/bin/env python
"""Developer aid: turn the ncu captures of one gpurun cycle (gpurun_out/) into
the tracked summaries under profiles/.

usage: tools/make_profile.py TAG [render.ncu-rep] [mix.ncu-rep] [launches.csv]
  e.g. tools/make_profile.py r01 gpurun_out/r01_render.ncu-rep gpurun_out/r01_mix.ncu-rep \
          gpurun_out/r01_launches.csv
Writes profiles/TAG_render_kernel.{json,txt}, profiles/TAG_mix_kernel.{json,txt},
profiles/TAG_launches.csv (+ a per-kernel share table at its top as comments).
bench.py reads profiles/<latest>_render_kernel.json for roofline.traffic."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "duration_ns": "gpu__time_duration.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "dyn_smem_bytes": "launch__shared_mem_per_block_dynamic",
    "warp_insts": "smsp__inst_executed.sum",
    "ipc_active": "sm__inst_executed.avg.per_cycle_active",
    "ipc_elapsed": "sm__inst_executed.avg.per_cycle_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "pipe_fp64_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "local_ld_requests": "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "local_st_requests": "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "thread_inst_per_warp_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "branch_efficiency_pct": "smsp__sass_average_branch_targets_threads_uniform.pct",
    "cycles_elapsed": "sm__cycles_elapsed.max",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return hdr, dict(zip(hdr, units)), dict(zip(hdr, vals))


def num(m, u, k):
    if k not in m or m[k] in ("", "n/a"):
        return None
    v = float(m[k].replace(",", ""))
    return v * UNIT.get(u.get(k, ""), 1.0)


def summarize(rep, name, units_per_launch, unit_name):
    hdr, u, m = raw_page(rep)
    d = {"kernel": m.get("Kernel Name", name), "source": os.path.basename(rep)}
    for k, mk in KEYS.items():
        d[k] = num(m, u, mk)
    stalls = {}
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            stalls[k.split("stalled_")[1].split("_per_issue")[0]] = float(m[k])
    d["stall_cycles_per_issue"] = {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])
                                   if v >= 0.05}
    d["dram_bytes_per_launch"] = (d["dram_read"] or 0) + (d["dram_write"] or 0)
    d["issue_slot_frac"] = (d["issue_active_pct"] or 0) / 100.0
    d["fp64_pipe_frac"] = (d["pipe_fp64_pct"] or 0) / 100.0
    d["xu_pipe_frac"] = (d["pipe_xu_pct"] or 0) / 100.0
    d[unit_name + "_per_launch"] = units_per_launch
    if d["warp_insts"] and units_per_launch:
        d["warp_insts_per_32_" + unit_name] = d["warp_insts"] / (units_per_launch / 32.0)
    if d["smem_wavefronts"]:
        d["smem_conflict_share"] = (d["smem_bank_conflicts"] or 0) / d["smem_wavefronts"]
    return d


def sass_mix(rep, per):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, byop, tot, nstatic = None, collections.Counter(), 0, 0
    for r in rows:
        if "Instructions Executed" in r and "Source" in r:
            h = r
            ia, isrc = h.index("Instructions Executed"), h.index("Source")
            continue
        if h is None or len(r) <= ia:
            continue
        try:
            cnt = int(r[ia])
        except ValueError:
            continue
        t = r[isrc].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        parts = op.split(".")
        op = ".".join(parts[:3]) if parts[0] in ("F2F", "I2F", "F2I", "MUFU", "LDS", "STS", "LDL", "STL",
                                                 "LDG", "STG", "I2FP") else parts[0]
        byop[op] += cnt
        tot += cnt
        nstatic += 1
    lines = [f"static SASS instructions: {nstatic}"]
    for op, cnt in byop.most_common(32):
        lines.append(f"  {op:18s} {cnt / per:8.2f} per 32 units  {100.0 * cnt / max(tot, 1):5.1f}%")
    return lines


def main():
    tag = sys.argv[1]
    render = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/{tag}_render.ncu-rep"
    mix = sys.argv[3] if len(sys.argv) > 3 else f"gpurun_out/{tag}_mix.ncu-rep"
    launches = sys.argv[4] if len(sys.argv) > 4 else f"gpurun_out/{tag}_launches.csv"
    pdir = os.path.join(ROOT, "profiles")
    os.makedirs(pdir, exist_ok=True)
    c4 = f"gpurun_out/{tag}_c4.ncu-rep"
    jobs = [(render, "render_kernel", 3 * 4096 * 24576, "op_samples"),
            (mix, "mix_kernel", 4096 * 24576, "voice_samples"),
            (c4, "c4_render_kernel", 1024 * 24576, "voice_samples")]
    for rep, name, units, uname in jobs:
        if not os.path.exists(rep):
            continue
        d = summarize(rep, name, units, uname)
        if name == "c4_render_kernel":
            # C4 (1024 self-PM voices): every feedback operator is one serial chain of 24576
            # dependent iterations per call; the figure of merit is cycles per iteration
            d["feedback_iterations_per_launch"] = 24576
            d["cycles_per_feedback_iteration"] = (d["cycles_elapsed"] or 0) / 24576.0
        with open(os.path.join(pdir, f"{tag}_{name}.json"), "w") as f:
            json.dump(d, f, indent=1)
        with open(os.path.join(pdir, f"{tag}_{name}.txt"), "w") as f:
            what = ("the C4 step (1024 self-PM voices x 24576 frames)" if name == "c4_render_kernel" else
                    "the C3 step (4096 voices x 24576 frames)")
            f.write(f"# {name}: ncu --set full --clock-control none --cache-control none, one steady-state launch "
                    f"of {what}, from {os.path.basename(rep)}\n")
            for k, v in d.items():
                f.write(f"{k:32s} {v}\n")
            f.write("\n".join(sass_mix(rep, units / 32.0)) + "\n")
        print("wrote", name, "dram bytes/launch", d["dram_bytes_per_launch"], "issue", d["issue_slot_frac"])
    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
        hdr = rows[0]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.defaultdict(list)
        for r in rows[1:]:
            try:
                agg[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")))
            except ValueError:
                pass
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(pdir, f"{tag}_launches.csv"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py "
                    "--steps 5 --warmup 3 --no-cpu\n# kernel, launches, mean ns, share of GPU time\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                f.write(f"# {k}, {len(v)}, {sum(v) / len(v):.0f}, {sum(v) / tot:.3f}\n")
            w = csv.writer(f)
            w.writerow(["ID", "Kernel", "Block", "Grid", "ns"])
            ib, ig = hdr.index("Block Size"), hdr.index("Grid Size")
            for r in rows[1:]:
                w.writerow([r[0], r[ik].split("(")[0], r[ib], r[ig], r[iv]])
        print({k: (len(v), sum(v) / len(v)) for k, v in agg.items()})


if __name__ == "__main__":
    main()
