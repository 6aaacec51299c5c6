#!/usr/bin/env python
"""Turns raw Nsight Compute output (gpurun_out/, scratch) into the small tracked summaries under profiles/.
  python profiles/summarize.py launches <launches.csv> <out.md>        per-kernel time shares of one bench run
  python profiles/summarize.py full <tag> <prof_*.ncu-rep | raw_*.csv ...>   key counters of `ncu --set full` captures -> profiles/<tag>.json + .md
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_membar"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("s2s::", "")


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    rows = [r for r in rows if "spin_kernel" not in r[4] and "at::native" not in r[4]]     # torch.cuda._sleep / L2-flush fills of bench.py: not the library's
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += float(r[14])
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path.split('/')[-1]}): {len(rows)} launches, {tot / 1e6:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)\n\n")
        streams = collections.Counter()
        for r in rows:
            streams[r[6]] += float(r[14])
        f.write("kernel time per CUDA stream id as ncu reports it (eager warm-up steps, the captured step and the graph replays appear under different ids; "
                "within one step the recurrence kernels run on the main stream and the deferred weight gradients / V1 on the low-priority side stream): "
                + ", ".join(f"stream {k}: {v / 1e6:.2f} ms" for k, v in streams.most_common()) + "\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v[1] / tot < 0.002:
                continue
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.1f} |\n")
    print(open(out).read())


def full(tag, reps):
    out = []
    for rep in reps:
        # a .ncu-rep, or the `ncu -i X.ncu-rep --page raw --csv` export made on the GPU box (the reports themselves are too large to bring back)
        txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"kernel": short(r[hdr.index("Kernel Name")]), "report": rep.split("/")[-1]}
            for k in KEYS:
                if k in hdr:
                    try:
                        d[k] = float(r[hdr.index(k)]); d[k + ".unit"] = units[hdr.index(k)]
                    except ValueError:
                        pass
            out.append(d)
    json.dump(out, open(f"profiles/{tag}.json", "w"), indent=1)
    with open(f"profiles/{tag}.md", "w") as f:
        f.write(f"# ncu --set full summaries ({tag}); one row per captured launch\n\n")
        f.write("| kernel | time | DRAM read | DRAM write | regs | grid x block | warps active % | DRAM thr % | SM thr % | tensor pipe % (active / elapsed) | top stalls (pc samples) |\n|---|---:|---:|---:|---:|---|---:|---:|---:|---|---|\n")
        for d in out:
            def g(k, dflt=0.0):
                return d.get(k, dflt)
            def to_bytes(k):
                u = d.get(k + ".unit", "byte"); v = g(k)
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tu = d.get("gpu__time_duration.sum.unit", "ns"); t = g("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(tu, 1)
            stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): v for k, v in d.items() if k.startswith("smsp__pcsamp") and isinstance(v, float)}
            top = ", ".join(f"{k} {int(v)}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
            f.write(f"| `{d['kernel']}` | {t:.1f} us | {to_bytes('dram__bytes_read.sum') / 1e6:.2f} MB | {to_bytes('dram__bytes_write.sum') / 1e6:.2f} MB | "
                    f"{int(g('launch__registers_per_thread'))} | {int(g('launch__grid_size'))} x {int(g('launch__block_size'))} | "
                    f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                    f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                    f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} / {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {top} |\n")
    print(open(f"profiles/{tag}.md").read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3:])
