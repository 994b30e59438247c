"""Turn ncu reports into the committed summaries under profiles/ (run in the build container: no GPU needed).

    python tools/ncu_summary.py full  gpurun_out/p_tc.ncu-rep  profiles/r02_k_tc_pass_ncu_full.md  "title" [--traffic k_tc_pass]
    python tools/ncu_summary.py list  gpurun_out/p_launches_render.csv  profiles/r02_launch_list.md  "title"

`full`: one column per captured launch with the metrics B200_PROFILING.md names; `--traffic NAME` also writes
profiles/r02_ncu_traffic.json[NAME] = dram bytes (read + write) per launch, which bench.py reports as `roofline.traffic`
with its provenance.  `list`: per-kernel launch counts, total time and share of the captured window."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    return names, units, rows[hdr + 2:]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full(rep, dst, title, traffic_name=None):
    names, units, rows = raw_rows(rep)
    kcol = names.index("Kernel Name")
    cols = [(m, names.index(m)) for m in METRICS if m in names]
    lines = [f"# {title}", "", f"Source: `{os.path.relpath(rep, ROOT)}` (kept out of the repository; parsed with `ncu -i ... --page raw --csv` by "
             "tools/ncu_summary.py).  Numbers taken under the profiler are NOT bench values.", ""]
    head = "| metric | unit | " + " | ".join(f"{r[kcol].split('(')[0][-34:]} #{i}" for i, r in enumerate(rows)) + " |"
    lines += [head, "|---|---|" + "---|" * len(rows)]
    for m, c in cols:
        lines.append(f"| {m} | {units[c]} | " + " | ".join(r[c] for r in rows) + " |")
    open(dst, "w").write("\n".join(lines) + "\n")
    if traffic_name:
        rd, wr = names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
        per = [to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]) for r in rows]
        p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        # k_tc_pass<FMT, PASSES, STASH, FUSED>: the one-launch frame variant has FUSED = 1; the others are the separate coarse and
        # fine launches (in that order) of the same frame
        ok = [(b, r) for b, r in zip(per, rows) if b == b]          # (a launch ncu failed to collect reads as NaN: left out)
        frame = [b for b, r in ok if r[kcol].split("(")[0].rstrip().endswith(", 1>")]
        sep = [b for b, r in ok if not r[kcol].split("(")[0].rstrip().endswith(", 1>")] or [b for b, _ in ok]
        d[traffic_name] = {"frame_bytes_per_launch": int(frame[-1]) if frame else None,
                           "mean_bytes_per_launch": int(sum(sep) / len(sep)), "fine_bytes_per_launch": int(sep[-1]),
                           "bytes_per_launch": [int(x) if x == x else None for x in per],
                           "source": f"{os.path.relpath(dst, ROOT)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
        json.dump(d, open(p, "w"), indent=1)
    print("\n".join(lines[:12]))


def launch_list(src, dst, title):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][-44:]
        t = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {title}", "", "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    mode, src, dst, title = sys.argv[1:5]
    if mode == "full":
        tn = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
        full(src, dst, title, tn)
    else:
        launch_list(src, dst, title)
