"""Turn gpurun_out/<round>_launches.csv and <round>_forward.ncu-rep into the tracked summaries under profiles/.

    python scripts/make_profile_summary.py r01

Writes profiles/<round>_launches.csv.gz, <round>_launches_summary.md, <round>_kernels.json (per-kernel DRAM bytes,
duration and pipe utilisation of one forward; bench.py reads the DRAM traffic of the dominant kernel from it) and
<round>_kernels.md."""
import csv, gzip, json, os, re, subprocess, sys
from collections import OrderedDict, defaultdict

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
# optional: what was captured (default: one inference forward); e.g. "the backward / loss / optimiser kernels of ..."
WHAT = sys.argv[2] if len(sys.argv) > 2 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0]


def launches():
    path = os.path.join(GO, f"{R}_launches.csv")
    if not os.path.exists(path):
        return
    raw = open(path).read()
    lines = [l for l in raw.splitlines() if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    i_name, i_val, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        v = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        a = agg[short(r[i_name])]
        a[0] += 1
        a[1] += v
        n += 1
    total = sum(a[1] for a in agg.values())
    with gzip.open(os.path.join(PR, f"{R}_launches.csv.gz"), "wt") as f:
        f.write(raw)
    with open(os.path.join(PR, f"{R}_launches_summary.md"), "w") as f:
        f.write(f"# {R} -- ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train`\n\n"
                "Command (under gpurun, directly after the same command exited 0 without ncu; scripts/profile_round.sh):\n\n"
                f"    ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/{R}_launches.csv \\\n"
                "        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train\n\n"
                f"{n} launches, {total / 1000.0:.2f} ms of kernel time in total (3 warm-up + 2 per-kernel-table + 2 timed + 2 end-to-end "
                "forward passes of 32 x 1028 clouds, plus the L2 flush fills).  Per-launch times are cold-cache and serialised: compare SHARES with "
                "bench.py's `kernels` table, not absolutes.  `--no-graph` so that every kernel is its own launch row "
                "(the benchmark proper replays the same launches as one CUDA graph).\n"
                f"Raw list: `profiles/{R}_launches.csv.gz`.\n\n| share | launches | total us | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * a[1] / total:.2f}% | {a[0]} | {a[1]:.1f} | `{k[:110]}` |\n")
    print("launch summary:", n, "launches")


WANT = OrderedDict([
    ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
])
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_US = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}


def forward():
    raw = os.path.join(GO, f"{R}_forward_raw.csv")
    rep = os.path.join(GO, f"{R}_forward.ncu-rep")
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units = rows[0], rows[1]
    cols = {b: hdr.index(a) for a, b in WANT.items() if a in hdr}
    i_name = hdr.index("Kernel Name")
    per = OrderedDict()
    launches_ = []
    for r in rows[2:]:
        d = {"kernel": short(r[i_name])}
        for b, i in cols.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                v = None
            if v is not None and b in ("dram_rd", "dram_wr"):
                v *= TO_BYTES.get(units[i], 1)
            if v is not None and b == "us":
                v *= TO_US.get(units[i], 1)
            d[b] = v
        launches_.append(d)
        a = per.setdefault(d["kernel"], {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "tensor_pct_time_weighted": 0.0,
                                         "fma_pct_time_weighted": 0.0, "ipc_time_weighted": 0.0})
        a["launches"] += 1
        a["us"] += d["us"] or 0
        a["dram_bytes"] += (d.get("dram_rd") or 0) + (d.get("dram_wr") or 0)
        for key, src in (("tensor_pct_time_weighted", "tensor_pct"), ("fma_pct_time_weighted", "fma_pct"), ("ipc_time_weighted", "ipc")):
            a[key] += (d.get(src) or 0) * (d["us"] or 0)
    for a in per.values():
        for key in ("tensor_pct_time_weighted", "fma_pct_time_weighted", "ipc_time_weighted"):
            a[key] = a[key] / a["us"] if a["us"] else 0.0
    json.dump({"source": f"ncu --set full --clock-control none, one forward of 32 x 1028 clouds (scripts/profile_round.sh {R})",
               "per_kernel": per, "launches": launches_}, open(os.path.join(PR, f"{R}_kernels.json"), "w"), indent=1)
    with open(os.path.join(PR, f"{R}_kernels.md"), "w") as f:
        if WHAT:
            f.write(f"# {R} -- ncu sections of {WHAT}\n\nCommand: `scripts/profile_train.sh {R}`.  Times are under the profiler (cold caches, "
                    "serialised) -- use them for shares and per-kernel diagnosis, never as benchmark values.  dram = `dram__bytes_read.sum + "
                    "dram__bytes_write.sum`.\n\n")
        else:
            f.write(f"# {R} -- `ncu --set full` of every library kernel in one forward (32 x 1028 clouds, eval, eager launches)\n\n"
                "Command: `scripts/profile_round.sh` step 2 (`-k regex:<library kernel names> -s <launches of the two warm-up forwards> -c <launches of "
                "one forward>`, i.e. the third forward of `scripts/profile_forward.py`).  Times are under the profiler (cold caches, serialised) -- use them for "
                "shares and per-kernel diagnosis, never as benchmark values.  dram = `dram__bytes_read.sum + dram__bytes_write.sum`.\n\n")
        f.write(
                "## Per kernel (summed over its launches in the forward)\n\n"
                "| kernel | launches | us | share | DRAM MB | tensor pipe % | FMA pipe % | IPC |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        tot = sum(a["us"] for a in per.values())
        for k, a in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"| `{k[:70]}` | {a['launches']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['dram_bytes'] / 1e6:.1f} | "
                    f"{a['tensor_pct_time_weighted']:.1f} | {a['fma_pct_time_weighted']:.1f} | {a['ipc_time_weighted']:.2f} |\n")
        f.write("\n## Every launch\n\n| # | kernel | us | grid | block | regs | DRAM rd MB | DRAM wr MB | dram % | sm % | IPC | occ % | FMA % | tensor % | "
                "stall long_sb | short_sb | barrier | wait |\n|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        g = lambda d, k, fmt="{:.1f}": "-" if d.get(k) is None else fmt.format(d[k])
        for i, d in enumerate(launches_):
            f.write(f"| {i} | `{d['kernel'][:48]}` | {g(d, 'us')} | {g(d, 'grid', '{:.0f}')} | {g(d, 'block', '{:.0f}')} | {g(d, 'regs', '{:.0f}')} | "
                    f"{'-' if d.get('dram_rd') is None else '%.2f' % (d['dram_rd'] / 1e6)} | {'-' if d.get('dram_wr') is None else '%.2f' % (d['dram_wr'] / 1e6)} | "
                    f"{g(d, 'dram_pct')} | {g(d, 'sm_pct')} | {g(d, 'ipc', '{:.2f}')} | {g(d, 'occ_pct')} | {g(d, 'fma_pct')} | {g(d, 'tensor_pct')} | "
                    f"{g(d, 'stall_long_sb', '{:.2f}')} | {g(d, 'stall_short_sb', '{:.2f}')} | {g(d, 'stall_barrier', '{:.2f}')} | {g(d, 'stall_wait', '{:.2f}')} |\n")
    print("forward summary:", len(launches_), "launches")


launches()
forward()
