"""brief of an .ncu-rep: headline metrics + stall reasons + top stalled SASS lines.  usage: ncu_brief.py rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:90], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
        if k in d: print(f"  {k} = {d[k]}")
    st = sorted(((float(v or 0), h) for h, v in d.items() if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h), reverse=True)
    for v, h in st[:7]:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in data)
print("samples", tot)
for r in sorted(data, key=lambda r: -int(r[isamp]))[:ntop]:
    print(f"  {r[ia][-5:]} {int(r[isamp]):6d} {100.0*int(r[isamp])/max(tot,1):5.1f}%  {r[isrc][:100]}")
