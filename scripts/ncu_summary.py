"""Extract the judged metrics of a .ncu-rep (ncu --set full) into a small CSV for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof_c2.ncu-rep profiles/r01x_ncu_full_c2_summary.csv"""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
out = [["metric", "unit"] + [f"launch{i}" for i in range(len(data))]]
out.append(["Kernel Name", ""] + [r[hdr.index("Kernel Name")] for r in data])
for i, h in enumerate(hdr):
    if h in WANT or (h.startswith(STALLS) and "not_issued" not in h):
        out.append([h, units[i]] + [r[i] for r in data])
with open(sys.argv[2], "w", newline="") as f:
    csv.writer(f).writerows(out)
print(f"{sys.argv[2]}: {len(out) - 2} metrics x {len(data)} launch(es)")
