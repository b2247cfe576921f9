"""Print selected metrics of every kernel in an .ncu-rep (reads `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys, re
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fp16.sum", "smsp__inst_executed_pipe_uniform.sum",
        "smsp__inst_executed_pipe_tmem.sum", "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fmalite.sum"]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:60])
    for i, h in enumerate(hdr):
        if h in KEYS or (pat and pat.search(h)):
            print(f"  {h:90s} {r[i]:>18s} {units[i]}")
