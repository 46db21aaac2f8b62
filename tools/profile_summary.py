"""Summarise an .ncu-rep into a small text file under profiles/ (the raw reports stay in gpurun_out/).
usage: python tools/profile_summary.py gpurun_out/ncu4_decode.ncu-rep profiles/r1_decode_awq.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
    for i, r in enumerate(rows[2:]):
        f.write(f"\n## launch {i}\n")
        for k in hdr:
            if k in KEYS or ("tensor" in k and "avg.pct_of_peak_sustained_active" in k and "sparsity" not in k and "mem_tensor" not in k):
                j = hdr.index(k)
                f.write(f"{k:75s} {r[j]:>18s} {units[j]}\n")
        try:
            rd, wr = float(r[hdr.index('dram__bytes_read.sum')].replace(',', '')), float(r[hdr.index('dram__bytes_write.sum')].replace(',', ''))
            UM = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            # ncu's raw page scales every column by the unit that suits its first row: convert both columns to bytes
            rd *= UM.get(units[hdr.index('dram__bytes_read.sum')], 1.0)
            wr *= UM.get(units[hdr.index('dram__bytes_write.sum')], 1.0)
            mult = 1.0
            dur = float(r[hdr.index('gpu__time_duration.sum')].replace(',', ''))
            du = units[hdr.index('gpu__time_duration.sum')]
            dmult = {"us": 1e-6, "ns": 1e-9, "ms": 1e-3}.get(du, 1e-6)
            f.write(f"{'derived: dram traffic bytes (read+write)':75s} {(rd + wr) * mult:18.0f} B\n")
            f.write(f"{'derived: dram GB/s over the (cold, serialised) ncu duration':75s} {(rd + wr) * mult / (dur * dmult) / 1e9:18.1f} GB/s\n")
        except Exception as e:
            f.write(f"derived: n/a ({e})\n")
print("wrote", out)
