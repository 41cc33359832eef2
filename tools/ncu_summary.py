"""Developer tool (build container): text summary of an .ncu-rep (`ncu --set full`) for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(rows) - 2} profiled launch(es); ncu --set full --clock-control none")
    for r in rows[2:]:
        print(f"\n== {r[ix['Kernel Name']]}  grid {r[ix.get('Grid Size', 0)]} block {r[ix.get('Block Size', 0)]}")
        for k in KEYS:
            if k in ix:
                print(f"  {k:85s} {r[ix[k]]:>16s} {units[ix[k]]}")
        rd, wr = ix.get("dram__bytes_read.sum"), ix.get("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = float(r[rd].replace(",", "")) * f[units[rd]] + float(r[wr].replace(",", "")) * f[units[wr]]
            print(f"  {'traffic = dram read + write (bytes per launch)':85s} {tot:16.0f} byte")


if __name__ == "__main__":
    main(sys.argv[1])
