#!/usr/bin/env python
"""Condense an `ncu --set full` report (.ncu-rep, read here without a GPU) into the handful of numbers the roofline
discussion needs, as markdown:   python scripts/ncu_summary.py gpurun_out/x.ncu-rep [...] > profiles/r02_x_ncu_full.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % (gpu)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smem wavefronts read by the tensor core"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem tensor-core wavefronts % of peak"),
    ("l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_red.sum", "TMA reduce-add bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM fabric read bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts (all)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem bank conflicts (ld)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smem bank conflicts (st)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    units = rd[1]
    return hdr, units, rd[2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, rows = rows_of(path)
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            print(f"### `{r[col['Kernel Name']]}`  ({path.split('/')[-1]}, ncu --set full --clock-control none)\n")
            print("| metric | value | unit |\n|---|---|---|")
            for key, label in WANT:
                if key in col:
                    print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
            print()


if __name__ == "__main__":
    main()
