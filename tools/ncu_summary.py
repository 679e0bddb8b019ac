"""Turn an `ncu --set full` report of mh_lanes_kernel into the markdown summary kept under profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> <sampler.o> <out.md> <title> [<note file>]"""
import csv, subprocess, sys, os
rep, obj, out, title = sys.argv[1:5]
note = open(sys.argv[5]).read() if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
md = f"# {title}\n\n{note}\n| metric | value | unit |\n|---|---|---|\n"
for h, u, v in zip(hdr, units, vals):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h):
        md += f"| {h} | {v} | {u} |\n"
here = os.path.dirname(os.path.abspath(__file__))
lines = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, obj, "mh_lanes", "24"], capture_output=True, text=True).stdout.strip().split("\n")
md += "\nPer-source-line stall attribution (SASS samples joined with nvdisasm line info, tools/ncu_lines.py):\n\n```\n" + "\n".join(l[:170] for l in lines[-25:]) + "\n```\n"
open(out, "w").write(md)
print(md[:1500])
