"""ncu .ncu-rep -> small CSV of the metrics the roofline discussion uses (profiles/*.ncu_summary.csv).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum']


def main(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u, body = rows[0], rows[1], rows[2:]
    with open(out + '.ncu_summary.csv', 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'metric', 'value', 'unit'])
        for r in body:
            for k in KEYS:
                if k in h:
                    w.writerow([r[h.index('Kernel Name')], k, r[h.index(k)], u[h.index(k)]])
    print('wrote', out + '.ncu_summary.csv')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
