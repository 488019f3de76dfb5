"""Summarise an .ncu-rep (raw page + source page hot spots) into text; used for profiles/."""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('== kernel:', r[idx['Kernel Name']][:110])
        for w in WANT:
            if w in idx:
                print(f'  {w} = {r[idx[w]]} {units[idx[w]]}')
    src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv']))))
    cur = None
    kernels = []
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = dict(name=r[1], rows=[])
            kernels.append(cur)
        elif r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and len(r) > 5:
            cur['rows'].append(r)
    for k in kernels:
        h = k['hdr']
        si, so, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
        tot = sum(int(r[si]) for r in k['rows']) or 1
        print('== hot instructions (share of warp-stall samples):', k['name'][:100])
        for r in sorted(k['rows'], key=lambda r: -int(r[si]))[:top]:
            print(f'  {int(r[si]) * 100 / tot:5.1f}%  exec={r[ie]:>9}  {r[so].strip()[:100]}')


if __name__ == '__main__':
    main()
