"""Summarise an .ncu-rep (raw metrics + SASS opcode mix + stall reasons) -- run on the CPU box."""
import collections
import csv
import json
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']


def main(rep, n_steps=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = {'report': rep, 'kernels': []}
    for r in rows[2:]:
        d = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = '%s %s' % (r[i], units[i])
        out['kernels'].append(d)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if start:
        h = rows[start[0]]
        end = start[1] - 1 if len(start) > 1 else len(rows)
        iA, iE = h.index('Source'), h.index('Instructions Executed')
        ops, tot = collections.Counter(), 0
        stalls = collections.Counter()
        scol = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
        for r in rows[start[0] + 1:end]:
            try:
                n = int(r[iE])
            except (ValueError, IndexError):
                continue
            toks = r[iA].split()
            op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
            ops[op] += n
            tot += n
            for i, nm in scol:
                try:
                    stalls[nm] += int(r[i])
                except ValueError:
                    pass
        out['warp_instructions'] = tot
        out['opcode_mix'] = {k: v for k, v in ops.most_common(30)}
        out['stall_samples'] = {k: v for k, v in stalls.most_common(12)}
        if n_steps:
            out['warp_instructions_per_warp_step'] = {k: round(v / n_steps, 2) for k, v in ops.most_common(30)}
            out['total_per_warp_step'] = round(tot / n_steps, 1)
    return out


if __name__ == '__main__':
    rep = sys.argv[1]
    n = float(sys.argv[2]) if len(sys.argv) > 2 else None
    print(json.dumps(main(rep, n), indent=1))
