#!/usr/bin/env python
"""Digest an .ncu-rep (read in the build container): key raw metrics and the
per-region instruction / stall-sample distribution from the source page.

    python tools/ncu_digest.py gpurun_out/prof.ncu-rep [--md out.md]
"""
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    lines = []
    rows = page(rep, 'raw')
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines += ['| metric | unit | value |', '|---|---|---|']
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            lines.append(f'| {h} | {u} | {v} |')
    rows = page(rep, 'source')
    hdr = rows[1]
    ia, isrc = hdr.index('Address'), hdr.index('Source')
    iex, ith = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    isamp = hdr.index('# Samples')
    data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ith]), int(r[isamp]))
            for r in rows[2:] if len(r) > iex]
    base = data[0][0]
    tot = sum(d[2] for d in data)
    tots = sum(d[4] for d in data)
    lines += ['', f'SASS instructions: {len(data)}; executed warp-instructions: {tot}; '
              f'stall samples: {tots}', '',
              '| SASS offset | % inst | % samples | active lanes | top opcodes (% of all inst) |',
              '|---|---|---|---|---|']
    blk = 48
    for k in range(0, len(data), blk):
        seg = data[k:k + blk]
        ex = sum(d[2] for d in seg)
        sm = sum(d[4] for d in seg)
        th = sum(d[3] for d in seg)
        if ex / tot < 0.004 and sm / tots < 0.004:
            continue
        ops = {}
        for d in seg:
            t = d[1].split()
            op = (t[0] if not t[0].startswith('@') else t[1]).split('.')[0]
            ops[op] = ops.get(op, 0) + d[2]
        top = sorted(ops.items(), key=lambda x: -x[1])[:6]
        lines.append(f'| {seg[0][0] - base:#x} | {100 * ex / tot:.1f} | {100 * sm / tots:.1f} | '
                     f'{th / max(ex, 1):.1f} | ' + ' '.join(f'{o}:{100 * c / tot:.1f}' for o, c in top) + ' |')
    text = '\n'.join(lines) + '\n'
    if '--md' in sys.argv:
        open(sys.argv[sys.argv.index('--md') + 1], 'w').write(text)
    else:
        print(text)


if __name__ == '__main__':
    main()
