#!/usr/bin/env python
"""Turn an ncu report (--set full) and a launch list (gpu__time_duration) into the summaries committed under profiles/.

    python tools/summarize_ncu.py <report.ncu-rep> <launches.csv> <out_prefix>

Writes <out_prefix>_kernels.md (per-kernel table + stall breakdown + instruction mix) and copies the launch list.
"""
import csv
import io
import shutil
import subprocess
import sys
from collections import Counter, defaultdict

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'TF32 tcgen05 ops, % of peak (elapsed)'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor memory (TMEM) active %'),
    ('sm__inst_executed_pipe_tc.sum', 'tcgen05 MMA instructions'),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'smem wavefronts read by the tensor core'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts LSU (ld + st)'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/smem throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__inst_executed.sum.per_cycle_active', 'IPC (SM)'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__warps_active.avg.per_cycle_active', 'active warps / SM'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM written'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'smem load wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'smem load bank conflicts'),
    ('smsp__sass_inst_executed_op_local_st.sum', 'local (spill) stores'),
]
STALLS = ['selected', 'not_selected', 'long_scoreboard', 'short_scoreboard', 'wait', 'dispatch_stall', 'math_pipe_throttle',
          'barrier', 'no_instruction', 'branch_resolving', 'mio_throttle', 'lg_throttle', 'sleeping']


def ncu(args):
    return subprocess.run(['ncu'] + args, check=True, capture_output=True, text=True).stdout


def main():
    rep, launches, prefix = sys.argv[1:4]
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f'# ncu --set full summary of `{rep.split("/")[-1]}`', '',
           'Captured with `ncu --set full --clock-control none --import-source on` (tools/gpu_profile.sh) on one B200; '
           'numbers under the profiler are cold-cache and serialised, use them for shares and ratios only.', '']
    seen = set()
    for r in data:
        name = r[hdr.index('Kernel Name')].split('(')[0]
        if name in seen:
            continue
        seen.add(name)
        out += [f'## `{name}`', '', '| metric | value |', '|---|---|']
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                out.append(f'| {label} (`{key}`) | {r[i]} {units[i]} |')
        out += ['', 'Warp stall reasons (average warps in the state per issued instruction, `smsp__average_warps_issue_stalled_*_per_issue_active`):', '']
        parts = []
        for s in STALLS:
            key = f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio'
            if key in hdr:
                parts.append(f'{s} {float(r[hdr.index(key)]):.3f}')
        out += [', '.join(parts), '']
        # instruction mix from the source page
        try:
            src = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--kernel-name',
                                                   'regex:' + name.split('<')[0].split('::')[-1].replace('void ', '').strip()]))))
            sh = src[1]
            isrc, iex = sh.index('Source'), sh.index('Instructions Executed')
            mix = Counter()
            for row in src[2:]:
                if row and row[0] == 'Kernel Name':
                    break                                       # first captured instance only
                if len(row) <= iex or not row[iex].isdigit():
                    continue
                t = row[isrc].split()
                if not t:
                    continue
                op = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
                mix[op] += int(row[iex])
            tot = sum(mix.values())
            if tot:
                out += ['Executed warp-instruction mix: ' + ', '.join(f'{op} {100 * n / tot:.1f}%' for op, n in mix.most_common(8)), '']
        except Exception as e:                                  # noqa: BLE001 - best effort
            out += [f'(instruction mix unavailable: {e})', '']
    # launch list
    d = defaultdict(list)
    with open(launches) as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    lh = rows[0]
    ik, iv = lh.index('Kernel Name'), lh.index('Metric Value')
    for r in rows[1:]:
        try:
            d[r[ik].split('(')[0]].append(float(r[iv].replace(',', '')))
        except ValueError:
            pass
    own = {k: v for k, v in d.items() if 'fp32_peak_probe' not in k}
    tot = sum(sum(v) for v in own.values())
    out += ['## Launch list (`gpu__time_duration.sum`, all launches of the bench command, FP32 probe excluded)', '',
            '| kernel | launches | mean us | share |', '|---|---|---|---|']
    for k, v in sorted(own.items(), key=lambda kv: -sum(kv[1])):
        out.append(f'| `{k[:70]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |')
    with open(prefix + '_kernels.md', 'w') as f:
        f.write('\n'.join(out) + '\n')
    shutil.copyfile(launches, prefix + '_launches.csv')
    print('wrote', prefix + '_kernels.md')


if __name__ == '__main__':
    main()
