"""Summarise an .ncu-rep (run where ncu is installed; no GPU needed): key metrics per kernel + hottest source lines.
usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [kernel-regex-for-source]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_dispatch_stall',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_drain',
        'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_warps_issue_stalled_sleeping',
        'smsp__pcsamp_warps_issue_stalled_tex_throttle', 'smsp__pcsamp_warps_issue_stalled_imc_miss',
        'smsp__pcsamp_warps_issue_stalled_misc']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print('== kernel:', name[:90])
    for w in WANT:
        if w in hdr:
            print('   %-70s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name',
                          'regex:' + sys.argv[2], '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(r for r in rows if 'Instructions Executed' in r)
    iE, iS = hdr.index('Instructions Executed'), hdr.index('# Samples')
    out = [(int(r[iE]), int(r[iS]) if r[iS].isdigit() else 0, int(r[0]), r[1].strip()[:110]) for r in rows
           if len(r) > iE and r[0].isdigit() and r[iE].isdigit()]
    tot = sum(o[0] for o in out)
    stot = sum(o[1] for o in out)
    print('== hottest source lines of', sys.argv[2], '(instructions executed %, stall samples %)')
    for e, s, l, text in sorted(out, key=lambda t: -t[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        print('  %5.2f%% %5.2f%% L%4d %s' % (100.0 * e / tot, 100.0 * s / max(stot, 1), l, text))
