"""Extracts the metrics quoted in profiles/*.txt from an ncu report (`ncu --set full --import-source on ... -o X`).

  python tools/ncu_extract.py raw   X.ncu-rep [metric-substring ...]   # one line per metric, one column per captured kernel
  python tools/ncu_extract.py stall X.ncu-rep [top_n]                   # warp-stall reasons and the hottest SASS instructions

Reads the report with `ncu -i ... --page raw|source --csv` (works without a GPU)."""
import csv
import subprocess
import sys

DEFAULT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'inst_executed_pipe_fp64', 'pipe_tensor', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
           'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_sectors_op_read.sum', 'sm__throughput.avg.pct',
           'l1tex__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'stalled_long_scoreboard_per_warp_active', 'stalled_short_scoreboard_per_warp_active',
           'stalled_lg_throttle_per_warp', 'stalled_mio_throttle_per_warp', 'stalled_wait_per_warp', 'stalled_barrier_per_warp', 'stalled_math_pipe',
           'issue_active.avg.pct', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit', 'launch__waves', 'local_op', 'smsp__cycles_active.avg',
           'sm__cycles_elapsed.max', 'pipe_fp64_cycles_active']
REASONS = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_drain', 'stall_lg', 'stall_long_sb', 'stall_math', 'stall_membar', 'stall_mio',
           'stall_misc', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_sleep', 'stall_tex', 'stall_wait']


def page(rep, which):
    out = subprocess.run(['ncu', '-i', rep, '--page', which, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def raw(rep, want):
    r = page(rep, 'raw')
    hdr = r[0]
    print('kernels:', [row[hdr.index('Kernel Name')][:30] for row in r[2:]])
    for i, h in enumerate(hdr):
        if any(w in h for w in want):
            print(f"{h:85s}", [row[i] for row in r[2:]])


def stall(rep, top_n):
    rows = page(rep, 'source')
    start = next(i for i, row in enumerate(rows) if row and row[0] == 'Address')
    print(rows[start - 1][:2] if start else '')
    hdr, data = rows[start], [d for d in rows[start + 1:] if len(d) > 4]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda d, k: int(d[ix[k]] or 0) if k in ix and d[ix[k]].strip().isdigit() else 0
    total = sum(num(d, '# Samples') for d in data)
    print('samples', total)
    for name, v in sorted(((r, sum(num(d, r) for d in data)) for r in REASONS), key=lambda x: -x[1]):
        if v:
            print(f'  {name:24s}{v:7d} {100.0 * v / max(total, 1):5.1f}%')
    print('hottest instructions:')
    for k in sorted(range(len(data)), key=lambda k: -num(data[k], '# Samples'))[:top_n]:
        d = data[k]
        why = ' '.join(f'{r[6:]}={num(d, r)}' for r in REASONS if num(d, r))
        print(f'  #{k:5d} {num(d, "# Samples"):5d}  {d[ix["Source"]].strip()[:64]:64s} {why}')


if __name__ == '__main__':
    if len(sys.argv) < 3 or sys.argv[1] not in ('raw', 'stall'):
        sys.exit(__doc__)
    if sys.argv[1] == 'raw':
        raw(sys.argv[2], sys.argv[3:] or DEFAULT)
    else:
        stall(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
