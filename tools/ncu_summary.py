"""Summarise an .ncu-rep (raw page + source page) into a short text: per kernel duration, DRAM
bytes, throughputs, occupancy and the top stall instructions."""
import collections
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


def source(rep, launch=None):
    cmd = ['ncu', '-i', rep, '--page', 'source', '--csv']
    if launch is not None:
        cmd += ['--launch-skip', str(launch), '--launch-count', '1']
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    sect, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            sect.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    return sect


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum']


def main():
    rep = sys.argv[1]
    h, rows = raw(rep)
    for n, r in enumerate(rows):
        src = [None] * n + source(rep, n)[:1]
        print('=' * 100)
        print(r[h.index('Kernel Name')][:95], '| grid', r[h.index('Grid Size')], 'block', r[h.index('Block Size')])
        for k in KEYS:
            if k in h:
                print('  %-72s %s' % (k, r[h.index(k)]))
        st = [(float(r[i]), k.replace('smsp__pcsamp_warps_issue_stalled_', '')) for i, k in enumerate(h)
              if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and r[i] not in ('', 'n/a')]
        tot = sum(v for v, _ in st) or 1
        print('  stalls: ' + ', '.join('%s %.0f%%' % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:6]))
        if n < len(src):
            s = src[n]
            hdr, data = s['rows'][0], s['rows'][1:]
            ia, isrc = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source')
            tot = sum(int(x[ia]) for x in data if len(x) > ia and x[ia].isdigit()) or 1
            top = sorted([(int(x[ia]), i, x[isrc]) for i, x in enumerate(data) if len(x) > ia and x[ia].isdigit()], reverse=True)[:8]
            for v, i, t in top:
                print('    %5.1f%% #%-5d %s' % (100 * v / tot, i, t.strip()[:80]))
            ops = collections.Counter()
            for x in data:
                if len(x) > ia and x[ia].isdigit():
                    tk = x[isrc].split()
                    ops[tk[1] if tk and tk[0].startswith('@') else (tk[0] if tk else '?')] += int(x[ia])
            print('    by opcode: ' + ', '.join('%s %.0f%%' % (k, 100 * v / tot) for k, v in ops.most_common(8)))


if __name__ == '__main__':
    main()
