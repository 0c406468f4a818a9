"""Aggregate the warp-stall samples of an .ncu-rep per CUDA source line (needs -lineinfo and
--import-source on):  python tools/ncu_lines.py report.ncu-rep [top] [kernel-name substring]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    want = sys.argv[3] if len(sys.argv) > 3 else ''
    func = ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    cur, hdr, acc = None, None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            func = r[1]
        elif r[0] == 'Line No':
            hdr = r
        elif hdr and r[0] != '' and len(r) > 5 and want in func:
            try:
                acc.append((float(r[4]), cur, r[0], r[1][:120]))
            except ValueError:
                pass
    tot = sum(a[0] for a in acc) or 1.0
    for s, f, ln, src in sorted(acc, reverse=True)[:top]:
        print('%5.1f%% %s:%s  %s' % (100 * s / tot, f, ln, src))


if __name__ == '__main__':
    main()
