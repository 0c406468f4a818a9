"""profiles/ncu_traffic.json from an ncu launch list (csv with gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch): DRAM bytes per launch of every pipeline kernel of the LAST step in the
list, plus the hash of the CUDA sources the capture was taken from (bench.py marks the figure stale when
the sources change).   python tools/make_traffic_json.py <launches.csv> "<workload name>" [out.json]"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stage_name(kernel):
    k = kernel
    if 'pb_walk_geo_kernel' in k:
        return 's1f_mass' if 'S1FMass' in k else 's1f'
    if 'pb_s32_kernel' in k:
        return 's23_mass' if 'PbS32Mass' in k else 's23'
    if 'pb_fields_row_kernel' in k:
        return 'k2_fields'
    if 'pb_basis_batch_kernel' in k:
        return 'k1_tables'
    if 'pb_lane_span_kernel' in k:
        return 's3'
    if 'pb_walk_kernel' in k:
        return 'walk'
    return None


def main():
    path, workload = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hi]
    ik, im, iv, iid = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
    launches = {}
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        d = launches.setdefault(int(r[iid]), {'kernel': r[ik]})
        d[r[im]] = float(r[iv].replace(',', ''))
    order = sorted(launches)
    # the last occurrence of every stage = the last step
    last = {}
    for i in order:
        st = stage_name(launches[i]['kernel'])
        if st:
            last[st] = launches[i]
    unit = 1.0
    from bench import source_hash
    res = {'workload': workload,
           'source': '%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, 1 GPU)' % os.path.basename(path),
           'source_hash': source_hash(),
           'dram_bytes_per_launch': {k: (v.get('dram__bytes_read.sum', 0.0) + v.get('dram__bytes_write.sum', 0.0)) * unit for k, v in last.items()},
           'dram_bytes_read_per_launch': {k: v.get('dram__bytes_read.sum', 0.0) for k, v in last.items()},
           'dram_bytes_write_per_launch': {k: v.get('dram__bytes_write.sum', 0.0) for k, v in last.items()},
           'ms_per_launch_under_ncu': {k: v.get('gpu__time_duration.sum', 0.0) * 1e-6 for k, v in last.items()}}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
