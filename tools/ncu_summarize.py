"""Turns ncu exports into the small tracked files under profiles/ (run here, no GPU needed):

  python tools/ncu_summarize.py launches gpurun_out/r02_launches_raw.csv profiles/r02_launches_by_kernel.csv
  ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_summarize.py full /tmp/raw.csv profiles/r02_ncu_full_summary.csv profiles/r02_ncu_dram_bytes_per_launch.json
"""
import csv
import json
import re
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_max_active', 'sm__inst_executed.sum', 'smsp__cycles_active.avg']


def short(name):
    name = re.sub(r'\(.*$', '', name).strip()
    return re.sub(r'^(void\s+)?(kb::)?', '', name)


def rows_of(path):
    with open(path, newline='') as fh:
        lines = [ln for ln in fh if not ln.startswith('==')]
    return list(csv.reader(lines))


def launches(src, dst):
    rows = rows_of(src)
    hdr = rows[0]
    ik, im, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    iu = hdr.index('Metric Unit')
    tot, cnt = {}, {}
    for r in rows[1:]:
        if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
            continue
        v = float(r[iv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[iu], 1e-3)
        k = short(r[ik])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] = cnt.get(k, 0) + 1
    s = sum(tot.values())
    with open(dst, 'w', newline='') as fh:
        w = csv.writer(fh)
        w.writerow(['kernel', 'launches', 'total_us', 'avg_us', 'share'])
        for k in sorted(tot, key=lambda k: -tot[k]):
            w.writerow([k, cnt[k], round(tot[k], 2), round(tot[k] / cnt[k], 2), round(tot[k] / s, 4)])
    print(open(dst).read())


def full(src, dst_csv, dst_json):
    rows = rows_of(src)
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    cols = [(c, hdr.index(c)) for c in KEEP if c in hdr]
    out, dram = [], {}
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        k = short(r[ik])
        rec = {'kernel': k}
        for c, i in cols:
            rec[c + (' [' + units[i] + ']' if units[i] else '')] = r[i]
        out.append(rec)
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')].replace(',', '')) * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0}.get(units[hdr.index('dram__bytes_read.sum')], 1.0)
            wr = float(r[hdr.index('dram__bytes_write.sum')].replace(',', '')) * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0}.get(units[hdr.index('dram__bytes_write.sum')], 1.0)
            dram.setdefault(k, []).append(rd + wr)
        except (ValueError, IndexError):
            pass
    keys = ['kernel'] + [k for k in out[0] if k != 'kernel'] if out else ['kernel']
    with open(dst_csv, 'w', newline='') as fh:
        w = csv.DictWriter(fh, fieldnames=keys)
        w.writeheader()
        for rec in out:
            w.writerow(rec)
    with open(dst_json, 'w') as fh:
        json.dump(dram, fh, indent=1)
    print(open(dst_csv).read())
    print(json.dumps(dram, indent=1))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
