import csv, sys
for v in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f'gpurun_out/rounds_{v}.csv')) if len(r) > 5]
    hdr = rows[0]; iid = hdr.index('ID'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value')
    d = {}
    for r in rows[1:]:
        d.setdefault(r[iid], {})[r[im]] = float(r[iv].replace(',', ''))
    tot = sum(m['gpu__time_duration.sum'] for m in d.values()) / 1e3
    print(v, 'total us %.1f' % tot)
    for k in sorted(d, key=int):
        m = d[k]
        print(' ', k, 'us %.1f' % (m['gpu__time_duration.sum'] / 1e3), 'inst %.3e' % m['smsp__inst_executed.sum'],
              'issue %.1f' % m['smsp__issue_active.avg.pct_of_peak_sustained_active'], 'dramMB %.0f' % (m['dram__bytes_read.sum'] / 1e6))
