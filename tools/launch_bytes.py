"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel the launches, average duration, DRAM bytes per launch and the resulting DRAM GB/s."""
import collections, csv, re, sys


def summarise(path, top=45):
    lines = [l for l in open(path) if not l.startswith('==')]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault((r['ID'], re.sub(r'\(.*', '', r['Kernel Name'])), {})
        v, u = float(r['Metric Value'].replace(',', '')), r['Metric Unit']
        if r['Metric Name'] == 'gpu__time_duration.sum':
            d['us'] = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        else:
            d[r['Metric Name']] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    agg = collections.OrderedDict()
    for (_, k), d in by.items():
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get('us', 0.0); a[2] += d.get('dram__bytes_read.sum', 0.0); a[3] += d.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    out = [f'total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches',
           f'{"us":>10} {"share":>6} {"n":>4} {"avg us":>9} {"rd MB":>9} {"wr MB":>9} {"GB/s":>7}  kernel']
    for k, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f'{us:10.1f} {100 * us / tot:5.1f}% {n:4d} {us / n:9.1f} {rd / n / 1e6:9.1f} {wr / n / 1e6:9.1f} {(rd + wr) / us / 1e3:7.0f}  {k[-60:]}')
    return '\n'.join(out)


if __name__ == '__main__':
    for a in sys.argv[1:]:
        print('=====', a)
        print(summarise(a))
