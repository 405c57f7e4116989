"""Per-kernel critical-path view of one decode layer from a bench.py --timeline file."""
import sys
ev = []
for l in open(sys.argv[1]):
    if l.startswith('#'):
        continue
    a = l.split(' ', 2)
    ev.append((float(a[0]), float(a[1]), a[2].strip()))
idx = [i for i, e in enumerate(ev) if 'rope_table_kernel' in e[2]]
print("step span ms", round((ev[-1][0] + ev[-1][1]) / 1e3, 2))
print("forward spans us:", [round(ev[idx[k + 1]][0] - ev[idx[k]][0], 1) for k in range(len(idx) - 1)])
s = idx[6]
def short(n):
    return n.replace('isst::', '').replace('void ', '').replace('tc::', '')[:44]
prev_end = None
for i in range(s + 2 + 8 * 3, s + 2 + 8 * 5 + 1):
    st, d, n = ev[i]
    end = st + d
    print(f"{st - ev[s][0]:9.2f} dur {d:7.2f} end {end - ev[s][0]:8.2f} crit {'' if prev_end is None else round(end - prev_end, 2):>6}  {short(n)}")
    prev_end = end
