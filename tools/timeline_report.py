"""Aggregate a tools/timeline.py slots CSV by (demangled) kernel: whole clip, pre-frame part, last frame."""
import collections, csv, re, subprocess, sys
rows = list(csv.DictReader(open(sys.argv[1])))
names = sorted(set(r['kernel'] for r in rows))
dm = subprocess.run(['c++filt'] + [n.replace(';', ',') for n in names], capture_output=True, text=True).stdout.splitlines()
m = dict(zip(names, dm))
def short(n):
    d = m[n].replace('(anonymous namespace)::', '').replace('void ', '').replace('keep::', '')
    return re.sub(r'\(.*$', '', d)
ends = [i for i, r in enumerate(rows) if 'nhwc_to_nchw' in m[r['kernel']]]
def report(title, lo, hi, top=30):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[lo:hi]:
        k = short(r['kernel'])
        if 'conv_tc_kernel' in k:
            g = int(r['grid'].split('x')[0]); k += ' grid=148' if g >= 148 else ' grid<148'
        agg[k][0] += 1; agg[k][1] += float(r['slot_us'])
    tot = sum(v[1] for v in agg.values())
    print("\n== %s: %d kernels, %.1f us" % (title, hi - lo, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-44s n=%5d %9.1f us avg %7.2f %5.1f%%" % (k[:44], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
report("whole clip", 0, len(rows))
report("before the first frame ends (LQ encoder, gains, frame 0)", 0, ends[0] + 1, 14)
report("last frame", ends[-2] + 1, ends[-1] + 1)
