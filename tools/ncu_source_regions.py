"""Per-region stall summary of an `ncu --page source --csv` dump: python tools/ncu_source_regions.py <csv> [bucket]
Groups consecutive SASS instructions into buckets of `bucket` instructions and prints the samples, executed
instructions and top stall reasons of the busiest ones, with the first/last opcodes of each bucket."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 100
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
tot_inst = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
print("total samples %d, warp instructions executed %d" % (tot, tot_inst))
for b in range(0, len(data), bucket):
    blk = data[b:b + bucket]
    s = sum(int(r[ix['# Samples']] or 0) for r in blk)
    n = sum(int(r[ix['Instructions Executed']] or 0) for r in blk)
    if s < tot * 0.004:
        continue
    st = sorted(((sum(int(r[ix[c]] or 0) for r in blk), c) for c in stall_cols), reverse=True)[:4]
    ops = {}
    for r in blk:
        op = r[ix['Source']].split()[0] if r[ix['Source']].split() else ''
        if op.startswith('@'):
            op = r[ix['Source']].split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top_ops = ' '.join('%s:%d' % (k, v) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    print("%5d-%5d  samples %6d (%4.1f%%)  inst %10d (%4.1f%%)  %s | %s" % (
        b, b + len(blk), s, 100.0 * s / tot, n, 100.0 * n / max(tot_inst, 1),
        ' '.join('%s=%d' % (c[6:], v) for v, c in st if v), top_ops))
