"""Summarise an IPDM_OP_TRACE file: python tools/op_trace.py trace.txt [forward_index=-1]
Per-op CUDA-event times of one UNet forward (warm, in stream order) grouped by kind and by layer shape."""
import sys
from collections import defaultdict

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
fwds, cur = [], None
for line in open(path):
    if line.startswith("#"):
        cur = []; fwds.append(cur); continue
    parts = line.split()
    cur.append((parts[0], " ".join(parts[1:-2]), float(parts[-2])))
ops = fwds[which]
total = sum(o[2] for o in ops)
print(f"forward {which}: {len(ops)} ops, {total / 1e3:.2f} ms")
by_kind = defaultdict(float)
for k, _, us in ops:
    by_kind[k] += us
for k, us in sorted(by_kind.items(), key=lambda x: -x[1]):
    print(f"  {k:12s} {us / 1e3:8.2f} ms  {100 * us / total:5.1f}%")
by_shape = defaultdict(lambda: [0, 0.0])
for k, sh, us in ops:
    e = by_shape[(k, sh)]; e[0] += 1; e[1] += us
print("top layers:")
for (k, sh), (n, us) in sorted(by_shape.items(), key=lambda x: -x[1][1])[:40]:
    print(f"  {k:12s} {sh:48s} x{n:<3d} {us:9.1f} us total {us / n:8.1f} us each")
