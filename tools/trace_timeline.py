"""Per-SM occupancy timeline of one solve from a -DSBQ_TRACE build (CTAs print smid, start and end globaltimer).

   SBQ_NVCC_EXTRA=-DSBQ_TRACE python strawberry_b200/build.py --force; cp strawberry_b200/libsbq.so build/variants/libsbq_trace.so
   SBQ_LIB_PATH=build/variants/libsbq_trace.so python tools/prof.py human 20000 2>&1 | python tools/trace_timeline.py
"""
import collections
import re
import sys

recs = []
for line in sys.stdin:
    m = re.match(r"TRACE (\S+) nt(\d+) locus (-?\d+) rank (\d+) sm (\d+) t0 (\d+) t1 (\d+) iters (\d+)", line)
    if m:
        recs.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(6)), int(m.group(7)), int(m.group(8))))
if not recs:
    sys.exit("no TRACE lines")
# the last solve only: records arrive in bursts, split on gaps > 2 ms
recs.sort(key=lambda r: r[5])
runs, cur = [], [recs[0]]
for r in recs[1:]:
    if r[5] - max(x[6] for x in cur[-50:]) > 2_000_000:
        runs.append(cur)
        cur = []
    cur.append(r)
runs.append(cur)
run = runs[-1]
t0 = min(r[5] for r in run)
t1 = max(r[6] for r in run)
print(f"{len(runs)} solves traced; last: {len(run)} CTAs, span {(t1 - t0) / 1e6:.3f} ms")
by_cls = collections.defaultdict(lambda: [0, 0.0, 1e18, 0.0])
for cls, nt, l, rank, sm, a, b, it in run:
    x = by_cls[(cls, nt)]
    x[0] += 1
    x[1] += (b - a) / 1e6
    x[2] = min(x[2], (a - t0) / 1e6)
    x[3] = max(x[3], (b - t0) / 1e6)
tot = 0.0
for k, (n, sm_ms, first, last) in sorted(by_cls.items()):
    print(f"  {k[0]:5s} nt{k[1]:4d}: {n:6d} CTAs, {sm_ms:9.2f} CTA-ms, first start {first:6.3f} ms, last end {last:6.3f} ms")
    tot += sm_ms
sms = sorted({r[4] for r in run})
print(f"  SMs seen: {len(sms)}; sum of CTA-ms {tot:.1f}")
# busy fraction per 0.25 ms bucket: an SM is busy if any CTA is resident
nb = int((t1 - t0) / 250_000) + 1
busy = [set() for _ in range(nb)]
heavy = [0.0] * nb
for cls, nt, l, rank, sm, a, b, it in run:
    for k in range(int((a - t0) / 250_000), int((b - t0) / 250_000) + 1):
        busy[k].add(sm)
print("  SMs with a resident CTA per 0.25 ms:", [len(b) for b in busy])
longest = sorted(run, key=lambda r: r[6] - r[5], reverse=True)[:8]
for cls, nt, l, rank, sm, a, b, it in longest:
    print(f"  longest: {cls} locus {l} rank {rank} sm {sm} start {(a - t0) / 1e6:.3f} end {(b - t0) / 1e6:.3f} iters {it}")
