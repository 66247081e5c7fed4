"""Summarises a chrome trace written by `bench.py --trace`: GPU busy time, idle gaps between
consecutive GPU activities and the runtime call the host was in while the GPU idled.
usage: trace_gaps.py trace.json [top-N] [kernels]   -- `kernels`: ignore copies (idle time of the SMs while a
host->device copy is in flight counts as a gap)."""
import json
import sys

ev = json.load(open(sys.argv[1]))["traceEvents"]
cats = ("kernel",) if len(sys.argv) > 3 and sys.argv[3] == "kernels" else ("kernel", "gpu_memcpy", "gpu_memset")
gpu = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in cats), key=lambda e: e["ts"])
cpu = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver")),
             key=lambda e: e["ts"])
t0, t1 = gpu[0]["ts"], max(e["ts"] + e["dur"] for e in gpu)
busy = sum(e["dur"] for e in gpu)
print("GPU activities %d, span %.3f ms, busy %.3f ms, idle %.3f ms" % (len(gpu), (t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3))
gaps = []
end = gpu[0]["ts"] + gpu[0]["dur"]
prev = gpu[0]
for e in gpu[1:]:
    if e["ts"] > end:
        gaps.append((e["ts"] - end, end, prev["name"][:40], e["name"][:40]))
    if e["ts"] + e["dur"] > end:
        end, prev = e["ts"] + e["dur"], e
gaps.sort(reverse=True)
print("gaps > 5 us: %d, total %.3f ms" % (sum(1 for g in gaps if g[0] > 5), sum(g[0] for g in gaps if g[0] > 5) / 1e3))
for g, at, a, b in gaps[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    calls = [c["name"] for c in cpu if c["ts"] < at + g and c["ts"] + c["dur"] > at]
    print("%8.1f us at %9.3f ms  after %-40s before %-40s host: %s" % (g, (at - t0) / 1e3, a, b, ",".join(calls[:6])))
