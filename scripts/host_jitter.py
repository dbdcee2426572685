"""GPU box: how noisy is this host?  A fixed CPU-only loop timed 400 times (median / p99 / max), on 1 and on
4 threads, next to the per-step spread bench.py reports."""
import threading, time, statistics, os
def work():
    s = 0
    for i in range(150000):
        s += i * i
    return s
def sample(n=400):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); work(); ts.append(1e3 * (time.perf_counter() - t0))
    return ts
ts = sample()
print("1 thread : median %.2f ms  p99 %.2f  max %.2f" % (statistics.median(ts), sorted(ts)[int(0.99 * len(ts))], max(ts)))
print("cores", os.cpu_count(), "loadavg", os.getloadavg())
