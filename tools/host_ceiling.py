"""What can the HOST side of this box feed?  Pure pinned-memory cudaMemcpyAsync traffic, no kernels: per device one
thread moves the byte mix of the end-to-end polymul leg (2 bytes H2D for every byte D2H) through three streams, for
1, 2, 4, ... all visible GPUs at once.  The aggregate GB/s is the ceiling the end-to-end number of bench.py is
measured against.
usage: python tools/host_ceiling.py [--mb 512] [--secs 2.0] [--json out.json]"""
import argparse
import json
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=512)
ap.add_argument("--secs", type=float, default=2.0)
ap.add_argument("--json", default=None)
args = ap.parse_args()

ndev = torch.cuda.device_count()
chunk = 24 << 20


def worker(dev, stop, counts, barrier):
    torch.cuda.set_device(dev)
    hin = torch.empty(args.mb << 20, dtype=torch.uint8).pin_memory()
    hout = torch.empty((args.mb << 20) // 2, dtype=torch.uint8).pin_memory()
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    dins = [torch.empty(chunk, dtype=torch.uint8, device="cuda:%d" % dev) for _ in range(3)]
    douts = [torch.empty(chunk // 2, dtype=torch.uint8, device="cuda:%d" % dev) for _ in range(3)]
    barrier.wait()
    moved = 0
    i = 0
    nchunks = (args.mb << 20) // chunk
    while not stop.is_set():
        s = i % 3
        off = (i % nchunks) * chunk
        with torch.cuda.stream(streams[s]):
            dins[s].copy_(hin[off:off + chunk], non_blocking=True)
            hout[off // 2:off // 2 + chunk // 2].copy_(douts[s], non_blocking=True)
        if i % 3 == 2:
            for st in streams:
                st.synchronize()
            moved += 3 * (chunk + chunk // 2)
            counts[dev] = moved
        i += 1
    for st in streams:
        st.synchronize()


res = {}
n = 1
while n <= ndev:
    stop = threading.Event()
    counts = [0] * ndev
    barrier = threading.Barrier(n + 1)
    th = [threading.Thread(target=worker, args=(d, stop, counts, barrier)) for d in range(n)]
    for t in th:
        t.start()
    barrier.wait()
    time.sleep(0.3)
    c0, t0 = sum(counts), time.perf_counter()
    time.sleep(args.secs)
    c1, t1 = sum(counts), time.perf_counter()
    stop.set()
    for t in th:
        t.join()
    gbs = (c1 - c0) / (t1 - t0) / 1e9
    res[str(n)] = {"aggregate_GBps": gbs, "per_gpu_GBps": gbs / n, "polymul_per_s_equivalent_n512": gbs * 1e9 / 6144}
    print("%d GPU(s): %.1f GB/s aggregate (H2D + D2H, 2:1), %.1f per GPU  -> ceiling %.3g polymul/s at 6144 B per product" % (n, gbs, gbs / n, gbs * 1e9 / 6144))
    n *= 2
if args.json:
    json.dump(res, open(args.json, "w"), indent=1)
