"""Experiment: do two engines on one GPU (each half of the histories, own stream, driven from two host threads) finish
sooner than one engine with all histories?  The event kernels of one would run under the trace kernels of the other."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skirt9_b200 import abi, configs  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
sim = configs.cfg2(num_packets=N).setup()


def run(engines, blocks, stream_id):
    def work(e, first, count):
        e.clear_instruments()
        e.run_segment(first, count, True, True, False, stream_id)
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(e, f, c)) for e, (f, c) in zip(engines, blocks)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


out = {}
for k in (1, 2, 3):
    os.environ["SK_BANK"] = str((1 << 23) // k)
    engines = [sim.configure(abi.Engine(sim.config_struct())) for _ in range(k)]
    for e in engines:
        e.prepare_primary(N)
    blocks = [(i * N // k, (i + 1) * N // k - i * N // k) for i in range(k)]
    times = [run(engines, blocks, s) for s in range(3)]
    out[f"{k}_engines"] = {"wall_s": times, "pkt_per_s": N / min(times[1:]), "device_ms_each": [e.last_kernel_ms() for e in engines]}
    for e in engines:
        e.close()
print(json.dumps(out))
