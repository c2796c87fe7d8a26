"""Developer aid (run via gpurun): where the wall time of the C5 batch path goes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import saugns_b200
from saugns_b200 import workloads, batch, generator as G
from oracle import pyref, pyport

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
group = int(sys.argv[2]) if len(sys.argv) > 2 else 256
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1
use_sink = len(sys.argv) > 4 and sys.argv[4] == "sink"
call_len = int(sys.argv[5]) if len(sys.argv) > 5 else None
t = pyport.ref_tables()
tabs = saugns_b200.WaveTables.from_buffer_copy(bytes(t))
tabs._keep = t
texts = [workloads.synth_c5_script(i) for i in range(n)]
prgs = [pyref.Program(x) for x in texts]
batch.render_batch(prgs[:16], srate=96000, tables=tabs, group_size=16)

# instrument
acc = {"create": 0.0, "run_many": 0.0, "close": 0.0, "render_ms": 0.0, "mix_ms": 0.0, "calls": 0}
_Gen = G.Generator
_init, _close = _Gen.__init__, _Gen.close


def init(self, *a, **k):
    t0 = time.perf_counter()
    _init(self, *a, **k)
    self.set_timing(True)
    acc["create"] += time.perf_counter() - t0


def close(self):
    if getattr(self, "ptr", None):
        r, m = self.kernel_ms()
        acc["render_ms"] += r
        acc["mix_ms"] += m
    t0 = time.perf_counter()
    _close(self)
    acc["close"] += time.perf_counter() - t0


_Gen.__init__, _Gen.close = init, close
L = G.lib()
_rm = L.saugen_batch_end


class Wrap:
    def __getattr__(self, k):
        return getattr(L, k)

    def saugen_batch_end(self, *a):
        t0 = time.perf_counter()
        r = _rm(*a)
        acc["run_many"] += time.perf_counter() - t0
        acc["calls"] += 1
        return r


G.lib = lambda: Wrap()
t0 = time.perf_counter()
out = batch.render_batch(prgs, srate=96000, tables=tabs, group_size=group, threads=threads, call_len=call_len,
                         sink=(lambda i, pcm: None) if use_sink else None)
wall = time.perf_counter() - t0
print(f"scripts {n} group {group} threads {threads} sink {use_sink} call_len {call_len} wall {wall:.3f}s -> {n / wall:.1f} scripts/s")
for k, v in acc.items():
    print(f"  {k:10s} {v:.3f}")
print("  other (python)", wall - acc["create"] - acc["run_many"] - acc["close"])
