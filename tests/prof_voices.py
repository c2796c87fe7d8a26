"""Developer aid (run via gpurun): render / mix kernel time per call for a C3-shaped script of N voices."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saugns_b200
from saugns_b200 import workloads
for voices in [int(x) for x in sys.argv[1:]] or [4096]:
    prg = workloads.build_c3(voices, 60, seed=1, fm={"mix": "mix", "1": True, "0": False}[os.environ.get("FM", "mix")])
    g = saugns_b200.Generator(prg, 96000, max_call_len=24576, sched=int(os.environ.get("SCHED", "0")))
    for _ in range(3):
        g.run_device(24576)
    g.set_timing(True)
    for _ in range(20):
        g.run_device(24576)
    r, m = g.kernel_ms()
    print(f"{voices} voices: render {r / 20:.3f} ms  mix {m / 20:.3f} ms  -> {voices * 24576 / ((r + m) / 20 * 1e-3) / 1e9:.1f} G voice-samples/s")
    g.close()
