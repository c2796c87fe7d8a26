"""Host logic of the multi-GPU drivers (saugns_b200/multigpu.py).

CPU part: partitioning, and the voice-sharded protocol (one sum-reduce of the
float mix planes per call + a 2-integer control all-reduce) driven over a
world-size-2 `gloo` group, with the oracle port standing in for each rank's
GPU shard renderer (test infrastructure; the product shard is CUDA-only).
GPU part (-m gpu): the same protocol with the CUDA shards of one process."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VOICES = [
    "Wsin f220 t0.30 c-0.5 p[Wtri r2 a0.8[g0.1 llin] p[Wsin r3.5 a0.5]]",
    "Wtri f331 t0.21 c0.4 a0.7[g0.2 lxpe]",
    "Wsqr f95 t0.25 a0.3 c0.1 p.f[Wsin f60 a0.4]",
    "Wsin f440 t0.33 p.a0.6 c-0.2",
    "Wpar f200 t0.18 c0.7 a0.5.r1[Wsin f9]",
    "Wsaw f150.r300[Wsin f5] t0.27 c-0.8",
]
HEAD = "S a.m0.2"      # W voices only: N/R seeds depend on the position in the script


def script(voices):
    return "\n".join([HEAD] + list(voices)) + "\n"


def test_shard_scripts_balanced_and_deterministic():
    from saugns_b200 import multigpu as M
    costs = [((i * 7919) % 97 + 1) * 10 for i in range(1000)]
    for world in (1, 2, 4, 8):
        plan = M.shard_scripts(costs, world)
        assert plan == M.shard_scripts(costs, world)
        flat = sorted(i for p in plan for i in p)
        assert flat == list(range(1000))
        loads = [sum(costs[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(costs)
    assert M.shard_scripts([], 4) == [[], [], [], []]
    assert M.shard_scripts([5], 2) == [[0], []]


def test_voice_ranges_cover_in_order():
    from saugns_b200 import multigpu as M
    for vo, world in [(4096, 8), (1024, 3), (5, 8), (0, 2), (7, 1)]:
        r = M.voice_ranges(vo, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == vo
        for a, b in zip(r, r[1:]):
            assert a[1] == b[0]
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1


def test_program_cost_reads_the_program(ref):
    from saugns_b200 import multigpu as M
    a = ref.Program(script(VOICES[:2]))
    b = ref.Program(script(VOICES))
    assert M.program_cost(b) > M.program_cost(a) > 0


class _PortShard:
    """CPU stand-in for one rank's shard: the oracle port renders the rank's
    voices (same ampmult as the whole script) in 1024-frame calls and exposes
    the float mix planes of the block (oracle_mix_buf)."""

    def __init__(self, text):
        import torch
        from oracle import pyref, pyport
        self.torch, self.pyport = torch, pyport
        self.prg = pyref.Program(text)
        self.g = pyport.PortGenerator(self.prg, 48000)
        self.done = False

    def run_mix(self, buf_len):
        import ctypes as C
        assert buf_len <= 1024
        planes = np.zeros(2 * buf_len, np.float32)
        if self.done:
            return False, self.torch.from_numpy(planes), 0
        more, _, n = self.g.run(buf_len)
        L = self.pyport.lib()
        for ch in range(2):
            p = L.oracle_mix_buf(self.g.ptr, ch)
            planes[ch * buf_len: ch * buf_len + n] = np.ctypeslib.as_array(p, shape=(1024,))[:n]
        self.done = not more
        return more, self.torch.from_numpy(planes), n

    def to_pcm(self, planes, buf_len, stereo):
        x = planes.numpy().reshape(2, buf_len)
        x = np.clip(x, -1.0, 1.0).astype(np.float32) * np.float32(32767.0)
        return np.rint(x).astype(np.int16).T.reshape(-1)      # interleaved L,R

    def close(self):
        self.g.close()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from saugns_b200 import multigpu as M
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = M.voice_ranges(len(VOICES), world)[rank]
        shard = _PortShard(script(VOICES[b:e]))
        vg = M.VoiceShardedGenerator(None, 48000, shard=shard)
        pcm = vg.render(1024)
        assert (pcm is None) == (rank != 0)
        q.put((rank, None if pcm is None else pcm.tolist(), vg.ended))
        vg.close()
    finally:
        dist.destroy_process_group()


def test_voice_sharded_protocol_gloo_world2(ref, port):
    """Two ranks, each rendering half of the voices; root's PCM == the oracle's
    render of the whole script within 1 LSB (summation order)."""
    import torch.multiprocessing as mp
    want = port.render(ref.Program(script(VOICES)), srate=48000, call_len=1024)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    mport = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, mport, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, pcm, ended = q.get(timeout=180)
        res[r] = (pcm, ended)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    got = np.array(res[0][0], dtype=np.int16)
    assert got.shape == want.shape
    assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1


@pytest.mark.gpu
def test_voice_sharded_cuda_single_process(ref, port):
    """World size 1 through the real CUDA shard: the driver degenerates to the
    plain generator (bit-exact, no reduce)."""
    import gpuutil
    import saugns_b200
    from saugns_b200 import multigpu as M
    tabs = gpuutil.ref_tables_for_gpu(port)
    prg = ref.Program(script(VOICES))
    want = ref.render(prg, srate=96000)
    vg = M.VoiceShardedGenerator(prg, 96000, device=0, tables=tabs)
    got = vg.render(24576)
    vg.close()
    assert np.array_equal(got, want)
    assert saugns_b200.device_count() >= 1


@pytest.mark.gpu
def test_nccl_voice_and_script_sharding_two_gpus():
    """torchrun x2 over NCCL (tests/mgpu_voice_shard.py); skipped on 1-GPU boxes."""
    import subprocess
    import saugns_b200
    if saugns_b200.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        str(29400 + os.getpid() % 500), os.path.join(ROOT, "tests", "mgpu_voice_shard.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 OK" in r.stdout and "rank 1 OK" in r.stdout
    out = os.path.join(ROOT, "gpurun_out")          # kept as evidence (copied to profiles/ by hand)
    if os.path.isdir(out):
        with open(os.path.join(out, "nccl_voice_shard_x2.log"), "w") as f:
            f.write(r.stdout)
