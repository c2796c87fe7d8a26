"""Multi-GPU drivers for the generator back end: one process per GPU
(`torchrun`), `torch.distributed` for the plumbing.

The path shards in two ways, both with the reference's semantics intact:

* script sharding -- independent programs are dealt to ranks; voices of
  different scripts never meet, so there is NO data-path collective
  (`shard_scripts`, `render_scripts`).
* voice sharding -- one large program: each rank renders a contiguous range of
  voice indices (voices interact only in mix_add, sau/generator.c:863-869) and
  the per-call float mix planes (2 x buf_len floats) are summed with ONE
  reduce to the root, which clips and converts to int16
  (`VoiceShardedGenerator`).  Every rank keeps the global `vo_count` (for
  amp_scale, generator.c:183-185) and the global event timeline, so segment
  boundaries are identical everywhere.  The summation order differs from the
  reference's serial voice loop => PCM may differ by 1 LSB (allowed).

Nothing here computes audio: rendering is `saugns_b200.Generator` (CUDA only).
The host logic is backend-agnostic so that the world-size-2 `gloo` tests can
drive it on CPU with a stand-in shard renderer.
"""
import ctypes as C

import numpy as np

from . import program as P


# ---------------------------------------------------------------------------
# partitioning (pure host logic)
# ---------------------------------------------------------------------------
def program_cost(prg):
    """Rendering cost estimate of a sauProgram: operators x duration (ms).
    `prg` has `.ptr` (address of a sauProgram, include/sau_program_abi.h)."""
    p = P.Program.from_address(prg.ptr)
    return max(1, int(p.op_count)) * max(1, int(p.duration_ms))


def shard_scripts(costs, world):
    """Greedy longest-processing-time assignment of scripts to ranks.
    -> list (per rank) of script indices, each list in ascending order.
    Deterministic: every rank computes the same plan without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    plan = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += costs[i]
    return [sorted(p) for p in plan]


def voice_ranges(vo_count, world):
    """Contiguous voice-index ranges [begin, end) per rank, sizes differing by <= 1."""
    base, extra = divmod(vo_count, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


# ---------------------------------------------------------------------------
# script sharding
# ---------------------------------------------------------------------------
def render_scripts(programs, srate=96000, rank=0, world=1, device=0, call_len=None, tables=None,
                   group_size=256, on_done=None):
    """Render this rank's share of `programs` (all ranks pass the same list).
    -> {script index: int16 array [frames, 2]}.  No collective is involved."""
    from . import batch
    plan = shard_scripts([program_cost(p) for p in programs], world)[rank]
    pcm = batch.render_batch([programs[i] for i in plan], srate=srate, device=device,
                             call_len=call_len, tables=tables, group_size=group_size)
    out = dict(zip(plan, pcm))
    if on_done is not None:
        for i in plan:
            on_done(i, out[i])
    return out


# ---------------------------------------------------------------------------
# voice sharding
# ---------------------------------------------------------------------------
MIX_TAIL = 64        # SAUGEN_MIX_TAIL (include/saugen_b200.h): floats after the planes for control words


def voice_shards(prg, srate, world):
    """Contiguous voice ranges per rank that keep together the voices an operator is handed
    over between (saugen_voice_groups: parseconv re-homes an operator whose voice slot was
    recycled; its state lives on one generator).  Boundaries of the even split move up past
    the end of any group they would cut; a script that is one big group lands on rank 0."""
    from .generator import voice_groups
    vo_count = P.Program.from_address(prg.ptr).vo_count
    ranges = voice_ranges(vo_count, world)
    n, groups = voice_groups(prg, srate)
    if n == 0:
        return ranges
    last = {}
    for v, g in enumerate(groups):
        last[g] = v
    cuts = [b for b, _ in ranges[1:]]
    out, begin = [], 0
    for c in cuts:
        c = max(c, begin)
        moved = True
        while moved:                      # no group may straddle the cut
            moved = False
            for v in range(begin, min(c, vo_count)):
                if last[groups[v]] >= c:
                    c = last[groups[v]] + 1
                    moved = True
        c = min(c, vo_count)
        out.append((begin, c))
        begin = c
    out.append((begin, vo_count))
    return out


class _CudaShard:
    """This rank's voices of the program on its GPU."""

    def __init__(self, prg, srate, voice_range, device, max_call_len, tables):
        from .generator import Generator
        import torch
        self.torch = torch
        self.device = device
        self.gen = Generator(prg, srate, tables=tables, device=device, voice_range=voice_range,
                             max_call_len=max_call_len)

    def run_mix(self, buf_len):
        """-> (more, tensor on the GPU: the float planes + MIX_TAIL spare floats, out_len)."""
        from .generator import planes_as_torch
        more, ptr, n = self.gen.run_mix(buf_len)
        row = self.row_len
        t = planes_as_torch(ptr, 2 * row + MIX_TAIL)
        # L plane at [0, buf_len), R plane at [row_len, row_len + buf_len)
        return more, t, n

    @property
    def row_len(self):
        return self._row_len

    @property
    def tail_offset(self):
        return 2 * self._row_len

    def set_row_len(self, n):
        self._row_len = n

    def to_pcm(self, planes, buf_len, stereo):
        return self.gen.mix_to_pcm(planes.data_ptr(), buf_len, stereo)

    def close(self):
        self.gen.close()


class VoiceShardedGenerator:
    """sau_create_Generator / sauGenerator_run for ONE program whose voices are
    spread over the ranks of a process group.

    run() must be called by every rank; rank `root` gets the PCM, the others
    get None.  All ranks get the same (more, out_len).  Each call costs exactly ONE
    collective: an all-reduce (sum) of a buffer holding the rank's float L / R planes
    followed by 2 x world control words -- slot r of the first `world` holds rank r's
    "more signal follows", slot r of the second its out_len, zeros elsewhere -- so the
    same reduction that mixes the voices tells every rank whether any shard is still
    alive and how long the longest one ran (`collectives` counts them).

    `shard` (optional) is this rank's renderer: an object with
    run_mix(buf_len) -> (more, planes, out_len), to_pcm(planes, buf_len,
    stereo) -> int16 array, close().  Default: the CUDA generator on
    `device`.  Tests pass a CPU stand-in to exercise the protocol over gloo.
    """

    def __init__(self, prg, srate=96000, group=None, root=0, device=None, max_call_len=0,
                 tables=None, shard=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.torch = torch
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if 2 * self.world > MIX_TAIL:
            raise ValueError("voice sharding over more than 32 ranks")
        self.root = root
        self.collectives = 0
        if shard is None:
            vo_count = P.Program.from_address(prg.ptr).vo_count
            self.voice_range = voice_shards(prg, srate, self.world)[self.rank]
            if device is None:
                device = torch.cuda.current_device()
            row = max_call_len if max_call_len else srate * 256 // 1000
            row = max(4, (row + 3) & ~3)           # runtime.cpp: row_len rounding
            vr = self.voice_range
            if vr[0] == vr[1]:
                vr = (vo_count, vo_count + 1)      # empty shard: renders nothing
            shard = _CudaShard(prg, srate, vr, device, max_call_len, tables)
            shard.set_row_len(row)
        self.shard = shard
        self.ended = False
        self._ctrl = None

    def run(self, buf_len, stereo=True):
        """-> (more, pcm or None, out_len); ONE all-reduce of planes + control words per call."""
        dist, torch = self.dist, self.torch
        more, planes, n = self.shard.run_mix(buf_len)
        body = planes
        if self.world > 1:
            w = self.world
            off = getattr(self.shard, "tail_offset", None)
            if off is None or planes.numel() < off + 2 * w:      # a stand-in without spare floats
                off = planes.numel()
                planes = torch.cat([planes, torch.zeros(2 * w, dtype=planes.dtype, device=planes.device)])
            if self._ctrl is None:
                self._ctrl = torch.zeros(2 * w, dtype=torch.float32)
                if planes.is_cuda:
                    self._ctrl = self._ctrl.pin_memory()
            self._ctrl.zero_()
            self._ctrl[self.rank] = 1.0 if more else 0.0
            self._ctrl[w + self.rank] = float(n)
            planes[off:off + 2 * w].copy_(self._ctrl, non_blocking=True)
            dist.all_reduce(planes, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
            tail = planes[off:off + 2 * w].cpu()                 # (the root needs the sum on the host anyway)
            more = bool(tail[:w].max().item() > 0)
            n = int(tail[w:].max().item())
            body = planes[:off]
        pcm = None
        if self.rank == self.root:
            pcm = self.shard.to_pcm(body, buf_len, stereo)
        if not more:
            self.ended = True
        return more, pcm, (buf_len if more else n)

    def render(self, call_len, stereo=True):
        """Whole program -> int16 [frames, ch] on the root, None elsewhere."""
        ch = 2 if stereo else 1
        chunks, more = [], True
        while more:
            more, pcm, n = self.run(call_len, stereo)
            if pcm is not None:
                chunks.append(pcm[:n * ch].copy())
        if self.rank != self.root:
            return None
        return np.concatenate(chunks).reshape(-1, ch) if chunks else np.zeros((0, ch), np.int16)

    def close(self):
        self.shard.close()
