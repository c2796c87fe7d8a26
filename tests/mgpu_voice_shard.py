"""Run under torchrun (one rank per GPU; launched by tests/test_multigpu.py when
the box has >= 2 GPUs): one script's voices sharded over the ranks, per-call
NCCL reduce of the float mix planes, root's PCM vs the reference <= 1 LSB; then
independent scripts dealt to ranks with no collective, bit-exact."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch
import torch.distributed as dist

import gpuutil
import scripts
from oracle import pyport, pyref
from saugns_b200 import multigpu as M


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tabs = gpuutil.ref_tables_for_gpu(pyport)
    # --- voice sharding: C3 sample and C4 sample (self-PM, AM) ---
    feats = scripts.feature_scripts()
    for text in [scripts.synth_c3(64, 1, fm="mix"), scripts.synth_c4(45, 1),
                 feats["seq_overlap"], feats["handover_twice"], feats["handover"]]:
        prg = pyref.Program(text)
        vg = M.VoiceShardedGenerator(prg, 96000, device=local, tables=tabs)
        got = vg.render(24576)
        ncalls = -(-max(1, pyref.render(prg, srate=96000).shape[0]) // 24576)
        assert vg.collectives == ncalls, (vg.collectives, ncalls)      # ONE NCCL op per call
        vg.close()
        if rank == 0:
            want = pyref.render(prg, srate=96000)
            assert got.shape == want.shape, (got.shape, want.shape)
            d = int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max())
            assert d <= 1, d
            print(f"voice-sharded x{world}: {want.shape[0]} frames, {ncalls} calls = {ncalls} all-reduces, "
                  f"max diff {d} LSB", flush=True)
        else:
            assert got is None
    # --- script sharding: no collective, bit-exact ---
    prgs = [pyref.Program(scripts.synth_c5_script(i)) for i in range(24)]
    mine = M.render_scripts(prgs, 96000, rank=rank, world=world, device=local, tables=tabs,
                            group_size=8)
    plan = M.shard_scripts([M.program_cost(p) for p in prgs], world)
    assert sorted(mine) == plan[rank]
    for i, pcm in mine.items():
        assert np.array_equal(pcm, pyref.render(prgs[i], srate=96000)), i
    cnt = torch.tensor([len(mine)], device="cuda")
    dist.all_reduce(cnt)
    assert int(cnt.item()) == len(prgs)
    if rank == 0:
        print(f"script-sharded x{world}: {len(prgs)} scripts bit-exact", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} OK", flush=True)


if __name__ == "__main__":
    main()
