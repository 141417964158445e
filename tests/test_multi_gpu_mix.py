"""N>1 path on CPU: world_size 2, gloo. Each rank owns the striped half of every room's pins, computes partial sums
(oracle standing in for msb200_mixer_partial_dev), all-reduces int32 over gloo, finishes its local pins; the gathered
result must equal the single-process 16-pin conference mix bit for bit (integer sums are order-independent)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rooms, pins, nwords, ticks, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _oracle as O
    from _oracle import ptr
    from mediastreamer2_b200 import conference as conf

    L = O.oracle()
    ok = True
    for t in range(ticks):
        pcm, present, gain, active = conf.cfg3_inputs(rooms, pins, nwords, t)
        if t == 1:
            present[:, 5] = 0  # a starving pin: contributes zeros, still receives the mix
        lpcm, lpres, lp = conf.shard_inputs(pcm, present, rank, world)
        lgain, lact = conf.shard_controls(gain, active, rank, world)
        # phase 1: local partial sums
        part = np.zeros((rooms, nwords), np.int32)
        L.orc_mixer_partial(rooms, len(lp), nwords, ptr(lgain), ptr(lact), ptr(lpcm), ptr(lpres), ptr(part))
        # phase 2: the one exchange step
        tsum = torch.from_numpy(part)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        # phase 3: finish local pins from the reduced sum  (== msb200_mixer_finish_dev)
        total = tsum.numpy()
        lout = np.zeros_like(lpcm)
        for r in range(rooms):
            for k in range(len(lp)):
                own = np.zeros(nwords, np.int64)
                if lact[r, k] and lpres[r, k]:
                    x = lpcm[r, k].astype(np.int64)
                    if lgain[r, k] != 1.0:
                        x = np.clip((np.float32(lgain[r, k]) * lpcm[r, k].astype(np.float32)).astype(np.int32), -32767, 32767).astype(np.int64)
                    own = x
                lout[r, k] = np.clip(total[r].astype(np.int64) - (own if lact[r, k] else 0), -32767, 32767).astype(np.int16)
        # gather everybody's outputs on rank 0 and compare with the single-process mixer
        lout32 = torch.from_numpy(lout.astype(np.int32))  # gloo has no int16 collectives
        gathered = [torch.zeros_like(lout32) for _ in range(world)] if rank == 0 else None
        dist.gather(lout32, gathered, dst=0)
        if rank == 0:
            full = np.zeros((rooms, pins, nwords), np.int16)
            for r2 in range(world):
                conf.scatter_outputs(full, gathered[r2].numpy().astype(np.int16), r2, world)
            exp = np.zeros_like(full)
            L.orc_mixer_process(rooms, pins, nwords, 1, ptr(gain), ptr(active), ptr(pcm), ptr(present), ptr(exp))
            ok = ok and bool(np.array_equal(full, exp))
    if rank == 0:
        Path(outdir, "ok").write_text("1" if ok else "0")
    dist.destroy_process_group()


def test_striped_conference_all_reduce_is_bit_exact_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 6, 16, 160, 3, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "1"


def test_local_pin_striping_partitions_the_room():
    from mediastreamer2_b200 import conference as conf

    for world in (1, 2, 4, 8):
        seen = np.concatenate([conf.local_pins(r, world, 16) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(16))
