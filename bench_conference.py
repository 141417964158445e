#!/usr/bin/env python
"""BASELINE cfg3: 16-party MSAudioMixer conference x 1024 rooms, pins striped over the GPUs (gpu = pin mod N), the only
place on the hot path with a real cross-GPU exchange: int32 partial sums -> all-reduce (NCCL over NVLink/NVSwitch, SUM) ->
local outputs. Launch: `python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
bench_conference.py --gpus N`. With N=1 it degenerates to the single-pass mixer kernel (no collective).

Checks bit-exactness on rank 0 against the oracle for the first ticks (gathered outputs == single-process 16-pin mix),
then times `--steps` ticks with CUDA events on the launching stream (max over ranks). Also reports the room-local
sharding (whole rooms per GPU, zero collectives), the layout a production deployment would use.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from mediastreamer2_b200 import _lib  # noqa: E402
from mediastreamer2_b200 import conference as conf  # noqa: E402
from mediastreamer2_b200 import filters as F  # noqa: E402

ROOMS, PINS, NWORDS = 1024, 16, 480


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--check-ticks", type=int, default=3)
    ap.add_argument("--exchange", choices=["nccl", "peer"], default="nccl",
                    help="nccl: partial -> ncclAllReduce -> finish; peer: one fused kernel loading the peers' partial sums over NVLink")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    ctx = F.Context(local_rank, cuda_stream=stream.cuda_stream)  # our kernels and NCCL share torch's stream order
    lib = ctx.lib
    lp = conf.local_pins(rank, world, PINS)
    nl = len(lp)
    mixer = F.AudioMixer(ctx, ROOMS, nl, NWORDS, True)
    # controls (cfg3): pin 3 gain 0.5, pin 7 muted
    _, _, gain, active = conf.cfg3_inputs(ROOMS, PINS, NWORDS, 0)
    lgain, lact = conf.shard_controls(gain, active, rank, world)
    for k in range(nl):
        if lgain[0, k] != 1.0 or not lact[0, k]:
            for r in range(ROOMS):
                if lgain[r, k] != 1.0:
                    mixer.set_input_gain(r, k, float(lgain[r, k]))
                if not lact[r, k]:
                    mixer.set_active(r, k, False)
    d_in = torch.empty((ROOMS, nl, NWORDS), dtype=torch.int16, device=dev)
    d_pr = torch.empty((ROOMS, nl), dtype=torch.uint8, device=dev)
    d_sum = torch.empty((ROOMS, NWORDS), dtype=torch.int32, device=dev)
    d_out = torch.empty((ROOMS, nl, NWORDS), dtype=torch.int16, device=dev)

    # ---- peer-memory exchange set-up: two alternating partial-sum buffers + an epoch flag per rank, mapped everywhere
    peer = {}
    if args.exchange == "peer" and world > 1:
        nbytes = ROOMS * NWORDS * 4
        own = {"sum0": ctx.dev_alloc(nbytes), "sum1": ctx.dev_alloc(nbytes), "flag": ctx.dev_alloc(256)}
        err = ctx.dev_alloc(256)
        _lib.check(lib.msb200_memset_dev(ctx.h, own["flag"], 0, 256))
        _lib.check(lib.msb200_memset_dev(ctx.h, err, 0, 256))
        ctx.sync()
        handles = {}
        for k, p in own.items():
            h = (C.c_uint8 * 64)()
            _lib.check(lib.msb200_ipc_export(ctx.h, C.c_void_p(p), h))
            handles[k] = bytes(h)
        allh = [None] * world
        dist.all_gather_object(allh, handles)
        mapped = []
        for r2 in range(world):
            if r2 == rank:
                mapped.append(own)
                continue
            m = {}
            for k, hb in allh[r2].items():
                q = C.c_void_p()
                _lib.check(lib.msb200_ipc_import(ctx.h, (C.c_uint8 * 64).from_buffer_copy(hb), C.byref(q)))
                m[k] = q.value
            mapped.append(m)
        peer = {"own": own, "err": err, "mapped": mapped, "epoch": 0,
                "sums": [(C.c_void_p * world)(*[m["sum0"] for m in mapped]), (C.c_void_p * world)(*[m["sum1"] for m in mapped])],
                "flags": (C.c_void_p * world)(*[m["flag"] for m in mapped])}
        dist.barrier()

    def tick():
        if peer:
            b = peer["epoch"] & 1
            peer["epoch"] += 1
            _lib.check(lib.msb200_mixer_partial_dev(mixer.h, d_in.data_ptr(), d_pr.data_ptr(), peer["own"][f"sum{b}"]))
            _lib.check(lib.msb200_signal_dev(ctx.h, peer["own"]["flag"], peer["epoch"]))
            _lib.check(lib.msb200_mixer_finish_peers_dev(mixer.h, d_in.data_ptr(), d_pr.data_ptr(), peer["sums"][b], peer["flags"],
                                                         world, peer["epoch"], d_out.data_ptr(), peer["err"]))
            return
        if world == 1:
            _lib.check(lib.msb200_mixer_process_dev(mixer.h, d_in.data_ptr(), d_pr.data_ptr(), d_out.data_ptr()))
            return
        _lib.check(lib.msb200_mixer_partial_dev(mixer.h, d_in.data_ptr(), d_pr.data_ptr(), d_sum.data_ptr()))
        dist.all_reduce(d_sum, op=dist.ReduceOp.SUM)  # ncclAllReduce(ncclInt32, ncclSum) over NVLink
        _lib.check(lib.msb200_mixer_finish_dev(mixer.h, d_in.data_ptr(), d_pr.data_ptr(), d_sum.data_ptr(), d_out.data_ptr()))

    # ---- parity: gathered striped outputs == oracle's single-process conference mix
    parity = True
    for t in range(args.check_ticks):
        pcm, present, gain, active = conf.cfg3_inputs(ROOMS, PINS, NWORDS, t)
        lpcm, lpres, _ = conf.shard_inputs(pcm, present, rank, world)
        d_in.copy_(torch.from_numpy(lpcm))
        d_pr.copy_(torch.from_numpy(lpres))
        tick()
        out = d_out.to(torch.int32)
        gathered = [torch.empty_like(out) for _ in range(world)] if world > 1 else [out]
        if world > 1:
            dist.all_gather(gathered, out)
        if rank == 0:
            import _oracle as O

            L = O.oracle()
            full = np.zeros((ROOMS, PINS, NWORDS), np.int16)
            for r2 in range(world):
                conf.scatter_outputs(full, gathered[r2].cpu().numpy().astype(np.int16), r2, world)
            exp = np.zeros_like(full)
            L.orc_mixer_process(ROOMS, PINS, NWORDS, 1, O.ptr(gain), O.ptr(active), O.ptr(pcm), O.ptr(present), O.ptr(exp))
            parity = parity and bool(np.array_equal(full, exp))
    # ---- timing
    for _ in range(max(args.warmup, 3)):
        tick()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record()
    for _ in range(args.steps):
        tick()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    if peer:
        e = np.zeros(1, np.uint32)
        ctx.d2h(e, peer["err"])
        if e[0]:
            raise RuntimeError("peer-memory exchange: a peer's epoch flag never arrived")
        dist.barrier()  # nobody unmaps while a peer may still be reading
        for r2, m in enumerate(peer["mapped"]):
            if r2 != rank:
                for q in m.values():
                    lib.msb200_ipc_close(ctx.h, C.c_void_p(q))
    if rank == 0:
        room_ticks = ROOMS * args.steps
        line = {
            "metric": "cfg3: 16-party conference rooms mixed per second (conference mode, bit-exact)",
            "value": room_ticks / (ms / 1000.0), "unit": "room-ticks/s", "n_gpus": world, "steps": args.steps,
            "ms_per_step": ms / args.steps, "scaling": "strong", "higher_is_better": True, "dtype": "s16/int32",
            "config": {"workload": f"{ROOMS} rooms x {PINS} pins x {NWORDS} samples, pins striped gpu = pin mod N",
                       "exchange": "none (single-pass kernel)" if world == 1 else
                       (f"ncclAllReduce int32 SUM of {ROOMS * NWORDS * 4} B per tick between partial and finish kernels"
                        if not peer else
                        f"fused: finish kernel loads the {world} partial-sum buffers ({ROOMS * NWORDS * 4} B each) through "
                        f"NVLink peer mappings after an epoch-flag handshake; no NCCL on the data path")},
            "bit_exact_vs_oracle": parity, "gpu_launches": int(launches),
            "stream_ticks_per_s": room_ticks * PINS / (ms / 1000.0),
        }
        print(json.dumps(line), flush=True)
    mixer.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
