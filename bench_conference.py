#!/usr/bin/env python
"""BASELINE cfg3: 16-party MSAudioMixer conference x 1024 rooms, pins striped over the GPUs (gpu = pin mod N) — the only
place on the hot path with a real cross-GPU exchange. `conference_block()` is what `bench.py --gpus N` (N > 1) adds to its
JSON line; run alone: `python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
bench_conference.py --gpus N`.

Three layouts per run, each checked bit for bit on rank 0 against the oracle's single-mixer result and then timed with CUDA
events on the launching stream (max over ranks), inputs rotating over a pool larger than the L2:
  room_local  whole rooms per GPU (rooms/N each, all 16 pins): the production sharding, no collective — the baseline
  nccl        striped; msb200_mixer_process_striped_dev = partial -> ncclAllReduce(int32) -> finish (3 launches)
  fused       striped; msb200_mixer_xchg_process_dev = ONE kernel pushing partial sums over NVLink peer memory
Everything goes through the C ABI (include/msb200dsp.h); torch.distributed is the side channel for ids / handles only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from mediastreamer2_b200 import _lib  # noqa: E402
from mediastreamer2_b200 import conference as conf  # noqa: E402
from mediastreamer2_b200 import filters as F  # noqa: E402

ROOMS, PINS, NWORDS = 1024, 16, 480
POOL_BYTES = 300 << 20  # inputs + outputs of the buffer sets a layout rotates over: > 2x the 126 MB L2


def _oracle_mix(pcm, present, gain, active):
    import _oracle as O

    L = O.oracle()
    exp = np.zeros_like(pcm)
    r, p, n = pcm.shape
    L.orc_mixer_process(r, p, n, 1, O.ptr(gain), O.ptr(active), O.ptr(pcm), O.ptr(present), O.ptr(exp))
    return exp


class _Layout:
    """device buffers of one layout on this rank: a pool of (in, present, out) sets"""

    def __init__(self, ctx, rooms: int, pins: int, nwords: int):
        self.ctx, self.shape = ctx, (rooms, pins, nwords)
        self.in_bytes = rooms * pins * nwords * 2
        self.n_sets = max(2, -(-POOL_BYTES // (2 * self.in_bytes)))
        self.d_in = ctx.dev_alloc(self.n_sets * self.in_bytes)
        self.d_out = ctx.dev_alloc(self.n_sets * self.in_bytes)
        self.d_pr = ctx.dev_alloc(rooms * pins)
        self.k = 0

    def load(self, pcm: np.ndarray, present: np.ndarray, every_set: bool):
        pcm = np.ascontiguousarray(pcm)
        for s in (range(self.n_sets) if every_set else [0]):
            self.ctx.h2d(self.d_in + s * self.in_bytes, pcm)
        self.ctx.h2d(self.d_pr, np.ascontiguousarray(present))
        self.k = 0

    def next(self):
        s = self.k % self.n_sets
        self.k += 1
        return self.d_in + s * self.in_bytes, self.d_pr, self.d_out + s * self.in_bytes

    def read_out(self, s: int = 0) -> np.ndarray:
        out = np.empty(self.shape, np.int16)
        self.ctx.d2h(out, self.d_out + s * self.in_bytes)
        return out

    def close(self):
        for p in (self.d_in, self.d_out, self.d_pr):
            self.ctx.dev_free(p)


def conference_block(ctx, rank: int, world: int, dist, steps: int = 200, warmup: int = 10, check_ticks: int = 3,
                     peak_gbs: float | None = None) -> dict | None:
    """returns the `conference` object of the bench line on rank 0 (None elsewhere). `ctx` is this rank's msb200 context."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())

    def allgather(obj):
        if world == 1:
            return [obj]
        got = [None] * world
        dist.all_gather_object(got, obj)
        return got

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(tick) -> float:
        for _ in range(max(warmup, 3)):
            tick()
        barrier()
        ctx.timer_start()
        for _ in range(steps):
            tick()
        return max_over_ranks(ctx.timer_stop_ms())

    results = {}
    slot_bytes = ROOMS * NWORDS * 4
    # ---------------------------------------------------------------- room-local baseline (no exchange)
    assert ROOMS % world == 0
    rl = ROOMS // world
    lay = _Layout(ctx, rl, PINS, NWORDS)
    mixer = F.AudioMixer(ctx, rl, PINS, NWORDS, True)
    _, _, gain, active = conf.cfg3_inputs(ROOMS, PINS, NWORDS, 0)
    for r, k in zip(*np.nonzero(gain[:rl] != 1.0)):
        mixer.set_input_gain(int(r), int(k), float(gain[r, k]))
    for r, k in zip(*np.nonzero(active[:rl] == 0)):
        mixer.set_active(int(r), int(k), False)
    ok = True
    for t in range(check_ticks):
        pcm, present, gain, active = conf.cfg3_inputs(ROOMS, PINS, NWORDS, t)
        mine = slice(rank * rl, (rank + 1) * rl)
        lay.load(pcm[mine], present[mine], every_set=(t == check_ticks - 1))
        d_in, d_pr, d_out = lay.next()
        _lib.check(ctx.lib.msb200_mixer_process_dev(mixer.h, d_in, d_pr, d_out))
        ctx.sync()
        parts = allgather(lay.read_out(0))
        if rank == 0:
            ok = ok and bool(np.array_equal(np.concatenate(parts), _oracle_mix(pcm, present, gain, active)))

    def tick_local():
        d_in, d_pr, d_out = lay.next()
        _lib.check(ctx.lib.msb200_mixer_process_dev(mixer.h, d_in, d_pr, d_out))

    l0 = ctx.launches
    ms = timed(tick_local)
    results["room_local"] = {"ms_per_step": ms / steps, "room_ticks_per_s": ROOMS * steps / (ms / 1000.0),
                             "bit_exact_vs_oracle": ok, "nvlink_bytes_per_tick_per_gpu": 0,
                             "launches_per_tick": (ctx.launches - l0) / (steps + max(warmup, 3)),
                             "what": f"{rl} whole rooms x {PINS} pins per GPU, mixer_kernel, no exchange"}
    mixer.close()
    lay.close()
    # ---------------------------------------------------------------- striped: the two exchanges
    for exchange in conf.EXCHANGES:
        if exchange == "nccl" and not ctx.lib.msb200_comm_available():
            results[exchange] = {"error": _lib.load().msb200_last_error().decode()}
            continue
        sc = conf.StripedConference(ctx, rank, world, ROOMS, PINS, NWORDS, exchange, allgather, barrier)
        lay = _Layout(ctx, ROOMS, sc.nl, NWORDS)
        sc.set_controls(gain, active)
        ok = True
        for t in range(check_ticks):
            pcm, present, gain, active = conf.cfg3_inputs(ROOMS, PINS, NWORDS, t)
            if t == 1:
                present[:, 5] = 0  # a starving pin: contributes zeros, still receives the mix (audiomixer.c:88)
            lpcm, lpres, _ = conf.shard_inputs(pcm, present, rank, world)
            lay.load(lpcm, lpres, every_set=(t == check_ticks - 1))
            sc.tick_dev(*lay.next())
            ctx.sync()
            parts = allgather(lay.read_out(0))
            if rank == 0:
                full = np.zeros((ROOMS, PINS, NWORDS), np.int16)
                for r2 in range(world):
                    conf.scatter_outputs(full, parts[r2], r2, world)
                ok = ok and bool(np.array_equal(full, _oracle_mix(pcm, present, gain, active)))
        l0 = ctx.launches
        ms = timed(lambda: sc.tick_dev(*lay.next()))
        launches = (ctx.launches - l0) / (steps + max(warmup, 3))
        timeouts = sum(allgather(sc.timeouts()))
        wire = sc.wire_bytes_per_tick()
        res = {"ms_per_step": ms / steps, "room_ticks_per_s": ROOMS * steps / (ms / 1000.0), "bit_exact_vs_oracle": ok,
               "nvlink_bytes_per_tick_per_gpu": wire, "launches_per_tick": launches + (1 if exchange == "nccl" else 0),
               "nvlink_gbs_per_gpu": wire / (ms / steps / 1000.0) / 1e9}
        if exchange == "nccl":
            res["what"] = (f"partial -> ncclAllReduce(int32 SUM, {slot_bytes} B, NCCL {ctx.lib.msb200_comm_nccl_version()}, "
                           f"dlopen'ed, same stream) -> finish; launches_per_tick counts the NCCL kernel")
        else:
            res["what"] = (f"one kernel per tick: {world - 1} x {slot_bytes} B pushed to the peers' receive slots, per-CTA epoch "
                           f"flags polled in local memory, no NCCL on the data path")
            res["flag_wait_timeouts"] = timeouts
        results[exchange] = res
        lay.close()
        sc.close()
    if rank != 0:
        return None
    best = min((k for k in conf.EXCHANGES if "ms_per_step" in results[k]), key=lambda k: results[k]["ms_per_step"])
    return {"workload": f"cfg3: {ROOMS} rooms x {PINS} pins x {NWORDS} samples (48 kHz, 10 ms), conference mode, pin 3 gain 0.5, "
                        f"pin 7 muted, pins striped gpu = pin mod {world}",
            "unit": "room-ticks/s", "n_gpus": world, "steps": steps, "scaling": "strong",
            "l2": f"inputs/outputs rotate over buffer sets totalling >= {POOL_BYTES >> 20} MiB per layout (> 2x L2)",
            "value": results[best]["room_ticks_per_s"], "best_exchange": best,
            "bit_exact_vs_oracle": all(r.get("bit_exact_vs_oracle", False) for r in results.values()),
            **results}


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--check-ticks", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = F.Context(local_rank)
    block = conference_block(ctx, rank, world, dist, args.steps, args.warmup, args.check_ticks)
    if rank == 0:
        print(json.dumps(block), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
