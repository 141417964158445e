"""Cross-GPU conference mixing (BASELINE cfg3): host-side sharding logic.

A room's pins are striped over the ranks (``gpu = pin mod world``, the layout cfg3 asks for; production would keep whole
rooms on one GPU and need no exchange at all). Per tick every rank:
  1. computes the int32 partial sum of its LOCAL pins per room          (msb200_mixer_partial_dev)
  2. all-reduces the [rooms][nwords] int32 buffer with SUM               (NCCL over NVLink; integer => order-independent
                                                                          => bit-exact regardless of the reduction tree)
  3. emits sat(total - own) for its local pins                           (msb200_mixer_finish_dev)
This module holds only index arithmetic so that the same code drives the GPU path (bench_conference.py) and the
world-size-2 gloo test on CPU (tests/test_multi_gpu_mix.py, where the oracle stands in for the kernels).
"""
from __future__ import annotations

import numpy as np


def local_pins(rank: int, world: int, n_pins: int) -> np.ndarray:
    """global pin indices owned by `rank` under the striped layout"""
    return np.arange(rank, n_pins, world)


def shard_inputs(pcm: np.ndarray, present: np.ndarray, rank: int, world: int):
    """[rooms][pins][nwords] -> this rank's [rooms][local_pins][nwords] (contiguous) and presence flags"""
    lp = local_pins(rank, world, pcm.shape[1])
    return np.ascontiguousarray(pcm[:, lp, :]), np.ascontiguousarray(present[:, lp]), lp


def shard_controls(gain: np.ndarray, active: np.ndarray, rank: int, world: int):
    lp = local_pins(rank, world, gain.shape[1])
    return np.ascontiguousarray(gain[:, lp]), np.ascontiguousarray(active[:, lp])


def scatter_outputs(full: np.ndarray, local_out: np.ndarray, rank: int, world: int) -> None:
    """place this rank's [rooms][local_pins][nwords] outputs into the full [rooms][pins][nwords] array"""
    full[:, local_pins(rank, world, full.shape[1]), :] = local_out


def cfg3_inputs(n_rooms: int, n_pins: int, nwords: int, tick: int):
    """BASELINE cfg3 synthetic tick: uniform noise in +-6000, every 97th block at +-30000 (saturation paths);
    gains: pin 3 = 0.5; pin 7 muted."""
    rng = np.random.default_rng(0xC3000 + tick)
    pcm = rng.integers(-6000, 6001, size=(n_rooms, n_pins, nwords)).astype(np.int16)
    blk = np.arange(n_rooms * n_pins).reshape(n_rooms, n_pins) + tick * 31
    hot = (blk % 97) == 0
    pcm[hot] = np.where(pcm[hot] >= 0, 30000, -30000).astype(np.int16)
    present = np.ones((n_rooms, n_pins), np.uint8)
    gain = np.ones((n_rooms, n_pins), np.float32)
    active = np.ones((n_rooms, n_pins), np.uint8)
    if n_pins > 3:
        gain[:, 3] = 0.5
    if n_pins > 7:
        active[:, 7] = 0
    return pcm, present, gain, active


# ------------------------------------------------------------------------------------------------ device-side driver
# Plumbing over the C ABI (include/msb200dsp.h "cross-GPU conference exchange"); nothing below computes on the CPU.
EXCHANGES = ("nccl", "fused")


class StripedConference:
    """One rank's share of `rooms` conferences of `pins` pins striped over `world` GPUs (gpu = pin mod world).

    exchange = "nccl":  msb200_mixer_process_striped_dev (partial -> ncclAllReduce int32 -> finish, 3 launches per tick)
    exchange = "fused": msb200_mixer_xchg_process_dev (ONE kernel: push partial sums over NVLink, per-CTA epoch flags, finish)
    `allgather(obj) -> list` is the caller's side channel (torch.distributed.all_gather_object, or a list built by hand when
    the ranks are contexts of one process); `barrier()` likewise.
    """

    def __init__(self, ctx, rank: int, world: int, rooms: int, pins: int, nwords: int, exchange: str, allgather, barrier):
        import ctypes as C

        from . import _lib
        from . import filters as F

        assert exchange in EXCHANGES
        self.C, self._lib = C, _lib
        self.ctx, self.lib, self.rank, self.world, self.exchange = ctx, ctx.lib, rank, world, exchange
        self.rooms, self.pins, self.nwords = rooms, pins, nwords
        self.lp = local_pins(rank, world, pins)
        self.nl = len(self.lp)
        self.mixer = F.AudioMixer(ctx, rooms, self.nl, nwords, True)
        self.barrier = barrier
        self.comm = self.xchg = None
        self.d_sum = None
        if exchange == "nccl":
            uid = (C.c_uint8 * 128)()
            if rank == 0:
                _lib.check(self.lib.msb200_comm_unique_id(uid))
            uid0 = allgather(bytes(uid))[0]
            h = C.c_void_p()
            _lib.check(self.lib.msb200_comm_create(ctx.h, (C.c_uint8 * 128).from_buffer_copy(uid0), rank, world, C.byref(h)))
            self.comm = h
            self.d_sum = ctx.dev_alloc(rooms * nwords * 4)
        else:
            h = C.c_void_p()
            _lib.check(self.lib.msb200_mixer_xchg_create(self.mixer.h, rank, world, C.byref(h)))
            self.xchg = h
            mine = (C.c_uint8 * 64)()
            _lib.check(self.lib.msb200_mixer_xchg_export(h, mine))
            handles = b"".join(allgather(bytes(mine)))
            _lib.check(self.lib.msb200_mixer_xchg_connect(h, (C.c_uint8 * len(handles)).from_buffer_copy(handles)))
            barrier()  # every rank's flags are zeroed and mapped before the first push

    def set_controls(self, gain: np.ndarray, active: np.ndarray) -> None:
        """full [rooms][pins] tables -> this rank's pins (MS_AUDIO_MIXER_SET_INPUT_GAIN / SET_ACTIVE)"""
        lgain, lact = shard_controls(gain, active, self.rank, self.world)
        for r, k in zip(*np.nonzero(lgain != 1.0)):
            self.mixer.set_input_gain(int(r), int(k), float(lgain[r, k]))
        for r, k in zip(*np.nonzero(lact == 0)):
            self.mixer.set_active(int(r), int(k), False)

    def tick_dev(self, d_in: int, d_present: int, d_out: int) -> None:
        C, check = self.C, self._lib.check
        if self.comm is not None:
            check(self.lib.msb200_mixer_process_striped_dev(self.mixer.h, self.comm, C.c_void_p(d_in), C.c_void_p(d_present),
                                                            C.c_void_p(self.d_sum), C.c_void_p(d_out)))
        else:
            check(self.lib.msb200_mixer_xchg_process_dev(self.xchg, C.c_void_p(d_in), C.c_void_p(d_present), C.c_void_p(d_out)))

    def timeouts(self) -> int:
        if self.xchg is None:
            return 0
        n = self.C.c_uint32()
        self._lib.check(self.lib.msb200_mixer_xchg_status(self.xchg, self.C.byref(n)))
        return int(n.value)

    def wire_bytes_per_tick(self) -> int:
        """bytes this rank moves over NVLink per tick"""
        if self.xchg is not None:
            return int(self.lib.msb200_mixer_xchg_wire_bytes_per_tick(self.xchg))
        # ring all-reduce: 2 (world-1)/world of the buffer out of every rank
        return int(2 * (self.world - 1) / self.world * self.rooms * self.nwords * 4)

    def close(self) -> None:
        self.ctx.sync()
        self.barrier()  # nobody unmaps / frees while a peer may still push or read
        if self.xchg is not None:
            self.lib.msb200_mixer_xchg_destroy(self.xchg)
        if self.comm is not None:
            self.lib.msb200_comm_destroy(self.comm)
        if self.d_sum:
            self.ctx.dev_free(self.d_sum)
        self.mixer.close()
        self.xchg = self.comm = self.d_sum = None
