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
