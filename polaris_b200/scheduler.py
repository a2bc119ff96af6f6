"""Block schedulers, restated from reference tracer/scheduler.go:6-106.

The CUDA backend is driven by these unchanged in the Go renderer; they are restated here because
the multi-GPU bench and the tests need the same row assignment without a Go toolchain.  A scheduler
only needs `speed()` and `stats()` of each tracer (`Tracer.Speed()`, `Tracer.Stats()`), so any object
with those two methods works; `StaticSpeed` is such a stand-in for tracers living in other processes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass
class _Stats:
    block_h: int = 0
    render_time: float = 0.0  # seconds


class StaticSpeed:
    """A (speed, last-frame stats) record standing in for a remote tracer."""

    def __init__(self, speed: int):
        self._speed = int(speed)
        self._stats = _Stats()

    def speed(self):
        return self._speed

    def stats(self):
        return self._stats

    def set_stats(self, block_h: int, render_time_s: float):
        self._stats = _Stats(int(block_h), float(render_time_s))

    @property
    def block_h(self):
        return self._stats.block_h


def _u32(x: float) -> int:
    # Go's uint32(float64) conversion truncates toward zero
    return int(x) & 0xFFFFFFFF


def assign_blocks_based_on_speed(tracers, frame_h: int):
    """scheduler.go:83-106"""
    speed_sum = sum(int(t.speed()) for t in tracers) & 0xFFFFFFFF
    scaler = float(frame_h) / float(speed_sum)
    out, assigned = [], 0
    for t in tracers:
        block_h = _u32(max(1.0, float(t.speed()) * scaler))
        assigned += block_h
        out.append(block_h)
    if assigned < frame_h:
        out[0] += frame_h - assigned
    return out


class NaiveScheduler:
    """scheduler.go:13-30: rows proportional to Speed(), computed once."""

    def __init__(self):
        self.block_assignment = []

    def schedule(self, tracers, frame_h: int):
        if len(self.block_assignment) != len(tracers):
            self.block_assignment = assign_blocks_based_on_speed(tracers, frame_h)
        return self.block_assignment


class PerfectScheduler:
    """scheduler.go:34-80: rows proportional to BlockH / RenderTime of the previous frame."""

    def __init__(self):
        self.block_assignment = []

    def schedule(self, tracers, frame_h: int):
        if len(self.block_assignment) != len(tracers):
            self.block_assignment = assign_blocks_based_on_speed(tracers, frame_h)
            return self.block_assignment
        total = 0.0
        for t in tracers:
            st = t.stats()
            total += float(st.block_h) / float(_ns(st.render_time))
        scaler = float(frame_h) / total
        assigned = 0
        for i, t in enumerate(tracers):
            st = t.stats()
            block_h = _u32(max(1.0, math.floor(float(st.block_h) / float(_ns(st.render_time)) * scaler)))
            assigned += block_h
            self.block_assignment[i] = block_h
        if assigned < frame_h:
            self.block_assignment[0] += frame_h - assigned
        return self.block_assignment


def _ns(seconds: float) -> int:
    """RenderTime.Nanoseconds()"""
    return int(round(seconds * 1e9))
