"""Host logic of the forward pipelining by direction halves (las_b200.functional.direction_half_schedule), checked
against a step-by-step simulation of a BiLSTM layer's two sweeps -- no GPU needed: the schedule is plain integer arithmetic, and a
wrong readiness count would let a gate-projection tile read frames the recurrence kernel has not written yet."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'attention-based-e2e-asr-dnn_b200'))


def _simulate_ready(T, fac, t0, t1, d):
    """Smallest number of completed recurrence steps after which direction d has written every frame of rows [t0, t1)."""
    frames = set(range(fac * t0, min(fac * t1, T)))
    done = set()
    for s in range(T):
        done.add(s if d == 0 else T - 1 - s)          # frame processed at step s (reference: PackedSequence order, src/modules.py:80)
        if frames <= done:
            return s + 1
    return T


@pytest.mark.parametrize('T,fac', [(1600, 2), (800, 2), (400, 2), (200, 2), (1601, 2), (97, 2), (33, 2), (1100, 2), (640, 1), (129, 1), (31, 1)])
def test_direction_half_schedule_matches_sweep_simulation(T, fac):
    from las_b200.functional import direction_half_schedule, _PROGRESS_EVERY as EVERY, _PIPE_TILE as TILE
    Tn = T // fac                                        # the pyramid drops an odd last frame (src/modules.py:171-175)
    early, late = direction_half_schedule(Tn, T, fac)
    kmax = (T - 1) // EVERY                              # the kernel publishes at steps EVERY, 2*EVERY, ... < T
    tiles = [(t0, min(t0 + TILE, Tn)) for t0 in range(0, Tn, TILE)]
    # every (tile, direction) is issued exactly once: as an early half, as a late half, or inside a late full-K GEMM
    seen = {}
    for k, d, t0, t1 in early:
        assert (t0, t1) in tiles and d in (0, 1)
        assert (t0, d) not in seen
        seen[(t0, d)] = ('early', k)
    for d, t0, t1 in late:
        assert (t0, t1) in tiles
        for dd in ((0, 1) if d is None else (d,)):
            assert (t0, dd) not in seen
            seen[(t0, dd)] = ('late', None)
    assert set(seen) == {(t0, d) for t0, _ in tiles for d in (0, 1)}
    # a full-K late GEMM only when NEITHER half was early (otherwise the other half has already written / will accumulate)
    for d, t0, t1 in late:
        if d is None:
            assert not any(e[2] == t0 for e in early)
    # readiness: progress count k means "steps < EVERY * k are complete"; the half must not run before its frames exist, and the
    # schedule must not wait longer than one publish interval beyond that
    assert early == sorted(early)
    for k, d, t0, t1 in early:
        need = _simulate_ready(T, fac, t0, t1, d)
        assert 1 <= k <= kmax
        assert EVERY * k >= need, (k, d, t0, t1, need)
        assert EVERY * (k - 1) < need, (k, d, t0, t1, need)
    # what is left for the end of the kernel really cannot be released by any published count
    for d, t0, t1 in late:
        for dd in ((0, 1) if d is None else (d,)):
            assert _simulate_ready(T, fac, t0, t1, dd) > EVERY * kmax


def test_direction_halves_release_work_from_the_start():
    """The point of the halves: whole tiles (both sweeps must have crossed them) release nothing before the middle of the kernel, halves
    release work within the first two publish intervals and at a uniform rate."""
    from las_b200.functional import direction_half_schedule, _time_tiles, _PROGRESS_EVERY as EVERY
    T, fac = 1600, 2
    early_h, late_h = direction_half_schedule(T // fac, T, fac)
    early_w, late_w = _time_tiles(T // fac, T, fac, True)
    assert min(k for k, *_ in early_w) * EVERY >= T // 2
    assert min(k for k, *_ in early_h) * EVERY <= 2 * EVERY
    assert len(early_h) >= 2 * len(early_w)
    ks = sorted(k for k, *_ in early_h)
    assert max(b - a for a, b in zip(ks, ks[1:])) * EVERY <= 256 + EVERY          # no gap longer than one tile's frames


@pytest.mark.parametrize('T,fac', [(1600, 2), (800, 2), (401, 2), (200, 2), (1600, 1), (800, 1), (130, 1), (64, 1)])
def test_whole_tile_schedule_matches_sweep_simulation(T, fac):
    """functional._time_tiles: whole 128-row tiles behind BOTH sweeps (LAS_FWD_KSPLIT=0, and every layer's dX GEMM behind its BPTT
    kernel, whose two sweeps run the other way round -- the condition is symmetric in the directions)."""
    from las_b200.functional import _time_tiles, _PROGRESS_EVERY as EVERY, _PIPE_TILE as TILE
    Tn = T // fac
    early, late = _time_tiles(Tn, T, fac, True)
    kmax = (T - 1) // EVERY
    tiles = [(t0, min(t0 + TILE, Tn)) for t0 in range(0, Tn, TILE)]
    assert sorted((t0, t1) for _, t0, t1 in early + late) == tiles
    assert early == sorted(early)
    for k, t0, t1 in early:
        need = max(_simulate_ready(T, fac, t0, t1, 0), _simulate_ready(T, fac, t0, t1, 1))
        assert 1 <= k <= kmax and EVERY * k >= need and EVERY * (k - 1) < need, (k, t0, t1, need)
    for _, t0, t1 in late:
        assert max(_simulate_ready(T, fac, t0, t1, 0), _simulate_ready(T, fac, t0, t1, 1)) > EVERY * kmax
    # a kernel that publishes nothing releases nothing early
    e0, l0 = _time_tiles(Tn, T, fac, False)
    assert e0 == [] and sorted((t0, t1) for _, t0, t1 in l0) == tiles
