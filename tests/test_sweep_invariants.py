"""The algebra the raycast sweep (k_sweep_compact / k_sweep_walk) relies on, checked on the CPU
against the oracle's grid_map LineIterator restatement:
  1. de-duplication identity: per-beam traversal == sum over distinct end cells of w_e * line(s,e);
  2. lines to side-by-side end cells at the same major-axis distance D have D + 1 cells each and,
     at every step k, the same major coordinate and minor coordinates that are MONOTONE in the end
     cell's minor coordinate (so equal cells form contiguous runs — one RED per run);
  3. the closed form of the iterator state the kernels use to start a line."""
import numpy as np

from oracle import gv_oracle as orc


def test_dedup_identity_matches_per_beam_oracle():
    g = orc.Grid.from_cells(96, 80, 0.5)
    rng = np.random.default_rng(0)
    n = 6000
    x = rng.uniform(-30, 30, n).astype(np.float32)
    y = rng.uniform(-25, 25, n).astype(np.float32)
    z = np.zeros(n, np.float32)
    T = np.eye(4, dtype=np.float32)
    T[0, 3], T[1, 3] = 1.3, -0.7
    upd, cells, flags = g.accumulate(T, x, y, z, r_max=18.0)
    sx, sy = g.get_index(1.3, -0.7)
    miss = np.zeros(g.nx * g.ny, np.int64)
    hit = np.zeros_like(miss)
    valid = cells >= 0
    ends, inv = np.unique(cells[valid], return_inverse=True)
    w = np.bincount(inv)
    h = np.bincount(inv, weights=((flags[valid] & orc.F_HIT) > 0)).astype(np.int64)
    for e, we, he in zip(ends, w, h):
        line = orc.bresenham_cells(sx, sy, int(e % g.nx), int(e // g.nx))
        lin = line[:, 0] + line[:, 1] * g.nx
        np.add.at(miss, lin[:-1], we)
        hit[lin[-1]] += he
        miss[lin[-1]] += we - he
    assert np.array_equal(miss, g.miss) and np.array_equal(hit, g.hit)
    assert len(ends) < valid.sum()          # duplicates existed, so the identity was exercised


def test_sweep_rows_are_equal_length_and_monotone():
    rng = np.random.default_rng(1)
    for _ in range(40):
        sx, sy = (int(v) for v in rng.integers(-20, 20, 2))
        D = int(rng.integers(1, 70))
        for direction in range(4):
            if direction < 2:   # x-major: column sx +- D, |dy| <= D
                ex = sx + D if direction == 0 else sx - D
                ends = [(ex, sy + a) for a in range(-D, D + 1)]
                major, minor = 0, 1
            else:               # y-major: row sy +- D, |dx| < D
                ey = sy + D if direction == 2 else sy - D
                ends = [(sx + a, ey) for a in range(-(D - 1), D)]
                major, minor = 1, 0
            if not ends:
                continue
            lines = np.stack([orc.bresenham_cells(sx, sy, ex, ey) for ex, ey in ends])   # [n, D+1, 2]
            assert lines.shape[1] == D + 1
            smaj = 1 if direction in (0, 2) else -1
            start = sx if major == 0 else sy
            assert np.all(lines[:, :, major] == start + smaj * np.arange(D + 1)[None, :])
            assert np.all(np.diff(lines[:, :, minor], axis=0) >= 0), (sx, sy, D, direction)
            assert np.all(np.diff(lines[:, :, minor], axis=0) <= 1)     # neighbours never jump a cell
            # a subsequence of the row (what compaction keeps) is still monotone
            pick = np.sort(rng.choice(len(ends), size=max(1, len(ends) // 3), replace=False))
            assert np.all(np.diff(lines[pick][:, :, minor], axis=0) >= 0)


def test_line_iterator_closed_form():
    rng = np.random.default_rng(2)
    for _ in range(300):
        sx, sy, ex, ey = (int(v) for v in rng.integers(-60, 60, 4))
        line = orc.bresenham_cells(sx, sy, ex, ey)
        dx, dy = abs(ex - sx), abs(ey - sy)
        D, A = max(dx, dy), min(dx, dy)
        if D == 0:
            continue
        k0 = int(rng.integers(0, D + 1))
        t0 = D // 2 + k0 * A
        minor_off = t0 // D                     # kernels: minor = origin + s * (t0 / den), num = t0 % den
        xmajor = dx >= dy
        sgn_minor = (1 if ey >= sy else -1) if xmajor else (1 if ex >= sx else -1)
        exp_minor = (sy if xmajor else sx) + sgn_minor * minor_off
        assert line[k0, 1 if xmajor else 0] == exp_minor
