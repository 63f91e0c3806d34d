"""CPU checks of the arithmetic k_points_pair (csrc/gv_points_pair.cuh) adds on top of the certified
scheme of fast_point, on millions of samples and without a GPU:
  1. fma(p, 1, c) rounds exactly like the separate add (the kernel's way of keeping ptxas from
     contracting mul.rn.f32x2 + add.rn.f32x2): checked in binary64-emulated binary32;
  2. the image test |RN(q - W/2)| against the thresholds of fill_pair_args (gv_api.cu) certifies
     only what the interval [q - E(q), q + E(q)] certifies;
  3. the tile taken from the bits of fma(q, 1/S, 192) is the tile of some pixel within 2^-16 tiles
     of q: the 1-px dilation of the tile masks covers it;
  4. the closed form k_sweep_walk uses to start a line piece after k0 steps;
  5. the error budget of the 16.16 cell index with the 2-unit margin."""
import numpy as np

from oracle import gv_oracle as orc

f32, f64 = np.float32, np.float64
U22 = 2.0 ** -22


def host_pair_thresholds(W, c):
    """The numbers the library itself uses (gv_debug_pair_thresholds: host-only, needs no GPU)."""
    import ctypes as C
    from grid_vision_b200 import _lib
    out = np.zeros(6, f32)
    rc = _lib.load().gv_debug_pair_thresholds(C.c_double(W), C.c_float(c), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def test_python_thresholds_are_the_librarys():
    """The emulation below restates fill_pair_args; this pins it to the real host code bit for bit."""
    for W, c in ((416.0, 208.0), (640.0, 321.5), (1280.0, 600.25), (64.0, 10.0), (1920.0, 959.5)):
        e6, e0, half, e, ain, aout = pair_thresholds(W, c)
        got = host_pair_thresholds(W, c)
        exp = np.array([half, e, ain, aout, e6, e0], f32)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (W, c, got, exp)


def pair_thresholds(W, c):
    """fill_pair_args for one axis (image size W, principal point c); all in double, rounded to the
    safe side exactly like the host code."""
    e6 = f64(f32(6.0) * f32(U22))
    e0 = f64(f32(U22) * (f32(1.5) * abs(f32(c)) + f32(1.0)) * f32(1.0001))
    half = 0.5 * W
    e = f32(np.nextafter(f32((e6 * (W + 1.0) + e0) * 1.001), f32(np.inf)))
    ain = f32(np.nextafter(f32(half - f64(e) - W * 2.0 ** -21), f32(0.0)))
    aout = f32(np.nextafter(f32(((W + e0) / (1.0 - e6) - half) * (1.0 + 2.0 ** -21)), f32(np.inf)))
    return e6, e0, f32(half), e, ain, aout


def test_fma_with_one_equals_separate_add():
    rng = np.random.default_rng(0)
    n = 4_000_000
    # products and addends of wildly different magnitudes, including cancellation
    a = (rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))).astype(f32)
    b = (rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))).astype(f32)
    c = np.where(rng.random(n) < 0.3, -(a.astype(f64) * b.astype(f64)), rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))).astype(f32)
    p = (a * b).astype(f32)                                  # mul.rn: one rounding
    sep = (p + c).astype(f32)                                # add.rn: one rounding
    fused = (p.astype(f64) * 1.0 + c.astype(f64)).astype(f32)  # fma(p, 1, c): p*1 exact, p + c exact in double, one rounding
    assert np.array_equal(sep.view(np.uint32), fused.view(np.uint32))
    # and it differs from the contracted fma(a, b, c) often enough that a contraction would show
    contracted = (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)
    assert (contracted.view(np.uint32) != sep.view(np.uint32)).mean() > 0.05


def test_image_thresholds_certify_only_what_the_interval_does():
    for W, c in ((416.0, 208.0), (640.0, 321.5), (1280.0, 600.25), (64.0, 10.0)):
        e6, e0, half, e, ain, aout = pair_thresholds(W, c)
        rng = np.random.default_rng(int(W))
        q = np.concatenate([
            rng.uniform(-3 * W, 4 * W, 1_000_000),
            rng.normal(0.0, 2e-3, 500_000), W + rng.normal(0.0, 2e-3, 500_000),   # both image edges
            np.array([0.0, W, -0.0, np.nextafter(W, 0.0), 1e30, -1e30]),
        ]).astype(f32)
        E = np.abs(q.astype(f64)) * e6 + e0                   # the proven interval half-width
        lo, hi = q.astype(f64) - E, q.astype(f64) + E
        inside_all = (lo >= 0.0) & (hi < W)                   # every value of the interval is in [0, W)
        outside_all = (hi < 0.0) | (lo >= W)
        with np.errstate(invalid="ignore"):
            d = np.abs((q - half).astype(f32))                # |RN(q - W/2)| as the kernel computes it
        cert_in, cert_out = d < ain, d > aout
        assert not np.any(cert_in & ~inside_all), (W, c)
        assert not np.any(cert_out & ~outside_all), (W, c)
        # inside the image the constant margin dominates the proven one
        assert np.all(f64(e) >= E[cert_in])
        # an overflowed projection (q = +-Inf) is certainly outside, NaN is neither (-> deferred)
        with np.errstate(invalid="ignore"):
            dinf = np.abs((np.array([np.inf, -np.inf, np.nan], f32) - half).astype(f32))
        assert list(dinf > aout) == [True, True, False] and not np.any(dinf < ain)
        # and the band left to the exact pass is thin
        band = ~(cert_in | cert_out)
        assert band[:1_000_000].mean() < 1e-4


def test_tile_from_float_bits_is_within_the_dilation():
    rng = np.random.default_rng(1)
    for shift in (4, 5, 6):
        S = float(1 << shift)
        W = 13.0 * S if shift == 5 else 16.0 * S
        q = np.concatenate([rng.uniform(1e-3, W - 1e-3, 2_000_000),
                            (np.arange(1, int(W / S)) * S)[None, :].repeat(2000, 0).ravel() + rng.normal(0, 1e-3, 2000 * (int(W / S) - 1))]).astype(f32)
        q = q[(q > 1e-3) & (q < W - 1e-3)]
        t = (q.astype(f64) * (1.0 / S) + 192.0).astype(f32)   # fma: exact product, one rounding
        tile = (t.view(np.uint32) >> 16).astype(np.int64) - 0x4340
        assert tile.min() >= 0 and tile.max() <= int(np.ceil(W / S)) - 1 + 1
        # the tile's pixel range, widened by 1 px, contains q
        assert np.all((q.astype(f64) >= tile * S - 1.0) & (q.astype(f64) < (tile + 1) * S + 1.0))
        # in fact it is off by at most 2^-16 tiles
        assert np.all(np.abs((tile + 0.5) * S - q.astype(f64)) <= 0.5 * S + S * 2.0 ** -16)


def test_line_piece_start_closed_form():
    """State of the grid_map LineIterator after k0 steps, as k_sweep_walk computes it for a piece."""
    rng = np.random.default_rng(3)
    for _ in range(400):
        sx, sy, ex, ey = (int(v) for v in rng.integers(-300, 300, 4))
        dx, dy = abs(ex - sx), abs(ey - sy)
        den, add = max(dx, dy), min(dx, dy)
        if den == 0:
            continue
        line = orc.bresenham_cells(sx, sy, ex, ey)
        xmajor = dx >= dy
        for k0 in (0, 1, den // 3, den - 1, int(rng.integers(0, den))):
            acc = den // 2 + k0 * add
            q, num = acc // den, acc % den
            smaj = (1 if ex >= sx else -1) if xmajor else (1 if ey >= sy else -1)
            smin = (1 if ey >= sy else -1) if xmajor else (1 if ex >= sx else -1)
            major = (sx if xmajor else sy) + smaj * k0
            minor = (sy if xmajor else sx) + smin * q
            cell = (major, minor) if xmajor else (minor, major)
            assert tuple(line[k0]) == cell
            # and the iteration continues identically from (num, cell)
            n2, mn = num, minor
            for k in range(k0, min(den, k0 + 5)):
                exp = line[k]
                got = ((sx if xmajor else sy) + smaj * k, mn) if xmajor else (mn, (sx if xmajor else sy) + smaj * k)
                assert tuple(exp) == got
                n2 += add
                if n2 >= den:
                    n2 -= den
                    mn += smin


def test_index_margin_budget():
    """16.16 index: the fixed-point value is within 1.03 units of the exact coordinate, so a margin
    of 2 units certifies.  Emulated in double for the maps of BASELINE configs 3 and 5."""
    rng = np.random.default_rng(4)
    for n_cells, res, reach in ((2048, 0.1, 1204), (8192, 0.05, 2404), (1000, 0.05, 0)):
        half = 0.5 * n_cells * res
        bias = reach if reach else 16
        magic = 1.5 * 2.0 ** 36
        C = f64(half / res + bias + magic)
        nires = f64(-1.0 / res)
        pad = min(3.0, (bias - 2) * res)   # stay where the biased coordinate is positive (the kernel checks the high word)
        p = rng.uniform(-half - pad, half + pad, 1_000_000).astype(f32)
        # fma in double: the product and the sum carried in extended precision, one rounding to double
        ld = np.longdouble
        r = (p.astype(ld) * ld(nires) + ld(C)).astype(f64)
        k = (r.view(np.uint64) & 0xFFFFFFFF).astype(np.int64) - (bias << 16)
        exact = (half - p.astype(ld)) / ld(res)                # index coordinate, extended precision
        err_units = np.abs(k.astype(ld) - exact * 65536.0)
        assert float(err_units.max()) < 1.03
        # a fraction in [2, 65534) therefore never disagrees with floor(exact)
        frac = k & 0xFFFF
        ok = (frac >= 2) & (frac < 65534)
        assert np.array_equal((k[ok] >> 16), np.floor(exact[ok]).astype(np.int64))


def test_pair_label_walk_certified_labels_are_the_oracles():
    """numpy emulation of k_points_pair's label decision (binary32 projection with the kernel's
    operation order, tile from the float bits, 1-px dilated tile masks, candidate walk with the
    constant margin): whenever the emulation does not defer, its label is the oracle's, and it
    defers rarely.  Checked on full C1 scans with integer-valued and with fractional boxes."""
    from grid_vision_b200 import synth
    wl = synth.C1
    S, shift = 32.0, 5
    tiles_x = (wl.image_w + 31) >> shift
    W, H = float(wl.image_w), float(wl.image_h)
    e6, e0u, half_w, eu, ain_u, aout_u = pair_thresholds(W, wl.cx)
    _, e0v, half_h, ev, ain_v, aout_v = pair_thresholds(H, wl.cy)
    Tc = synth.camera_extrinsics(1)[0]
    rng = np.random.default_rng(11)
    deferred = total_in = 0
    for frame, fractional in ((0, False), (1, True), (2, False)):
        xyz = synth.make_scans(wl, frame0=frame, frames=1).numpy()
        boxes = synth.make_boxes(wl, frame=frame, n=50)
        if fractional:
            for k in ("x_min", "y_min"):
                boxes[k] += rng.uniform(0, 1, len(boxes))
            for k in ("x_max", "y_max"):
                boxes[k] -= rng.uniform(0, 1, len(boxes))
        X, Y, Z = orc.transform_points(Tc, *xyz)
        elab, _, _, _ = orc.project_label(wl.K(), wl.image_w, wl.image_h, X, Y, Z, boxes)
        with np.errstate(all="ignore"):
            fr = Z > f32(0.001)
            rz = (f32(1.0) / Z).astype(f32)                       # rcp.approx: within 1 ulp; exact is one instance
            t = (X * rz).astype(f32)
            q = (f64(f32(wl.fx)) * t.astype(f64) + f64(f32(wl.cx))).astype(f32)   # fma: one rounding
            t = (Y * rz).astype(f32)
            r = (f64(f32(wl.fy)) * t.astype(f64) + f64(f32(wl.cy))).astype(f32)
            aq, ar = np.abs((q - half_w).astype(f32)), np.abs((r - half_h).astype(f32))
        inn = fr & (aq < ain_u) & (ar < ain_v)
        defer = fr & ~inn & ~((aq > aout_u) | (ar > aout_v))
        lab = np.full(q.shape, -1, np.int64)
        # rounded float bounds (k_round_boxes) and 1-px dilated tile masks (k_box_masks, rev32 mode)
        bx0 = np.nextafter(boxes["x_min"].astype(f32), f32(np.inf), where=boxes["x_min"].astype(f32) < boxes["x_min"], out=boxes["x_min"].astype(f32))
        by0 = np.nextafter(boxes["y_min"].astype(f32), f32(np.inf), where=boxes["y_min"].astype(f32) < boxes["y_min"], out=boxes["y_min"].astype(f32))
        bx1 = np.nextafter(boxes["x_max"].astype(f32), f32(-np.inf), where=boxes["x_max"].astype(f32) > boxes["x_max"], out=boxes["x_max"].astype(f32))
        by1 = np.nextafter(boxes["y_max"].astype(f32), f32(-np.inf), where=boxes["y_max"].astype(f32) > boxes["y_max"], out=boxes["y_max"].astype(f32))
        idx = np.flatnonzero(inn)
        qi, ri = q[idx], r[idx]
        tu = ((qi.astype(f64) / S + 192.0).astype(f32).view(np.uint32) >> 16).astype(np.int64) - 0x4340
        tv = ((ri.astype(f64) / S + 192.0).astype(f32).view(np.uint32) >> 16).astype(np.int64) - 0x4340
        assert tu.min() >= 0 and tu.max() < tiles_x and tv.min() >= 0 and tv.max() < tiles_x
        x0t, y0t = tu * S, tv * S
        ql, qh = (qi - eu).astype(f32), (qi + eu).astype(f32)
        rl, rh = (ri - ev).astype(f32), (ri + ev).astype(f32)
        pend = np.ones(idx.size, bool)
        li = np.full(idx.size, -1, np.int64)
        di = np.zeros(idx.size, bool)
        for b in range(len(boxes)):       # list order = candidate order
            cand = (bx1[b] >= x0t - 1) & (bx0[b] < x0t + S + 1) & (by1[b] >= y0t - 1) & (by0[b] < y0t + S + 1)
            out = (qh < bx0[b]) | (ql > bx1[b]) | (rh < by0[b]) | (rl > by1[b])
            ins = (ql >= bx0[b]) & (qh <= bx1[b]) & (rl >= by0[b]) & (rh <= by1[b])
            # a box outside the point's (dilated) tile mask is never even looked at: it must be certainly outside
            assert not np.any(pend & ~cand & ~out)
            hit = pend & cand & ~out
            li[hit & ins] = b
            di[hit & ~ins] = True
            pend &= ~hit
        lab[idx] = li
        defer[idx] |= di
        keep = ~defer
        assert np.array_equal(lab[keep], elab[keep].astype(np.int64)), f"frame {frame}"
        deferred += int(defer.sum())
        total_in += int(inn.sum())
    assert total_in > 50_000 and deferred < 0.01 * total_in
