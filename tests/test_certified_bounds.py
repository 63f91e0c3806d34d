"""CPU emulation of the two certified fast paths of k_points (csrc/gv_kernels.cuh:
fuse_point_fast and grid_get_index_cert).  The GPU tests check them against the oracle on the
device; these tests check the ERROR ANALYSIS itself on millions of samples, every round, without
a GPU: (1) the binary32 projection stays within a third of the margin E the kernel uses, even
with the division perturbed by the 2 ulp __fdividef is allowed; (2) whenever the kernel's
decision rules certify an answer, it is the exact one; (3) few points defer."""
import numpy as np

from oracle import gv_oracle as orc

f32, f64 = np.float32, np.float64
K = np.array([[208.0, 0, 208.0], [0, 208.0, 208.0], [0, 0, 1.0]])
FX = CX = f32(208.0)
E6 = f32(6.0) * f32(2.0 ** -22)
E0 = f32(2.0 ** -22) * (abs(CX) + f32(1.0)) * f32(1.0001)


def fast_q(X, Z, ulps=0):
    """q = fdividef(fmaf(fx, X, cx*Z), Z) in binary32; `ulps` perturbs the quotient."""
    t = (CX * Z).astype(f32)
    s = (FX.astype(f64) * X.astype(f64) + t.astype(f64)).astype(f32)   # fma: one rounding
    q = (s / Z).astype(f32)
    for _ in range(abs(ulps)):
        q = np.nextafter(q, f32(np.inf if ulps > 0 else -np.inf))
    return q


def test_projection_error_bound_holds_with_margin():
    rng = np.random.default_rng(0)
    n = 2_000_000
    Z = np.exp(rng.uniform(np.log(0.0011), np.log(300.0), n)).astype(f32)
    u_t = rng.uniform(-600, 1000, n)                                   # in and far outside the image
    X = (Z * ((u_t - 208.0) / 208.0)).astype(f32)
    Y = np.zeros(n, f32)
    _, _, u_ref, _ = orc.project_label(K, 416, 416, X, Y, Z, np.zeros(0, orc.BOX_DTYPE))
    worst = 0.0
    for ulps in (-2, 0, 2):
        q = fast_q(X, Z, ulps)
        E = (np.abs(q) * E6 + E0).astype(f32)
        ratio = np.abs(q.astype(f64) - u_ref.astype(f64)) / E.astype(f64)
        worst = max(worst, float(ratio.max()))
    assert worst < 1.0 / 3.0, worst      # the kernel's margin is > 3x the realised error


def certified_label(X, Y, Z, boxes, ulps):
    """Decision rules of fuse_point_fast; returns (label, certified)."""
    q, r = fast_q(X, Z, ulps), fast_q(Y, Z, -ulps)
    Eu, Ev = (np.abs(q) * E6 + E0).astype(f32), (np.abs(r) * E6 + E0).astype(f32)
    ql, qh, rl, rh = (q - Eu).astype(f32), (q + Eu).astype(f32), (r - Ev).astype(f32), (r + Ev).astype(f32)
    W = f32(416.0)
    inside = (ql >= 0) & (qh < W) & (rl >= 0) & (rh < W)
    outside = (qh < 0) | (ql >= W) | (rh < 0) | (rl >= W)
    label = np.full(X.size, -1, np.int32)
    cert = outside.copy()                                 # certainly outside -> label -1
    live = inside.copy()
    tile_ok = (ql.astype(np.int32) >> 5 == qh.astype(np.int32) >> 5) & (rl.astype(np.int32) >> 5 == rh.astype(np.int32) >> 5)
    live &= tile_ok                                        # straddling a tile edge defers
    undecided = live.copy()
    for b in range(len(boxes)):
        x0, y0 = f32(np.ceil(boxes["x_min"][b])), f32(np.ceil(boxes["y_min"][b]))   # integer bounds: RU/RD exact
        x1, y1 = f32(np.floor(boxes["x_max"][b])), f32(np.floor(boxes["y_max"][b]))
        cin = (ql >= x0) & (qh <= x1) & (rl >= y0) & (rh <= y1)
        cout = (qh < x0) | (ql > x1) | (rh < y0) | (rl > y1)
        hit = undecided & cin
        label[hit] = b
        cert |= hit
        amb = undecided & ~cin & ~cout
        undecided &= ~(cin | amb)                          # ambiguous -> deferred, stays uncertified
    cert |= undecided                                      # every box certainly missed -> label -1
    return label, cert


def test_certified_decisions_equal_exact_ones():
    rng = np.random.default_rng(1)
    nb = 40
    w, h = rng.integers(16, 129, nb), rng.integers(16, 129, nb)
    x0, y0 = rng.integers(0, 416 - w), rng.integers(0, 416 - h)
    boxes = orc.make_boxes(np.stack([x0, y0, x0 + w, y0 + h], 1).astype(f64))
    edges = np.unique(np.concatenate([x0, x0 + w, y0, y0 + h, np.arange(0, 417, 32)])).astype(f64)
    n = 400_000
    Z = np.exp(rng.uniform(np.log(0.5), np.log(120.0), n)).astype(f32)
    # half the points aimed at decision boundaries (+- a few float steps), half anywhere
    tgt_u = np.where(rng.random(n) < 0.5, rng.choice(edges, n) + rng.integers(-3, 4, n) * 3e-5, rng.uniform(-40, 460, n))
    tgt_v = np.where(rng.random(n) < 0.5, rng.choice(edges, n) + rng.integers(-3, 4, n) * 3e-5, rng.uniform(-40, 460, n))
    X = (Z * ((tgt_u - 208.0) / 208.0)).astype(f32)
    Y = (Z * ((tgt_v - 208.0) / 208.0)).astype(f32)
    exact, _, _, _ = orc.project_label(K, 416, 416, X, Y, Z, boxes)
    deferred = 0
    for ulps in (-2, 0, 2):
        lab, cert = certified_label(X, Y, Z, boxes, ulps)
        assert np.array_equal(lab[cert], exact[cert].astype(np.int32))
        deferred = max(deferred, int((~cert).sum()))
    assert deferred < 0.25 * n          # adversarial set: many defer, none is wrong
    # on ordinary points almost nothing defers
    X = (Z * ((rng.uniform(0, 416, n) - 208.0) / 208.0)).astype(f32)
    Y = (Z * ((rng.uniform(0, 416, n) - 208.0) / 208.0)).astype(f32)
    exact, _, _, _ = orc.project_label(K, 416, 416, X, Y, Z, boxes)
    lab, cert = certified_label(X, Y, Z, boxes, 0)
    assert np.array_equal(lab[cert], exact[cert].astype(np.int32))
    assert (~cert).mean() < 0.02


def test_fixed_point_index_certification():
    rng = np.random.default_rng(2)
    for nx, ny, res, px, py in [(2048, 2048, 0.1, 0.0, 0.0), (500, 200, 0.1, 16.0, 0.0), (8192, 8192, 0.05, 3.3, -7.1)]:
        g = orc.Grid.from_cells(nx, ny, res, px, py)
        c0x, c0y, mres = 0.5 * g.len_x + px, 0.5 * g.len_y + py, 65536.0 / res
        n = 60_000
        i = rng.integers(-2, nx + 3, n)
        bx = (c0x - i * res).astype(f32)
        bx = (bx + rng.integers(-3, 4, n).astype(f32) * np.spacing(bx)).astype(f32)   # on cell boundaries
        bx = np.concatenate([bx, rng.uniform(px - g.len_x * 0.6, px + g.len_x * 0.6, n).astype(f32)])
        by = rng.uniform(py - g.len_y * 0.6, py + g.len_y * 0.6, bx.size).astype(f32)
        with np.errstate(invalid="ignore"):
            kx = np.trunc((c0x - bx.astype(f64)) * mres).astype(np.int64)
            ky = np.trunc((c0y - by.astype(f64)) * mres).astype(np.int64)
        fx, fy = kx & 0xFFFF, ky & 0xFFFF
        frac_ok = (fx >= 8) & (fx <= 0xFFFF - 8) & (fy >= 8) & (fy <= 0xFFFF - 8)
        klx, kly = nx << 16, ny << 16
        cert_in = frac_ok & (kx >= 8) & (ky >= 8) & (kx < klx - 8) & (ky < kly - 8)
        cert_out = (kx < -8) | (ky < -8) | (kx >= klx + 8) | (ky >= kly + 8)
        exact = [g.get_index(float(a), float(b)) for a, b in zip(bx, by)]
        for j in np.flatnonzero(cert_in):
            assert exact[j] == (int(kx[j] >> 16), int(ky[j] >> 16)), (nx, j)
        for j in np.flatnonzero(cert_out):
            assert exact[j] is None, (nx, j)
        rand = slice(n, None)
        assert (~(cert_in | cert_out))[rand].mean() < 1e-3      # ordinary points: almost never defer
