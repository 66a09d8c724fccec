"""BASELINE.json full-size runs: configs[1] 100 M points x 10 k polygons (EVERY id compared with the oracle, in
chunks), configs[2] 50 M LineString pairs (10 M rows against the oracle), configs[4] 10 M 256-vertex polygons (1 M
hull rings against the oracle), plus the size-independent properties (disjointness, symmetry, idempotence, rigid
motions).  Data is generated on the device (geopolars_b200/csrc/synth.cu), bit-identical to the host generators
where no libm call is involved."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_close
from geopolars_b200 import GeoArrowArray, GeometryType, synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def tctx():
    from geopolars_b200 import engine as E

    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ctx = E.Context(0, st.cuda_stream)
    yield ctx, st
    ctx.synchronize()


def test_config2_full_size_contains_join(tctx, og, conv):
    from geopolars_b200 import engine as E

    ctx, st = tctx
    n, m = 100_000_000, 10_000
    dev = torch.device("cuda", 0)
    with torch.cuda.stream(st):
        pts = torch.empty((n, 2), dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_gen_uniform_points(ctx._h, 2, 0, n, 1000.0, pts.data_ptr()))
        xy, ro, go = synth.star_polygons(m, 100)
        polys = GeoArrowArray.polygons(xy, ro, go)
        idx = E.PipIndex(ctx.upload(polys))
        ids = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        idx.query_device(pts.data_ptr(), n, ids.data_ptr(), cnt.data_ptr())
        st.synchronize()
        # range, disjointness (each point in at most one polygon), hit rate of the generator
        assert int(ids.min()) >= -1 and int(ids.max()) < m
        assert int(cnt.max()) <= 1 and bool(((cnt == 1) == (ids >= 0)).all())
        hits = int((ids >= 0).sum())
        assert 0.35 < hits / n < 0.38
        # a hit lies in the lattice cell of its polygon (polygon j sits in cell j of the 100 x 100 lattice)
        hit = ids >= 0
        cell = (pts[:, 1] / 10.0).floor().long() * 100 + (pts[:, 0] / 10.0).floor().long()
        assert bool((cell[hit] == ids[hit].long()).all())
        # histogram of ids == per-polygon counts, sums to the hit count
        counts = torch.zeros(m, dtype=torch.int64, device=dev)
        E.check(ctx.lib.gpl_join_histogram(ctx._h, ids.data_ptr(), n, counts.data_ptr(), m, E.GPL_DEVICE))
        st.synchronize()
        assert int(counts.sum()) == hits
        assert torch.equal(counts, torch.bincount(ids[hit].long(), minlength=m))
        # idempotence + independence of the launch partitioning: second half alone gives the same ids
        ids2 = torch.empty(n // 2, dtype=torch.int32, device=dev)
        idx.query_device(pts[n // 2 :].data_ptr(), n // 2, ids2.data_ptr())
        st.synchronize()
        assert torch.equal(ids2, ids[n // 2 :])
        # EVERY id against the oracle (device generator == host generator bit for bit), 10 M rows at a time
        k = 300_000
        assert np.array_equal(pts[:k].cpu().numpy(), synth.uniform_points(k))
        chunk = 10_000_000
        opolys = conv(polys)
        for lo in range(0, n, chunk):
            hp = og.gen_uniform_points(2, lo, chunk, 1000.0)
            if lo in (0, 50_000_000):
                assert np.array_equal(pts[lo : lo + chunk].cpu().numpy(), hp)
            want, _ = og.contains_join(opolys, hp, use_grid=True, threads=0)
            assert np.array_equal(ids[lo : lo + chunk].cpu().numpy(), want), f"ids differ from the oracle in rows [{lo}, {lo + chunk})"
        # the exact kernel re-evaluated the points the FP32 filter could not certify: a small, non-zero share
        deferred = idx.stats()["deferred"]
        assert 0 < deferred < 0.01 * n * 2, deferred  # two full queries ran on this index (and one half)
        # fused per-polygon counts == histogram of the id column
        fused = torch.zeros(m, dtype=torch.int64, device=dev)
        idx.query_device_counts(pts.data_ptr(), n, ids.data_ptr(), fused.data_ptr())
        st.synchronize()
        assert torch.equal(fused, counts)
        # end-to-end host path gives the same column
        host_pts = torch.empty((4_000_000, 2), dtype=torch.float64, pin_memory=True)
        host_pts.copy_(pts[: host_pts.shape[0]])
        host_ids = torch.empty(host_pts.shape[0], dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()
        idx.query_host_pipelined(host_pts.data_ptr(), host_pts.shape[0], host_ids.data_ptr(), 1_000_000)
        assert torch.equal(host_ids, ids[: host_pts.shape[0]].cpu())


def test_config3_full_size_linestring_pairs(tctx, og, conv):
    from geopolars_b200 import engine as E

    ctx, st = tctx
    n, k = 50_000_000, 16
    dev = torch.device("cuda", 0)
    with torch.cuda.stream(st):
        axy = torch.empty((n * k, 2), dtype=torch.float64, device=dev)
        bxy = torch.empty((n * k, 2), dtype=torch.float64, device=dev)
        aoff = torch.empty(n + 1, dtype=torch.int64, device=dev)
        boff = torch.empty(n + 1, dtype=torch.int64, device=dev)
        E.check(ctx.lib.gpl_gen_walk_linestrings(ctx._h, 3, -1, 0, n, k, axy.data_ptr(), aoff.data_ptr()))
        E.check(ctx.lib.gpl_gen_walk_linestrings(ctx._h, 4, 3, 0, n, k, bxy.data_ptr(), boff.data_ptr()))
        A = ctx.wrap_device(GeometryType.LINESTRING, n, n * k, axy.data_ptr(), geom_off_ptr=aoff.data_ptr(), keepalive=(axy, aoff))
        B = ctx.wrap_device(GeometryType.LINESTRING, n, n * k, bxy.data_ptr(), geom_off_ptr=boff.data_ptr(), keepalive=(bxy, boff))
        nb = (n + 7) // 8
        i_ab = torch.empty(nb, dtype=torch.uint8, device=dev)
        i_ba = torch.empty(nb, dtype=torch.uint8, device=dev)
        d_ab = torch.empty(n, dtype=torch.float64, device=dev)
        d_ba = torch.empty(n, dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_intersects(ctx._h, A._h, B._h, C.c_void_p(i_ab.data_ptr()), E.GPL_DEVICE))
        E.check(ctx.lib.gpl_intersects(ctx._h, B._h, A._h, C.c_void_p(i_ba.data_ptr()), E.GPL_DEVICE))
        E.check(ctx.lib.gpl_distance(ctx._h, A._h, B._h, C.c_void_p(d_ab.data_ptr()), None, E.GPL_DEVICE))
        E.check(ctx.lib.gpl_distance(ctx._h, B._h, A._h, C.c_void_p(d_ba.data_ptr()), None, E.GPL_DEVICE))
        st.synchronize()
        assert torch.equal(i_ab, i_ba)  # the predicate is symmetric, exactly
        bits = ((i_ab[:, None] >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).reshape(-1)[:n].bool()
        assert torch.equal(bits, d_ab == 0.0)  # distance is 0 exactly on the intersecting rows
        assert 0.3 < float(bits.float().mean()) < 0.7
        assert bool((d_ab >= 0).all()) and bool(torch.isfinite(d_ab).all())
        assert bool(((d_ab - d_ba).abs() <= 1e-12 * torch.maximum(d_ab, d_ba)).all())  # symmetric up to rounding
        # the distance never exceeds the distance between the first vertices
        first = (axy[::k] - bxy[::k]).norm(dim=1)
        assert bool((d_ab <= first * (1 + 1e-12)).all())
        # 10 M rows against the oracle (two 5 M-row blocks from both ends of the column)
        m = 5_000_000
        for lo in (0, n - m):
            ah, _ = synth.walk_linestrings(m, k, first=lo, stream=3)
            bh, _ = synth.walk_linestrings(m, k, first=lo, stream=4, other_of=3)
            assert np.array_equal(axy[lo * k : (lo + m) * k].cpu().numpy(), ah) and np.array_equal(bxy[lo * k : (lo + m) * k].cpu().numpy(), bh)
            off = np.arange(m + 1) * k
            HA, HB = GeoArrowArray.linestrings(ah, off), GeoArrowArray.linestrings(bh, off)
            assert np.array_equal(bits[lo : lo + m].cpu().numpy(), og.intersects_rowwise(conv(HA), conv(HB), threads=0))
            assert rel_close(d_ab[lo : lo + m].cpu().numpy(), og.distance_rowwise(conv(HA), conv(HB), threads=0), 1e-9)


def test_config5_full_size_polygons(tctx, og, conv):
    from geopolars_b200 import engine as E

    ctx, st = tctx
    g, nv = 10_000_000, 256
    nc = g * (nv + 1)
    dev = torch.device("cuda", 0)
    with torch.cuda.stream(st):
        xy = torch.empty((nc, 2), dtype=torch.float64, device=dev)
        ro = torch.empty(g + 1, dtype=torch.int64, device=dev)
        go = torch.empty(g + 1, dtype=torch.int64, device=dev)
        E.check(ctx.lib.gpl_gen_blob_polygons(ctx._h, 5, 0, g, nv, xy.data_ptr(), ro.data_ptr(), go.data_ptr()))
        polys = ctx.wrap_device(GeometryType.POLYGON, g, nc, xy.data_ptr(), geom_off_ptr=go.data_ptr(), ring_off_ptr=ro.data_ptr(), n_rings=g,
                                keepalive=(xy, ro, go))
        area = torch.empty(g, dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_area(ctx._h, polys._h, C.c_void_p(area.data_ptr()), E.GPL_DEVICE))
        # star-shaped blobs with radii in [0.5, 1]: area between the inscribed and circumscribed discs
        st.synchronize()
        assert float(area.min()) > 0.7 and float(area.max()) < 3.2
        # rigid motion (BASELINE config 5's matrix is a rotation + translation): area preserved, centroid mapped
        m = (0.8, -0.6, 10.0, 0.6, 0.8, -5.0)
        moved = E.affine_transform(polys, m)
        area2 = torch.empty(g, dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_area(ctx._h, moved._h, C.c_void_p(area2.data_ptr()), E.GPL_DEVICE))
        c0 = E.centroid(polys)
        c1 = E.centroid(moved)
        v0, v1 = c0.view(), c1.view()
        t0 = torch.empty((g, 2), dtype=torch.float64, device=dev)
        t1 = torch.empty((g, 2), dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_array_copy_out(ctx._h, c0._h, C.c_void_p(t0.data_ptr()), None, None, None, None, E.GPL_DEVICE))
        E.check(ctx.lib.gpl_array_copy_out(ctx._h, c1._h, C.c_void_p(t1.data_ptr()), None, None, None, None, E.GPL_DEVICE))
        st.synchronize()
        assert v0.n_geoms == g and v1.n_geoms == g
        assert bool(((area - area2).abs() <= 1e-9 * area).all())
        want_x = 0.8 * t0[:, 0] - 0.6 * t0[:, 1] + 10.0
        want_y = 0.6 * t0[:, 0] + 0.8 * t0[:, 1] - 5.0
        assert bool(((t1[:, 0] - want_x).abs() <= 1e-9 * want_x.abs().clamp(min=1.0)).all())
        assert bool(((t1[:, 1] - want_y).abs() <= 1e-9 * want_y.abs().clamp(min=1.0)).all())
        del moved, c0, c1, t0, t1, area2
        ctx.trim()
        # inverse transform returns the coordinates to 1e-9 (on a 1 M-polygon prefix: memory)
        g1 = 1_000_000
        sub = ctx.wrap_device(GeometryType.POLYGON, g1, g1 * (nv + 1), xy.data_ptr(), geom_off_ptr=go.data_ptr(), ring_off_ptr=ro.data_ptr(),
                              n_rings=g1, keepalive=(xy, ro, go))
        back = E.affine_transform(E.affine_transform(sub, m), (0.8, 0.6, -(0.8 * 10.0 + 0.6 * -5.0), -0.6, 0.8, -(-0.6 * 10.0 + 0.8 * -5.0)))
        bxy = torch.empty((g1 * (nv + 1), 2), dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_array_copy_out(ctx._h, back._h, C.c_void_p(bxy.data_ptr()), None, None, None, None, E.GPL_DEVICE))
        st.synchronize()
        bxy.sub_(xy[: g1 * (nv + 1)])
        assert float(bxy.abs_().max()) < 1e-9 * 1010.0
        del back, sub, bxy
        ctx.trim()
        # convex hull: closed rings of input vertices, hull area >= polygon area, hull(hull) == hull
        hull = E.convex_hull(polys)
        hv = hull.view()
        hro = torch.empty(g + 1, dtype=torch.int64, device=dev)
        hxy = torch.empty((hv.n_coords, 2), dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_array_copy_out(ctx._h, hull._h, C.c_void_p(hxy.data_ptr()), None, None, C.c_void_p(hro.data_ptr()), None, E.GPL_DEVICE))
        harea = torch.empty(g, dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_area(ctx._h, hull._h, C.c_void_p(harea.data_ptr()), E.GPL_DEVICE))
        st.synchronize()
        sizes = hro[1:] - hro[:-1]
        assert int(sizes.min()) >= 4 and int(sizes.max()) <= nv + 1
        assert torch.equal(hxy[hro[:-1]], hxy[hro[1:] - 1])  # closed
        assert bool((harea >= area * (1 - 1e-12)).all())
        hull2 = E.convex_hull(hull)
        h2xy = torch.empty((hull2.view().n_coords, 2), dtype=torch.float64, device=dev)
        E.check(ctx.lib.gpl_array_copy_out(ctx._h, hull2._h, C.c_void_p(h2xy.data_ptr()), None, None, None, None, E.GPL_DEVICE))
        st.synchronize()
        assert hull2.view().n_coords == hv.n_coords and torch.equal(h2xy, hxy)  # idempotent, same vertex order
        # 1 M hull rings against the oracle, bit for bit (vertex set AND geo's quick_hull order); the device generator uses
        # CUDA sincos, so the oracle gets the device's coordinates
        g1 = 1_000_000
        hx = xy[: g1 * (nv + 1)].cpu().numpy()
        harr = og.OGArray(og.POLYGON, hx, geom_off=np.arange(g1 + 1, dtype=np.int64), ring_off=np.arange(g1 + 1, dtype=np.int64) * (nv + 1))
        want_off, want_xy = og.convex_hull(harr, threads=0)
        assert np.array_equal(hro[: g1 + 1].cpu().numpy(), want_off)
        assert np.array_equal(hxy[: int(want_off[-1])].cpu().numpy(), want_xy)
