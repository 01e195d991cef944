"""CUDA RCS-aware radar scatter vs the CPU oracle (itself pinned to the reference's Python loop
by tests/golden/radar_ref.npz)."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


def _ulp_close(got, want, ulps=1):
    a = got.view(np.int32).astype(np.int64)
    b = want.view(np.int32).astype(np.int64)
    return int(np.abs(a - b).max()) <= ulps


@pytest.mark.parametrize("name", ["r32", "r64", "r128"])
def test_golden_cases(golden_radar, name):
    import rcbevdet_b200 as rcb
    g = golden_radar
    B, ny, nx, cin = (int(v) for v in g[f"{name}.shape"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    f, h, hf = rcb.radar_rcs_scatter(t(g[f"{name}.point_features"]), t(g[f"{name}.rcs"]),
                                     t(g[f"{name}.coors"]), B, ny, nx)
    assert np.array_equal(f.cpu().numpy(), g[f"{name}.features"])
    assert np.array_equal(hf.cpu().numpy(), g[f"{name}.heatmap_feat"])
    assert _ulp_close(h.cpu().numpy(), g[f"{name}.heatmap"])   # fp64 exp, CUDA vs numpy: <= 1 ulp after the fp32 cast


def test_config4_shape_vs_oracle_and_gradient():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B, ny, nx = 2, 128, 128
    pf, rc, co = rig.radar_pillars(B, ny, nx, points_per_sample=1500, seed=4)
    wf, wh, whf = oracle.radar_rcs_scatter(pf.numpy(), rc.numpy(), co.numpy(), B, ny, nx)
    p = pf.cuda().requires_grad_(True)
    f, h, hf = rcb.radar_rcs_scatter(p, rc.cuda(), co.cuda(), B, ny, nx)
    assert np.array_equal(f.detach().cpu().numpy(), wf)
    assert np.array_equal(hf.cpu().numpy(), whf)
    assert _ulp_close(h.cpu().numpy(), wh)
    g = torch.randn(f.shape, generator=torch.Generator().manual_seed(0))
    f.backward(g.cuda())
    want = g[co[:, 0].long(), :, co[:, 2].long(), co[:, 3].long()]
    assert torch.equal(p.grad.cpu(), want)


def test_module_matches_manual_composition():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B, ny, nx = 1, 64, 64
    pf, rc, co = rig.radar_pillars(B, ny, nx, points_per_sample=300, in_channels=16, seed=2)
    m = rcb.PointPillarsScatterRCS(16, (ny, nx)).cuda()
    out = m((pf.cuda(), rc.cuda()), co.cuda(), batch_size=B)
    assert out.shape == (B, 16, ny, nx)
    wf, wh, whf = oracle.radar_rcs_scatter(pf.numpy(), rc.numpy(), co.numpy(), B, ny, nx)
    att = m.rcs_att(torch.cat([torch.from_numpy(wh).cuda().unsqueeze(1), torch.from_numpy(whf).cuda()], 1))
    want = m.compress(torch.cat([torch.from_numpy(wf).cuda(), att], 1))
    assert torch.allclose(out, want, rtol=1e-5, atol=1e-5)


def test_empty_input_gives_zeros():
    import rcbevdet_b200 as rcb
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device="cuda")
    f, h, hf = rcb.radar_rcs_scatter(z(0, 8), z(0, 7), z(0, 4, dt=torch.int32), 2, 16, 16)
    assert f.shape == (2, 8, 16, 16) and float(f.abs().max()) == 0 and float(h.abs().max()) == 0
    assert float(hf.abs().max()) == 0


@pytest.mark.parametrize("ny", [128, 512])
def test_config4_full_size_properties_and_timing(ny, capsys):
    """BASELINE config 4 (B=8, 5 sweeps x 5 radars x 125 points per sample) on the 128x128 grid and
    on the shipped 512x512 pillar grid: exact scatter identities at full size + a timing line."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B = 8
    pf, rc, co = rig.radar_pillars(B, ny, ny, points_per_sample=3125, seed=11)
    pf, rc, co = pf.cuda(), rc.cuda(), co.cuda()
    f, h, hf = rcb.radar_rcs_scatter(pf, rc, co, B, ny, ny)
    # features: exactly the pillar rows at the pillar cells, zero elsewhere
    picked = f[co[:, 0].long(), :, co[:, 2].long(), co[:, 3].long()]
    assert torch.equal(picked, pf)
    assert int((f != 0).sum()) == int((pf != 0).sum())
    # heat-map: 1.0 exactly at every pillar centre (Gaussian peak), within [0, 1] everywhere
    assert float(h[co[:, 0].long(), co[:, 2].long(), co[:, 3].long()].min()) == 1.0
    assert float(h.min()) >= 0.0 and float(h.max()) == 1.0
    # heatmap_feat holds RCS values only where the heat-map is covered
    assert bool(((hf[:, 0] != 0) <= (h > 0)).all())
    # idempotence / determinism
    f2, h2, hf2 = rcb.radar_rcs_scatter(pf, rc, co, B, ny, ny)
    assert torch.equal(f, f2) and torch.equal(h, h2) and torch.equal(hf, hf2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        rcb.radar_rcs_scatter(pf, rc, co, B, ny, ny)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    out_bytes = B * ny * ny * (64 + 2) * 4 + pf.numel() * 4 + rc.numel() * 4 + co.numel() * 4
    with capsys.disabled():
        print(f"\n[timing] radar RCS scatter B=8 V={pf.shape[0]} grid {ny}x{ny}: {ms * 1e3:.1f} us, "
              f"{out_bytes / ms / 1e6:.0f} GB/s algorithmic ({B / ms * 1e3:.0f} samples/s)")


def _edge_case(seed, B, ny, nx, n_per_sample, rcs_hi, cin=8, drop_sample=None):
    """Pillars in arbitrary order (samples interleaved), radii up to relu(rcs_hi * 2) + 1, unique cells
    per sample."""
    g = torch.Generator().manual_seed(seed)
    feats, rcss, coors = [], [], []
    for b in range(B):
        if b == drop_sample:
            continue
        lin = torch.randperm(ny * nx, generator=g)[:n_per_sample]
        r = torch.zeros(n_per_sample, 7)
        r[:, 0] = torch.rand(n_per_sample, generator=g)
        r[:, 1] = torch.rand(n_per_sample, generator=g)
        r[:, 5] = torch.rand(n_per_sample, generator=g) * (rcs_hi + 5.0) - 5.0
        c = torch.zeros(n_per_sample, 4, dtype=torch.int32)
        c[:, 0] = b
        c[:, 2] = (lin // nx).int()
        c[:, 3] = (lin % nx).int()
        feats.append(torch.randn(n_per_sample, cin, generator=g))
        rcss.append(r)
        coors.append(c)
    pf, rc, co = torch.cat(feats), torch.cat(rcss), torch.cat(coors)
    perm = torch.randperm(pf.shape[0], generator=g)          # samples interleaved
    return pf[perm].contiguous(), rc[perm].contiguous(), co[perm].contiguous()


@pytest.mark.parametrize("B,ny,nx,n,rcs_hi,drop", [
    (3, 100, 176, 60, 30.0, None),     # ragged tiles (nx, ny not multiples of 32 / 8), radii <= 61
    (2, 200, 176, 40, 120.0, None),    # radii up to ~240: beyond the table (>= 64), windows larger than the grid
    (4, 64, 64, 50, 30.0, 2),          # a sample without pillars
    (1, 40, 520, 30, 60.0, None),      # wide grid: several rows per thread at small batch
])
def test_tile_form_edge_cases_against_oracle(B, ny, nx, n, rcs_hi, drop):
    """The tile (gather) kernels on what the goldens do not cover: pillars in arbitrary order, radii
    beyond the tabulated range, an empty sample, ragged tile borders.  features / heatmap_feat
    bit-exact, heatmap <= 1 ulp (fp32)."""
    import rcbevdet_b200 as rcb
    pf, rc, co = _edge_case(7 + B, B, ny, nx, n, rcs_hi, drop_sample=drop)
    wf, wh, whf = oracle.radar_rcs_scatter(pf.numpy(), rc.numpy(), co.numpy(), B, ny, nx)
    f, h, hf = rcb.radar_rcs_scatter(pf.cuda(), rc.cuda(), co.cuda(), B, ny, nx)
    assert np.array_equal(f.cpu().numpy(), wf)
    assert np.array_equal(hf.cpu().numpy(), whf)
    assert _ulp_close(h.cpu().numpy(), wh)
    if drop is not None:
        assert float(h[drop].abs().max()) == 0.0 and float(f[drop].abs().max()) == 0.0


def test_tile_and_splat_forms_agree(monkeypatch):
    """RCB_RADAR_SPLAT=1 selects the scatter kernels (kept for reference measurements): same
    features / heatmap_feat bits, heat-map within 1 ulp of the tile form."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    pf, rc, co = (t.cuda() for t in rig.radar_pillars(2, 128, 128, points_per_sample=800, seed=9))
    a = rcb.radar_rcs_scatter(pf, rc, co, 2, 128, 128)
    monkeypatch.setenv("RCB_RADAR_SPLAT", "1")
    b = rcb.radar_rcs_scatter(pf, rc, co, 2, 128, 128)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    assert _ulp_close(a[1].cpu().numpy(), b[1].cpu().numpy())


@pytest.mark.parametrize("ny,nx,B", [(100, 176, 3), (33, 41, 2), (8, 1000, 1), (300, 20, 2)])
def test_tile_form_writes_inside_its_buffers(ny, nx, B):
    """Guard bands around every output and the workspace of the C-ABI call (ragged tiles, grids narrower
    than a tile, several rows per thread): nothing outside the declared extents is written."""
    import ctypes
    from rcbevdet_b200 import _lib
    from rcbevdet_b200.radar import _desc
    pf, rc, co = _edge_case(3, B, ny, nx, min(40, ny * nx // 4), 40.0, cin=8)
    pf, rc, co = pf.cuda(), rc.cuda(), co.cuda()
    d = _desc(pf, rc, B, ny, nx)
    lib = _lib.lib()
    G = 1024                                                # guard elements on each side (16-byte aligned)

    def guarded(n, dtype=torch.float32):
        t = torch.full((n + 2 * G,), -123.0 if dtype == torch.float32 else 77, dtype=dtype, device="cuda")
        return t, t[G:G + n]

    cells = B * ny * nx
    fa, f = guarded(cells * 8)
    ha, h = guarded(cells)
    hfa, hf = guarded(cells)
    ws_bytes = lib.rcb_radar_workspace_bytes(ctypes.byref(d))
    wa, ws = guarded(ws_bytes, torch.uint8)
    _lib.check(lib.rcb_radar_rcs_scatter(ctypes.byref(d), _lib.ptr(pf), _lib.ptr(rc), _lib.ptr(co), _lib.ptr(f), _lib.ptr(h),
                                         _lib.ptr(hf), _lib.ptr(ws), ws_bytes, 0, None), "radar")
    torch.cuda.synchronize()
    for name, full, n in (("features", fa, cells * 8), ("heatmap", ha, cells), ("heatmap_feat", hfa, cells)):
        assert float((full[:G] + 123.0).abs().max()) == 0.0 and float((full[G + n:] + 123.0).abs().max()) == 0.0, name
    assert int((wa[:G] != 77).sum()) == 0 and int((wa[G + ws_bytes:] != 77).sum()) == 0, "workspace"
    wf, wh, whf = oracle.radar_rcs_scatter(pf.cpu().numpy(), rc.cpu().numpy(), co.cpu().numpy(), B, ny, nx)
    assert np.array_equal(f.view(B, 8, ny, nx).cpu().numpy(), wf)
    assert np.array_equal(hf.view(B, 1, ny, nx).cpu().numpy(), whf)
    assert _ulp_close(h.view(B, ny, nx).cpu().numpy(), wh)
