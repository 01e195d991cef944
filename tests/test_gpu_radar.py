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
