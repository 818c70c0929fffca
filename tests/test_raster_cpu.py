"""CPU checks of the rasteriser oracle's helpers (test infrastructure for the -m gpu parity tests)."""
import math

import numpy as np
import torch

from oracle import gsplat_ref as gr


def _small_scene(n=3000, seed=2, H=96, W=128):
    g = torch.Generator().manual_seed(seed)
    means = torch.rand((n, 3), generator=g) * 2 - 1
    scales = torch.exp(torch.randn((n, 3), generator=g) * 0.4 + math.log(0.04))
    quats = torch.randn((n, 4), generator=g)
    quats = quats / quats.norm(dim=-1, keepdim=True)
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([0.1, -0.2, 2.5])
    fx = fy = 110.0
    vm = gr.viewmat_from_c2w(c2w)
    pm = gr.projection_matrix(0.001, 1000, 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy)))
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    out = gr.project_gaussians(means, scales, 1, quats, vm[:3], pm @ vm, fx, fy, W / 2, H / 2, H, W, tb)
    return out, tb, (H, W), g


def test_vectorized_binning_equals_loop():
    (xys, depths, radii, conics, nth, _), tb, _, _ = _small_scene()
    k1, g1, b1 = gr.bin_and_sort(xys, depths, radii, nth, tb)
    k2, g2, b2 = gr.bin_and_sort_vectorized(xys, depths, radii, nth, tb)
    assert len(k1) > 1000
    assert np.array_equal(k1, k2) and np.array_equal(g1, g2) and np.array_equal(b1, b2)


def test_rasterize_sorted_accepts_tensor_bins_and_background_device():
    (xys, depths, radii, conics, nth, _), tb, (H, W), g = _small_scene(n=800)
    _, gids, bins = gr.bin_and_sort(xys, depths, radii, nth, tb)
    col = torch.rand((xys.shape[0], 3), generator=g)
    op = torch.rand((xys.shape[0],), generator=g)
    a = gr.rasterize_sorted(xys, conics, col, op, gids, bins, H, W, torch.tensor([0.1, 0.2, 0.3]))
    b = gr.rasterize_sorted(xys, conics, col, op, gids, torch.from_numpy(bins), H, W, torch.tensor([0.1, 0.2, 0.3]))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_sh_basis_is_the_real_spherical_harmonics_of_scipy():
    """Pin of the oracle's spherical-harmonics evaluation (gsplat 0.1.3 sh_coeffs_to_color, gc_model.py:166) against an
    independent implementation of the published functions: with one-hot coefficients the oracle returns basis function
    k = l^2 + l + m, which must equal scipy's spherical harmonic Y_l^m (Condon-Shortley phase included) combined the
    3D-Gaussian-splatting way: Y_l^0 for m = 0, sqrt(2) Re Y_l^m for m > 0, sqrt(2) Im Y_l^|m| for m < 0.  Degrees 0..3
    (all 16 coefficients splatfacto uses), fp64, to 1e-12."""
    import numpy as np
    import torch
    from scipy import special
    from oracle import gsplat_ref as g
    rng = np.random.default_rng(0)
    d = rng.normal(size=(257, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    theta, phi = np.arccos(d[:, 2]), np.arctan2(d[:, 1], d[:, 0])   # polar angle, azimuth

    def ylm(l, m):
        if hasattr(special, "sph_harm_y"):
            return special.sph_harm_y(l, m, theta, phi)
        return special.sph_harm(m, l, phi, theta)

    for l in range(4):
        for m in range(-l, l + 1):
            k = l * l + l + m
            co = torch.zeros((d.shape[0], 16, 3), dtype=torch.float64)
            co[:, k, :] = 1.0
            got = g.spherical_harmonics(3, torch.tensor(d), co).numpy()
            want = ylm(l, 0).real if m == 0 else (np.sqrt(2) * ylm(l, m).real if m > 0 else np.sqrt(2) * ylm(l, -m).imag)
            assert np.abs(got - want[:, None]).max() < 1e-12, (l, m)
            # lower degrees ignore the higher coefficients
            for deg in range(l):
                assert np.abs(g.spherical_harmonics(deg, torch.tensor(d), co).numpy()).max() == 0.0


def test_projection_oracle_against_autograd_jacobian_and_scipy_rotations():
    """Pin of the oracle's project_gaussians (gsplat 0.1.3 project_gaussians_forward, gc_model.py:140-154) against an
    independent derivation of the published EWA-splatting math: 3-D covariance R S S^T R^T with scipy's quaternion ->
    rotation matrix, camera-space covariance through the view rotation, 2-D covariance J Sigma J^T with J the AUTOGRAD
    Jacobian of the pinhole projection (the oracle uses the closed-form J), + the 0.3 px low-pass; conic = its inverse,
    radius = ceil(3 sqrt(lambda_max)), depth = camera z, centre = pinhole projection - 0.5 (pixel centres on integers)."""
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(5)
    n, H, W = 400, 96, 128
    means = (torch.rand((n, 3), generator=g) * 2 - 1) * 0.6        # well inside the frustum: no tan-fov clamping of J
    scales = torch.exp(torch.randn((n, 3), generator=g) * 0.4 + math.log(0.04))
    quats = torch.randn((n, 4), generator=g)
    quats = quats / quats.norm(dim=-1, keepdim=True)
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([0.1, -0.2, 2.5])
    fx = fy = 110.0
    vm = gr.viewmat_from_c2w(c2w)
    pm = gr.projection_matrix(0.001, 1000, 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy)))
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    xys, depths, radii, conics, _, _ = gr.project_gaussians(means, scales, 1, quats, vm[:3], pm @ vm, fx, fy, W / 2, H / 2,
                                                             H, W, tb)
    assert int((radii > 0).sum()) == n
    Rw = Rotation.from_quat(quats[:, [1, 2, 3, 0]].numpy()).as_matrix()      # gsplat quaternions are (w, x, y, z)
    S2 = scales.numpy().astype(np.float64) ** 2
    sig = np.einsum("nij,nj,nkj->nik", Rw, S2, Rw)
    Rv, tv = vm[:3, :3].numpy().astype(np.float64), vm[:3, 3].numpy().astype(np.float64)
    pc = means.numpy().astype(np.float64) @ Rv.T + tv

    def proj(p):
        return torch.stack([fx * p[0] / p[2] + W / 2, fy * p[1] / p[2] + H / 2])

    off_by_one = 0
    for i in range(n):
        p = torch.tensor(pc[i], dtype=torch.float64)
        J = torch.autograd.functional.jacobian(proj, p).numpy()
        cov = J @ (Rv @ sig[i] @ Rv.T) @ J.T + 0.3 * np.eye(2)
        con = np.linalg.inv(cov)
        want = np.array([con[0, 0], con[0, 1], con[1, 1]])
        assert np.abs(conics[i].numpy() - want).max() < 1e-5 * np.abs(want).max(), i
        assert np.abs(xys[i].numpy() - (proj(p).numpy() - 0.5)).max() < 1e-3, i
        assert abs(float(depths[i]) - pc[i, 2]) < 1e-5
        rad = math.ceil(3 * math.sqrt(np.linalg.eigvalsh(cov).max()))
        assert abs(rad - int(radii[i])) <= 1
        off_by_one += rad != int(radii[i])
    assert off_by_one <= n // 50     # fp32 vs fp64 at a ceil() boundary
