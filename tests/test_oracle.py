"""CPU tests: the oracle restatement (oracle/neus_oracle.py) against the reference's pins.

The reference holds no numerical tests for this path, so the pins are
(1) tests/golden/kav_sphere_init.json -- known-answer vectors of the reference ShapeNetwork on
    checkpoints/sphere_init.pt recorded in SURVEY.md section 8(c) and re-derived by oracle/gen_golden.py;
(2) tests/golden/<case>.npz -- outputs of the unmodified reference NeuSRenderer.render (fp32 and fp64).
"""
import json
import os

import pytest
import torch

from helpers import CASES, GOLDEN, OUT_KEYS, linf, load_case, load_params
from oracle import neus_oracle as O


def _run(name, dtype, **kw):
    meta, inp, r32, r64 = load_case(name)
    P = load_params(meta["params"], dtype)
    a = {k: v.to(dtype) for k, v in inp.items()}
    w = O.style_mlp(P, a["z"])
    out = O.render(P, a["rays_o"], a["rays_d"], a["near"], a["far"], z=a["z"], w=w,
                   n_samples=meta["n_samples"], n_importance=meta["n_importance"],
                   cos_anneal_ratio=meta["cos_anneal_ratio"], t_rand=a.get("t_rand"),
                   up_sample_steps=meta.get("up_sample_steps", 1), **kw)
    return meta, inp, r32, r64, w, out


def test_known_answer_vectors_sphere_init():
    with open(os.path.join(GOLDEN, "kav_sphere_init.json")) as f:
        kav = json.load(f)
    P = load_params("params_D8.npz")
    w = O.style_mlp(P, torch.zeros(1, 64))
    assert torch.allclose(w[0, :4], torch.tensor(kav["w_first4"]), atol=1e-6)
    assert abs(float(w.sum()) - kav["w_sum"]) < 1e-5
    pts = torch.tensor(kav["points"], dtype=torch.float32)
    out = O.sdf_network_forward(P, pts, w)
    assert torch.allclose(out[:, 0], torch.tensor(kav["sdf"]), atol=2e-6)
    assert torch.allclose(out[:, 1:].sum(-1), torch.tensor(kav["feat_sum"]), atol=5e-5)
    g = O.sdf_gradient(P, pts, w)
    assert torch.allclose(g, torch.tensor(kav["grad"]), atol=5e-6)
    _, _, ga = O.sdf_gradient_analytic(P, pts, w)
    assert torch.allclose(ga, torch.tensor(kav["grad"]), atol=5e-6)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp32_matches_reference_fp32(name):
    # same torch build, same op sequence -> expected bit-identical; allow a few ulp for other BLAS builds
    meta, inp, r32, r64, w, out = _run(name, torch.float32)
    assert linf(w, inp["w"]) <= 1e-6
    tol = 1e-5 if meta["n_importance"] == 0 else 2e-3   # inverse-CDF step is ill-conditioned
    for k in OUT_KEYS:
        assert out[k].shape == r32[k].shape, k
        assert linf(out[k], r32[k]) <= tol, (k, linf(out[k], r32[k]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_matches_reference_fp64(name):
    meta, inp, r32, r64, w, out = _run(name, torch.float64)
    for k in OUT_KEYS:
        assert linf(out[k], r64[k]) <= 1e-7, (k, linf(out[k], r64[k]))


@pytest.mark.parametrize("name", ["cfg1_n16_m0", "cfg2_n64_m0"])
def test_analytic_gradient_equals_autograd(name):
    _, _, _, _, _, out_a = _run(name, torch.float64, analytic_gradient=True)
    _, _, _, _, _, out_b = _run(name, torch.float64)
    for k in OUT_KEYS:
        assert linf(out_a[k], out_b[k]) <= 1e-12, k


def test_init_params_shapes_match_reference_params():
    ref = load_params("params_D8.npz")
    mine = O.init_params(D=8, W=128, style_dim=64)
    assert set(ref) == set(mine)
    for k in ref:
        assert ref[k].shape == mine[k].shape, k


def test_degenerate_inputs():
    P = O.init_params(D=4, W=128, style_dim=64, seed=3)
    ro, rd, near, far = O.synthetic_rays(1, 2, seed=1)
    z = torch.zeros(1, 64)
    out = O.render(P, ro, rd, near, far, z=z, n_samples=2, n_importance=0, cos_anneal_ratio=0.0)
    assert out["weights"].shape == (4, 2) and torch.isfinite(out["color_fine"]).all()
    # rays that miss the unit sphere entirely: weights ~ 0 but finite
    ro2 = ro + torch.tensor([0.0, 0.0, 50.0])
    n2, f2 = O.near_far_from_sphere(ro2, rd)
    out2 = O.render(P, ro2, rd, n2, f2, z=z, n_samples=8, n_importance=4, cos_anneal_ratio=1.0)
    assert torch.isfinite(out2["weights"]).all() and out2["inside_sphere"].sum() == 0


def test_bench_input_generator_equals_the_oracles():
    """bench.py's repo arm builds its rays with bench_inputs.py (no oracle import on that arm); same rays, bit for bit."""
    import bench_inputs as BI
    for bs, patch, seed in ((1, 64, 1234), (4, 16, 1237), (3, 5, 5)):
        for a, b in zip(BI.synthetic_rays(bs, patch, seed=seed), O.synthetic_rays(bs, patch, seed=seed)):
            assert torch.equal(a, b)
    ref = load_params("params_D8.npz")
    mine = BI.load_flat_params("params_D8.npz")
    assert set(ref) == set(mine) and all(torch.equal(ref[k], mine[k]) for k in ref)
