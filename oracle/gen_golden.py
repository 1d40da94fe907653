"""TEST INFRASTRUCTURE -- generates tests/golden/* by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    python -m oracle.gen_golden

For each case it stores, in one .npz: the inputs (rays, near/far, z, w, t_rand), the flat reference
parameter dict (state_dict tensors; data, not code), and the reference's `NeuSRenderer.render`
outputs computed twice through the *same reference code*: in fp32 (`ref32/<key>`) and in fp64
(`ref64/<key>`, modules .double() + default dtype float64) -- SURVEY.md section 8c tolerance plan.

Cases (BASELINE.json configs): cfg1 = 32x32-patch shape family (D=4, n=16, m=0 and the shipped m=4),
cfg2 = headline (D=8, W=128, n=64, m=0; sphere_init SDF weights), cfg4 = 64+64 hierarchical.
Ray counts are sliced down so that the fixtures stay small; instance structure (bs>1) is kept.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import neus_oracle as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

OUT_KEYS = ['s_val', 'cdf_fine', 'weight_sum', 'weight_max', 'gradients', 'weights', 'gradient_error',
            'inside_sphere', 'mid_z_vals', 'surface_loss', 'sdf', 'pts_norm', 'pts', 'color_fine', 'raw_color']

CASES = [
    # name,            D, bs, patch, n,  m, sphere_init, anneal, jitter
    ("cfg1_n16_m0",    4, 2, 12, 16, 0, False, 0.3, False),
    ("cfg1_n16_m4",    4, 2, 12, 16, 4, False, 1.0, False),
    ("cfg1_n16_m4_jit", 4, 1, 12, 16, 4, False, 0.0, True),
    ("cfg2_n64_m0",    8, 2, 7, 64, 0, True, 1.0, False),
    ("cfg2_n64_m0_jit", 8, 1, 8, 64, 0, True, 0.5, True),
    ("cfg4_n64_m64",   8, 1, 6, 64, 64, True, 1.0, False),
    ("cfgd_n16_m4_D8", 8, 3, 6, 16, 4, True, 1.0, False),
    # multi-step up-sampling (renderer.py:400-413; dead in the reference's configs, up_sample_steps: 1 in train.yaml)
    ("cfgs_n16_m8_s2", 4, 2, 8, 16, 8, False, 1.0, False, 2),
    ("cfgs_n16_m12_s4_D8", 8, 2, 6, 16, 12, True, 0.5, True, 4),
]


def run_reference(nets, inp, n, m, anneal, dtype, steps=1):
    sdf, col, dev = [x.to(dtype) for x in nets]
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ro, rd, near, far, z = [inp[k].to(dtype) for k in ("rays_o", "rays_d", "near", "far", "z")]
        with RH.cpu_mode(), torch.no_grad():
            w = sdf.style(z)
        t_rand = inp.get("t_rand")
        if t_rand is not None:
            # inject the jitter deterministically: the reference draws torch.rand([R,1]) at renderer.py:372
            real_rand = torch.rand

            def fake_rand(*a, **k):
                return (t_rand.to(dtype) + 0.5)
            torch.rand = fake_rand
        try:
            out = RH.reference_render(sdf, col, dev, ro, rd, near, far, z, w, n_samples=n, n_importance=m,
                                      cos_anneal_ratio=anneal, perturb_overwrite=(1 if t_rand is not None else 0),
                                      up_sample_steps=steps)
        finally:
            if t_rand is not None:
                torch.rand = real_rand
    finally:
        torch.set_default_dtype(old)
    return w, out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    # ---- known-answer vectors recorded in SURVEY.md section 8c (reference + sphere_init.pt, no RNG)
    kav = {
        "w_first4": [-0.0100471303, 0.0826901346, -0.0250231009, -0.0067402455],
        "w_sum": 0.6362632513,
        "points": [[0, 0, 0], [0.5, 0, 0], [0, 0, 1], [0.3, -0.4, 0.5]],
        "sdf": [-0.4264975786, -0.0221938491, 0.5117361546, 0.2348149121],
        "feat_sum": [-4.5445504189, 16.6883735657, 12.1179466248, 0.2172311544],
        "grad": [[-0.2558322847, -0.0593644455, -0.2747395635], [0.8766956329, -0.1936466545, 0.2837001383],
                 [0.0079841306, 0.0001418108, 0.6624215841], [0.3521813750, -0.1444926858, 0.8254041672]],
        "source": "SURVEY.md section 8(c): reference ShapeNetwork(sphere_init.pt), z=zeros(1,64), CPU fp32 torch 2.11",
    }
    # re-derive them from the reference right now and refuse to write if they disagree
    nets = RH.build_reference_nets(D=8, sphere_init=True, seed=0)
    sdfn = nets[0]
    with RH.cpu_mode():
        w0 = sdfn.style(torch.zeros(1, 64))
        pts = torch.tensor(kav["points"], dtype=torch.float32)
        out = sdfn(pts.clone(), z=None, w=w0)
        grad = sdfn.gradient(pts.clone(), z=None, w=w0)
    assert np.allclose(w0[0, :4].detach().numpy(), kav["w_first4"], atol=1e-6)
    assert abs(float(w0.sum()) - kav["w_sum"]) < 1e-5
    assert np.allclose(out[:, 0].detach().numpy(), kav["sdf"], atol=2e-6)
    assert np.allclose(out[:, 1:].sum(-1).detach().numpy(), kav["feat_sum"], atol=5e-5)
    assert np.allclose(grad.detach().numpy(), kav["grad"], atol=5e-6)
    with open(os.path.join(GOLDEN, "kav_sphere_init.json"), "w") as f:
        json.dump(kav, f, indent=1)

    param_sets = {}
    only = [a for a in sys.argv[1:] if not a.startswith("-")]   # python oracle/gen_golden.py [case names]
    for case in CASES:
        name, D, bs, patch, n, m, sphere, anneal, jitter = case[:9]
        steps = case[9] if len(case) > 9 else 1
        key = (D, sphere)
        if key not in param_sets:
            nets = RH.build_reference_nets(D=D, sphere_init=sphere, seed=7)
            # perturb variance a little so inv_s is not the init constant everywhere
            param_sets[key] = nets
            P = O.extract_params(*nets)
            if not only:
                np.savez_compressed(os.path.join(GOLDEN, f"params_D{D}.npz"), **{k: v.numpy() for k, v in P.items()})
            else:   # adding cases: the rebuilt reference networks must be the committed ones, bit for bit
                with np.load(os.path.join(GOLDEN, f"params_D{D}.npz")) as f:
                    assert all(np.array_equal(f[k], v.numpy()) for k, v in P.items()), f"params_D{D}.npz differs"

        nets = param_sets[key]
        if only and name not in only:
            continue
        seed = 1234 + len(name)
        ro, rd, near, far = O.synthetic_rays(bs, patch, seed=seed)
        g = torch.Generator().manual_seed(seed)
        z = torch.randn(bs, 64, generator=g)
        inp = dict(rays_o=ro, rays_d=rd, near=near, far=far, z=z)
        if jitter:
            inp["t_rand"] = torch.rand(ro.shape[0], 1, generator=g) - 0.5
        w32, out32 = run_reference(nets, inp, n, m, anneal, torch.float32, steps)
        w64, out64 = run_reference(nets, inp, n, m, anneal, torch.float64, steps)
        blob = {f"in/{k}": v.numpy() for k, v in inp.items()}
        blob["in/w"] = w32.detach().numpy()
        blob["meta"] = np.array(json.dumps(dict(name=name, D=D, W=128, bs=bs, patch=patch, n_samples=n,
                                                n_importance=m, up_sample_steps=steps, cos_anneal_ratio=anneal, jitter=jitter,
                                                params=f"params_D{D}.npz", torch=torch.__version__)))
        for k in OUT_KEYS:
            blob[f"ref32/{k}"] = out32[k].detach().numpy()
            blob[f"ref64/{k}"] = out64[k].detach().numpy()
        blob["ref64/w"] = w64.detach().numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **blob)
        d = {k: float((out32[k].double() - out64[k]).abs().max()) for k in OUT_KEYS}
        print(name, "R=%d" % ro.shape[0], "fp32-vs-fp64 Linf:", {k: "%.1e" % v for k, v in d.items()})


if __name__ == "__main__":
    main()
