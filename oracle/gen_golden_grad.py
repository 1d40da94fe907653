"""TEST INFRASTRUCTURE -- parameter gradients of the UNMODIFIED reference render, as golden fixtures.

Run in the authoring container only (needs /root/reference):

    python -m oracle.gen_golden_grad

The reference obtains d(loss)/d(theta) by autograd through `NeuSRenderer.render` with the SDF normal built by
`autograd.grad(create_graph=True)` (src/models/fields.py:104-122) and differentiated a second time at
src/trainers/gan_pose_trainer.py:141.  For two cases this script runs exactly that -- the reference's own modules,
`render(...)` in grad mode with both `z` and `w = sdf_network.style(z)` -- in fp32 and in fp64, with

    loss = color_fine.sum() + weight_sum.sum() + 10 * gradient_error

and stores every parameter's gradient (all 56 `sdf_network` tensors incl. the 3 style layers reached through `w`, the
8 `color_network` tensors and `deviation_network.variance`), the gradient w.r.t. `w` itself, and the fine z-values
the reference rendered (captured from the arguments of `render_core`; the hierarchical resampling runs under
no_grad, renderer.py:389-415, so they are constants of the backward).  Cases reuse the inputs of the forward
fixtures `cfg2_n64_m0` (fixed z, D=8, 2 instances) and `cfgd_n16_m4_D8` (16+4 hierarchical, D=8, 3 instances).
tests/test_backward_reference.py compares oracle/backward_oracle.py (CPU) and the CUDA backward (-m gpu) with them.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as RH  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = ["cfg2_n64_m0", "cfgd_n16_m4_D8"]
LOSS = "color_fine.sum() + weight_sum.sum() + 10 * gradient_error"


def reference_gradients(meta, inp, dtype):
    nets = RH.build_reference_nets(D=meta["D"], sphere_init=True, seed=7)   # same construction as gen_golden.py
    sdf, col, dev = [x.to(dtype) for x in nets]
    with np.load(os.path.join(GOLDEN, meta["params"])) as f:               # ... and the same values, checked
        for prefix, mod in (("sdf_network.", sdf), ("color_network.", col), ("deviation_network.", dev)):
            for k, v in mod.state_dict().items():
                assert np.array_equal(f[prefix + k], v.float().numpy()), prefix + k
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ro, rd, near, far, z = [inp[k].to(dtype) for k in ("rays_o", "rays_d", "near", "far", "z")]
        ref = RH.import_reference()
        r = ref["NeuSRenderer"](nerf=None, sdf_network=sdf, deviation_network=dev, color_network=col,
                                n_samples=meta["n_samples"], n_importance=meta["n_importance"], n_outside=0,
                                up_sample_steps=1, perturb=1)
        seen = {}
        inner = r.render_core

        def spy(rays_o, rays_d, z_vals, *a, **k):      # records the z-values the reference actually renders
            seen["z_vals"] = z_vals.detach().clone()
            return inner(rays_o, rays_d, z_vals, *a, **k)
        r.render_core = spy
        with RH.cpu_mode():
            w = sdf.style(z)
            w.retain_grad()
            out = r.render(ro, rd, near, far, background_rgb=None, cos_anneal_ratio=meta["cos_anneal_ratio"],
                           perturb_overwrite=0, z=z, w=w)
            loss = out["color_fine"].sum() + out["weight_sum"].sum() + 10.0 * out["gradient_error"]
            loss.backward()
        grads = {}
        for prefix, mod in (("sdf_network.", sdf), ("color_network.", col), ("deviation_network.", dev)):
            for k, p in mod.named_parameters():
                assert p.grad is not None and float(p.grad.abs().max()) > 0, prefix + k
                grads[prefix + k] = p.grad.detach().clone()
        grads["w"] = w.grad.detach().clone()
    finally:
        torch.set_default_dtype(old)
    return grads, seen["z_vals"], float(loss)


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN)))
    from helpers import load_case
    for name in CASES:
        meta, inp, _, _ = load_case(name)
        g32, z32, l32 = reference_gradients(meta, inp, torch.float32)
        g64, z64, l64 = reference_gradients(meta, inp, torch.float64)
        blob = {"meta": np.array(json.dumps(dict(case=name, loss=LOSS, loss32=l32, loss64=l64, torch=torch.__version__)))}
        blob["z_vals32"], blob["z_vals64"] = z32.numpy(), z64.numpy()
        worst = (0.0, None)
        for k in g64:
            blob[f"g32/{k}"], blob[f"g64/{k}"] = g32[k].numpy(), g64[k].numpy()
            rel = float((g32[k].double() - g64[k]).abs().max()) / (float(g64[k].abs().max()) + 1e-30)
            if rel > worst[0]:
                worst = (rel, k)
        np.savez_compressed(os.path.join(GOLDEN, f"grad_{name}.npz"), **blob)
        print(f"{name}: {len(g64)} gradient tensors, loss {l64:.6f}, worst fp32-vs-fp64 relative Linf "
              f"{worst[0]:.2e} ({worst[1]}), |z32 - z64| {float((z32.double() - z64).abs().max()):.2e}")


if __name__ == "__main__":
    main()
