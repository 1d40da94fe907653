"""CPU tests: oracle/generator_oracle.py (gen_rays_at, render_maps, Phong light) against golden vectors produced
by the reference's own Generator methods and lighting module (tests/golden/generator_golden.npz)."""
import os

import numpy as np
import torch

from helpers import GOLDEN, linf
from oracle import generator_oracle as GO


def load():
    with np.load(os.path.join(GOLDEN, "generator_golden.npz")) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def test_gen_rays_at_matches_reference():
    G = load()
    cam_dist, res, scene_res = [float(v) for v in G["rays/scalars"]]
    ro, rd, xo, yo = GO.gen_rays_at(G["rays/b2w"], G["rays/c2b"], G["rays/w2c"], cam_dist, int(res), int(scene_res),
                                    G["rays/intrinsics_inv"])
    assert linf(ro, G["rays/rays_o"]) == 0.0
    assert linf(rd, G["rays/rays_d"]) < 1e-6
    assert linf(xo, G["rays/x_offset"]) < 1e-4 and linf(yo, G["rays/y_offset"]) < 1e-4
    assert linf(rd.norm(dim=-1), torch.ones(rd.shape[:-1])) < 1e-6


def test_render_maps_matches_reference():
    G = load()
    amb, dif, spec, shin = [float(v) for v in G["maps/light"]]
    bs = G["maps/bg"].shape[0]
    res = G["maps/out/image"].shape[-1]
    ro_rays = G["rays/rays_o"].reshape(-1, 3)
    render_out = {k[len("maps/in/"):]: v for k, v in G.items() if k.startswith("maps/in/")}
    bg_map = G["maps/bg"][:, :, None, None].expand(bs, 3, res, res)
    out = GO.render_maps(bs, res, render_out, ro_rays, G["maps/light_dir_b"], torch.full((3,), amb),
                         torch.full((3,), dif), torch.full((3,), spec), torch.tensor(shin), bg_map, return_raw=True)
    for k in out:
        ref = G[f"maps/out/{k}"]
        assert out[k].shape == ref.shape, k
        assert linf(out[k], ref) < 2e-6, (k, linf(out[k], ref))
    # light direction in the box frame = R(w2b) @ direction
    d = GO.light_batch_direction(G["maps/w2b"], torch.tensor([0.0, 0.0, -1.0]))
    assert linf(d, G["maps/light_dir_b"]) < 1e-6
