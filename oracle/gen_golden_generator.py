"""TEST INFRASTRUCTURE -- generates tests/golden/generator_golden.npz from the reference's own
`Generator.gen_rays_at`, `Generator.render_maps` (src/models/generator.py:80-174,255-279) and Phong light
(src/models/lighting.py), run on CPU in the authoring container:

    python -m oracle.gen_golden_generator

The two methods are called unbound on a small stand-in for `self` that carries exactly the attributes they read
(`camera`, `resolution`, `scene_resolution`, `bg_color`); camera and light are the reference's own classes.
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import neus_oracle as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def random_b2w(bs, g):
    """Box poses like the Plane prior produces (data/example/cfg.yaml): rotation about an axis + in-plane shift."""
    out = []
    for _ in range(bs):
        a = torch.randn(3, generator=g)
        a = a / a.norm()
        th = float(torch.rand(1, generator=g)) * 2 * np.pi
        K = torch.tensor([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        Rm = torch.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
        t = torch.tensor([float(torch.rand(1, generator=g) - 0.5) * 4.0, float(torch.rand(1, generator=g) - 0.5) * 2.5,
                          float(torch.rand(1, generator=g) - 0.5) * 0.5])
        M = torch.eye(4)
        M[:3, :3] = Rm
        M[:3, 3] = t
        out.append(M)
    return torch.stack(out)


def main():
    RH.import_reference()
    from src.models.camera_network import Camera
    from src.models.generator import Generator
    from src.models.lighting import DirectionalLightWithSpecularFixInit
    from src.utils.pose import invert_rot_t

    g = torch.Generator().manual_seed(5)
    blob = {}
    res, fov, img, img_scene = 12, 10.0, 256, 1588
    cam_dist = 1.0 / np.tan(np.radians(fov / 2))
    scene_res = int(res * img_scene / img)
    scene_fov = 2 * np.degrees(np.arctan(np.tan(np.radians(fov / 2)) * img_scene / img))
    bs = 3
    with RH.cpu_mode():
        camera = Camera(cam_dist=cam_dist, fov=scene_fov, resolution=scene_res)
        light = DirectionalLightWithSpecularFixInit(direction=np.array([0.0, 0.0, -1.0]), ambient_color=0.33,
                                                    diffuse_color=0.66, specular_color=0.2, shininess=10)
        b2w = random_b2w(bs, g)
        w2b = invert_rot_t(b2w)
        c2b = torch.einsum('bij,jk->bik', w2b, camera.c2w)
        fake = types.SimpleNamespace(camera=camera, resolution=res, scene_resolution=scene_res)
        with torch.no_grad():
            rays = Generator.gen_rays_at(fake, {}, {'b2w': b2w, 'c2b': c2b})
    for k in ('b2w', 'c2b'):
        blob[f"rays/{k}"] = locals()[k].numpy()
    blob["rays/w2c"], blob["rays/c2w"] = camera.w2c.numpy(), camera.c2w.numpy()
    blob["rays/intrinsics_inv"] = camera.intrinsics_inv.numpy()
    blob["rays/scalars"] = np.array([cam_dist, res, scene_res], dtype=np.float64)
    for k in ('rays_o', 'rays_d', 'x_offset', 'y_offset'):
        blob[f"rays/{k}"] = rays[k].contiguous().numpy()

    # ---- render_maps on a real render of those rays (sphere_init SDF net, D=8, n=16 + 4)
    nets = RH.build_reference_nets(D=8, sphere_init=True, seed=7)
    z = torch.randn(bs, 64, generator=g)
    ro = rays['rays_o'].reshape(-1, 3).contiguous()
    rd = rays['rays_d'].reshape(-1, 3).contiguous()
    near, far = O.near_far_from_sphere(ro, rd)
    with RH.cpu_mode(), torch.no_grad():
        w = nets[0].style(z)
    out = RH.reference_render(nets[0], nets[1], nets[2], ro, rd, near, far, z, w, n_samples=16, n_importance=4,
                              cos_anneal_ratio=1.0, perturb_overwrite=0)
    bg = torch.rand(bs, 3, generator=g)
    fake.bg_color = lambda n: bg[:, :, None, None].expand(n, 3, res, res)
    with RH.cpu_mode(), torch.no_grad():
        lb = light.batch_transform(w2b=w2b)
        maps = Generator.render_maps(fake, bs, {k: v.clone() for k, v in out.items()},
                                     {'rays_o': rays['rays_o']}, {'light': lb}, True)
        blob["maps/light_dir_b"] = lb.batch_direction(out['pts'].reshape(bs, -1, 3))[:, 0, :].numpy()
        blob["maps/light"] = np.array([float(light.ambient_color[0]), float(light.diffuse_color[0]),
                                       float(light.specular_color[0]), float(light.shininess)], dtype=np.float64)
    blob["maps/w2b"], blob["maps/bg"], blob["maps/z"] = w2b.numpy(), bg.numpy(), z.numpy()
    for k in ('pts', 'weights', 'weight_sum', 'color_fine', 'gradients', 'raw_color', 'mid_z_vals'):
        blob[f"maps/in/{k}"] = out[k].numpy()
    for k, v in maps.items():
        blob[f"maps/out/{k}"] = v.contiguous().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "generator_golden.npz"), **blob)
    print("wrote", len(blob), "arrays;", {k: tuple(v.shape) for k, v in maps.items()})


if __name__ == "__main__":
    main()
