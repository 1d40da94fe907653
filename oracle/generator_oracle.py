"""CPU ORACLE for the two callers either side of the render path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

SURVEY.md section 8(f): rank 3 = the step before the path (`Generator.gen_rays_at` + `build_rays` +
`near_far_from_sphere`, src/models/generator.py:255-279,317-342) and rank 1 = the step after it
(`Generator.render_maps` with the Phong light of src/models/lighting.py:61-76,94-119,126-225,
generator.py:80-174).  Plain-torch restatements citing the reference lines; pinned by
tests/golden/generator_golden.npz, generated from the reference's own functions by
oracle/gen_golden_generator.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# ray generation
# ------------------------------------------------------------------------------------------------
def build_rays(h_recp_size, w_recp_size, h_offset, w_offset, num_rays_h, num_rays_w, intrinsics_inv):
    """generator.py:317-333: pixel grid (linspace(0,1) * size + offset) through K^-1, normalised; [bs,h,w,3]."""
    tx = torch.linspace(0, 1, num_rays_w, dtype=intrinsics_inv.dtype)
    ty = torch.linspace(0, 1, num_rays_h, dtype=intrinsics_inv.dtype)
    pixels_x, pixels_y = torch.meshgrid(tx, ty, indexing='ij')            # (w, h)
    pixels_x = pixels_x * w_recp_size + w_offset[..., None, None]
    pixels_y = pixels_y * h_recp_size + h_offset[..., None, None]
    p = torch.stack([pixels_x, pixels_y, torch.ones_like(pixels_y)], dim=-1)  # (..., w, h, 3)
    p = torch.einsum('ij,...whj->...whi', intrinsics_inv[:3, :3], p)
    p = torch.einsum('...whi->...hwi', p)
    return p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)


def gen_rays_at(b2w, c2b, w2c, cam_dist, resolution, scene_resolution, intrinsics_inv):
    """generator.py:255-279: crop centre from the projected box origin, rays through the crop, rotated into the
    box frame; rays_o = camera centre in the box frame.  Returns rays_o, rays_d [bs,h,w,3], x_offset, y_offset."""
    b2c = torch.einsum('ij,bjk->bik', w2c, b2w)
    t = b2c[..., :3, 3]
    center_x = cam_dist / t[..., 2] * t[..., 0] * resolution / 2 + 1 / 2 * scene_resolution
    center_y = cam_dist / t[..., 2] * t[..., 1] * resolution / 2 + 1 / 2 * scene_resolution
    x_offset = center_x - resolution / 2
    y_offset = center_y - resolution / 2
    rays_v = build_rays(resolution, resolution, y_offset, x_offset, resolution, resolution, intrinsics_inv)
    rays_v = torch.einsum('bij,bhwj->bhwi', c2b[..., :3, :3], rays_v)
    rays_o = c2b[:, None, None, :3, 3].expand(rays_v.shape)
    return rays_o, rays_v, x_offset, y_offset


# ------------------------------------------------------------------------------------------------
# Phong shading (lighting.py:126-225)
# ------------------------------------------------------------------------------------------------
def light_batch_direction(w2b, direction):
    """lighting.py:115-119 (direction of the directional light in each box frame), [bs,3]."""
    return torch.einsum('bij,j->bi', w2b[:, :3, :3], direction)


def diffuse(normals, color, direction):
    normals = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    direction = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    angle = F.relu(torch.sum(normals * direction, dim=-1))
    return color * angle[..., None]


def specular(points, normals, direction, color, camera_position, shininess):
    normals = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    direction = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    cos_angle = torch.sum(normals * direction, dim=-1)
    mask = (cos_angle > 0).to(points.dtype)
    view_direction = F.normalize(camera_position - points, p=2, dim=-1, eps=1e-6)
    reflect_direction = -direction + 2 * (cos_angle[..., None] * normals)
    alpha = F.relu(torch.sum(view_direction * reflect_direction, dim=-1)) * mask
    return color * torch.pow(alpha, shininess)[..., None]


def render_maps(bs, resolution, render_out, rays_o, light_dir_b, ambient_color, diffuse_color, specular_color,
                shininess, bg_map, return_raw=True):
    """generator.py:80-174.  render_out: the renderer's dict ([R,S] / [R,S,3] tensors, R = bs*h*w); rays_o [R,3];
    light_dir_b [bs,3] (box-frame light direction); *_color: [3] tensors; bg_map [bs,3,h,w]."""
    h = w = resolution
    n_pts = render_out['pts'].shape[1]
    weights_pts = render_out['weights'].unsqueeze(-1)

    def to_map(x):                       # (bs*h*w, c) -> (bs, c, h, w)
        return x.reshape(bs, h, w, x.shape[-1]).permute(0, 3, 1, 2)

    def pts_to_map(x):                   # (bs*h*w, n_pts, c) -> (bs, c, h, w)
        return to_map((x * weights_pts).sum(-2))

    normal_pts, color_pts, pts = render_out['gradients'], render_out['raw_color'], render_out['pts']
    direction = light_dir_b[:, None, :].expand(bs, h * w * n_pts, 3).reshape(bs * h * w, n_pts, 3)
    ret = {'weight_sum_map': to_map(render_out['weight_sum']), 'color_map': to_map(render_out['color_fine'])}
    ambient = ambient_color[None, None, :].expand(bs * h * w, n_pts, 3)
    diff = diffuse(normal_pts, diffuse_color, direction)
    shading = ambient + diff
    ret['shading_map'] = pts_to_map(shading)
    no_spec = pts_to_map(shading * color_pts)
    spec = pts_to_map(specular(pts, normal_pts, direction, specular_color,
                               rays_o[:, None, :].expand_as(pts), shininess))
    rgb = no_spec + spec
    ret['image_no_bg'] = rgb
    ret['image'] = rgb + bg_map * (1 - ret['weight_sum_map'])
    ret['mask'] = ret['weight_sum_map'].clamp(1e-3, 1.0 - 1e-3)
    if return_raw:
        ret['amb_shading_map'] = pts_to_map(ambient)
        ret['diff_shading_map'] = pts_to_map(diff)
        ret['normal_map'] = pts_to_map(normal_pts)
        ret['no_specular_map'] = no_spec
        ret['specular_map'] = spec
        z_rays = torch.einsum('bn,bn->b', render_out['mid_z_vals'], render_out['weights']).unsqueeze(-1)
        ret['z_map'] = to_map(z_rays)
        ret['z_min'] = render_out['mid_z_vals'].min(-1).values.reshape(bs, -1).min(-1).values
    return ret
