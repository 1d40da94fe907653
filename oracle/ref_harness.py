"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference.

This module exists so that the oracle restatement (oracle/neus_oracle.py) can be pinned
against the reference's own code in the authoring container, and so that the golden
fixtures under tests/golden/ can be (re)generated (oracle/gen_golden.py).  Nothing on the
product path may import it, and nothing that runs on the GPU box may need it:
/root/reference does not exist there.

Shims applied (no reference source is modified or copied):
  * collections.MutableMapping & friends alias   (tu/configs.py:108 uses the py<3.10 name)
  * a TorchFunctionMode that rewrites device='cuda' kwargs to 'cpu' and makes Tensor.cuda()
    the identity (renderer.py:53,56,106,166,178,222,300,359,372 hard-code device='cuda')
  * torch.load forced to map_location='cpu'     (src/models/fields.py:33)
"""
import collections
import collections.abc
import contextlib
import os
import sys

import torch
from torch.overrides import TorchFunctionMode

REFERENCE_DIR = os.environ.get("OI_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_DIR, "src", "third_party", "neus"))


class _CudaToCpu(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        dev = kwargs.get("device", None)
        if dev is not None and "cuda" in str(dev):
            kwargs["device"] = "cpu"
        if func is torch.Tensor.cuda:
            return args[0]
        return func(*args, **kwargs)


_imported = {}


def import_reference():
    """Returns a namespace dict with the reference classes on the hot path."""
    if _imported:
        return _imported
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_DIR}")
    for n in ("MutableMapping", "Mapping", "Sequence", "Iterable"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/oi_ref_torch_ext")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    from src.third_party.neus.models.renderer import NeuSRenderer  # noqa
    from src.third_party.neus.models.fields import SingleVarianceNetwork  # noqa
    from src.models.fields import ShapeNetwork, ColorNetwork  # noqa
    _imported.update(NeuSRenderer=NeuSRenderer, SingleVarianceNetwork=SingleVarianceNetwork,
                     ShapeNetwork=ShapeNetwork, ColorNetwork=ColorNetwork)
    return _imported


@contextlib.contextmanager
def cpu_mode():
    """Context in which the reference's hard-coded cuda device strings land on the CPU."""
    real_load = torch.load

    def load_cpu(*a, **k):
        k["map_location"] = "cpu"
        k.setdefault("weights_only", False)
        return real_load(*a, **k)

    torch.load = load_cpu
    cwd = os.getcwd()
    os.chdir(REFERENCE_DIR)
    try:
        with _CudaToCpu():
            yield
    finally:
        torch.load = real_load
        os.chdir(cwd)


def build_reference_nets(D=8, W=128, style_dim=64, sphere_init=True, seed=0, dtype=torch.float32):
    """ShapeNetwork / ColorNetwork / SingleVarianceNetwork exactly as configs/train.yaml:36-57 builds them."""
    ref = import_reference()
    torch.manual_seed(seed)
    with cpu_mode():
        kw = dict(D=D, W=W, input_ch=3, input_ch_views=3, style_dim=style_dim)
        ckpt = "./checkpoints/sphere_init.pt" if (sphere_init and D == 8 and W == 128) else None
        sdf = ref["ShapeNetwork"](checkpoint_path=ckpt, **kw)
        col = ref["ColorNetwork"](**kw)
        dev = ref["SingleVarianceNetwork"](init_val=0.3)
    return sdf.to(dtype), col.to(dtype), dev.to(dtype)


def reference_render(sdf, col, dev, rays_o, rays_d, near, far, z, w, *, n_samples, n_importance,
                     up_sample_steps=1, cos_anneal_ratio=1.0, perturb_overwrite=0, grad=False):
    """Runs the reference's NeuSRenderer.render (renderer.py:351-473) on CPU."""
    ref = import_reference()
    r = ref["NeuSRenderer"](nerf=None, sdf_network=sdf, deviation_network=dev, color_network=col,
                            n_samples=n_samples, n_importance=n_importance, n_outside=0,
                            up_sample_steps=up_sample_steps, perturb=1)
    ctx = contextlib.nullcontext() if grad else torch.no_grad()
    with cpu_mode(), ctx:
        out = r.render(rays_o, rays_d, near, far, background_rgb=None, cos_anneal_ratio=cos_anneal_ratio,
                       perturb_overwrite=perturb_overwrite, z=z, w=w)
    return out
