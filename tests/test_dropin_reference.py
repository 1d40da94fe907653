"""Drop-in boundary against the real reference tree (runs only where /root/reference exists, i.e. the authoring
container): the reference's own `build_from_config` instantiates `object_intrinsics_b200.renderer.NeuSRenderer`
from the `renderer` block of configs/train.yaml with only `__target__` changed -- exactly what the CLI override
`model.generator.kwargs.renderer.__target__=...` does (generator.py:51-57, tu/utils/config.py:19-25) -- and the
resulting object exposes what Generator and scripts/test.py read from it."""
import inspect
import os

import pytest
import torch
import yaml

from oracle import ref_harness as RH

pytestmark = pytest.mark.skipif(not RH.reference_available(), reason="reference tree not present")


def test_reference_build_from_config_instantiates_the_drop_in():
    RH.import_reference()
    from tu.utils.config import build_from_config
    with open(os.path.join(RH.REFERENCE_DIR, "configs", "train.yaml")) as f:
        cfg = yaml.safe_load(f)
    block = cfg["model"]["generator"]["kwargs"]["renderer"]
    assert block["__target__"] == "src.third_party.neus.models.renderer.NeuSRenderer"
    block = dict(block, __target__="object_intrinsics_b200.renderer.NeuSRenderer")
    sdf, col, dev = RH.build_reference_nets(D=8, sphere_init=True)
    r = build_from_config(block, nerf=None, sdf_network=sdf, deviation_network=dev, color_network=col)
    from object_intrinsics_b200.renderer import NeuSRenderer, collect_params
    assert isinstance(r, NeuSRenderer)
    assert (r.n_samples, r.n_importance, r.n_outside, r.up_sample_steps, r.perturb) == (16, 4, 0, 1, 1)
    assert r.sdf_network is sdf and r.color_network is col and r.deviation_network is dev
    # the renderer owns no parameters and finds every tensor of the path on the reference's own modules
    names = [n for n, _ in collect_params(sdf, col, dev)]
    gen_keys = {f"sdf_network.{k}" for k in sdf.state_dict()} | {f"color_network.{k}" for k in col.state_dict()} | \
        {f"deviation_network.{k}" for k in dev.state_dict()}
    assert set(names) == gen_keys
    # same render() signature as the reference class (extra keyword-only arguments allowed)
    ref_sig = inspect.signature(RH.import_reference()["NeuSRenderer"].render)
    my_sig = inspect.signature(NeuSRenderer.render)
    ref_params = [(p.name, p.default) for p in ref_sig.parameters.values()]
    mine = [(p.name, p.default) for p in my_sig.parameters.values() if p.kind != p.KEYWORD_ONLY]
    assert mine == ref_params
    # CPU tensors fail loudly (no silent fallback to the reference path)
    ro = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.render(ro, ro, ro[:, :1], ro[:, :1], z=torch.zeros(1, 64), w=torch.zeros(1, 64))
