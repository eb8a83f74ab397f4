"""The shipped sibling YAMLs (configs/) instantiate the drop-in classes with the reference's own init_args
(configs/cv_builder/{std,concat}_cvbuilder.yaml:1-7, configs/base_model.yaml:28-34)."""
import importlib
from pathlib import Path

import yaml

ROOT = Path(__file__).resolve().parent.parent


def _instantiate(node, **override):
    mod, cls = node["class_path"].rsplit(".", 1)
    args = dict(node["init_args"])
    args.update(override)
    return getattr(importlib.import_module(mod), cls)(**args)


def test_cv_builder_yamls_instantiate():
    from mvs_gi_b200 import modules

    std = _instantiate(yaml.safe_load((ROOT / "configs/cv_builder/std_cvbuilder_b200.yaml").read_text()))
    assert isinstance(std, modules.SphericalSweepStdMasked) and (std.num_cams, std.feat_chs) == (3, 16)
    assert std.post_vol.conv_layer.kernel_size == (3, 3, 3)
    cat = _instantiate(yaml.safe_load((ROOT / "configs/cv_builder/concat_cvbuilder_b200.yaml").read_text()))
    assert isinstance(cat, modules.SphericalSweep) and (cat.num_cams, cat.feat_chs) == (3, 48)


def test_dist_regressor_yaml_instantiates():
    from mvs_gi_b200 import calibration as cal
    from mvs_gi_b200 import modules

    node = yaml.safe_load((ROOT / "configs/dist_regressor_b200.yaml").read_text())["dist_regressor"]
    assert node["init_args"]["dist_cands"] == "sweep@dist_list"  # resolved by the reference's config helper
    reg = _instantiate(node, dist_cands=cal.DIST_16CPD)
    assert isinstance(reg, modules.DistanceRegressorWithFixedCandidates)
    assert reg.interp_scale_factor == 2 and reg.pre_interp and reg.inv_dist_idx.shape == (1, 16, 1, 1)
