"""TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): stage the UNMODIFIED reference so that it can run on
the GPU box's host CPU, where /root/reference does not exist.

    python -m oracle.build_ref            # copies /root/reference/opencood -> oracle/_ref/opencood (git-ignored)

The reference is a pure-Python package on this path: nothing is compiled, nothing is edited.  Next to the
package go the five import-only stubs (tests/golden/_stubs: icecream, matplotlib, shapely, pyquaternion, turtle) for
third-party modules the reference imports at module scope but never executes between `create_model` and the head outputs.
`oracle/_ref/` is listed in .gitignore (no reference source enters the history) and NOT in .gpurunignore (it travels to the
GPU box like the built .so files).  Consumers: bench.py's `--impl reference` arm and its `cpu_baseline` leg
(`kind: "reference"`), nothing else; the product package never imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/opencood"
REF_DST = os.path.join(HERE, "_ref")
STUBS = os.path.join(os.path.dirname(HERE), "tests", "golden", "_stubs")
YAML_REL = os.path.join("opencood", "hypes_yaml", "opv2v", "lidar_only_with_noise", "coalign", "pointpillar_coalign.yaml")


def build_ref(force: bool = False) -> str:
    """Returns the staged directory ('' when the reference is not present here and nothing was staged before)."""
    marker = os.path.join(REF_DST, "opencood", "__init__.py")
    if not os.path.isdir(REF_SRC):
        return REF_DST if os.path.exists(marker) else ""
    if os.path.exists(marker) and not force:
        return REF_DST
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    shutil.copytree(REF_SRC, os.path.join(REF_DST, "opencood"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.png", "*.jpg", "*.gif", "logs"))
    shutil.copytree(STUBS, os.path.join(REF_DST, "_stubs"), ignore=shutil.ignore_patterns("__pycache__"))
    return REF_DST


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "opencood", "__init__.py"))


def import_reference():
    """Put the staged reference (and its stubs) on sys.path and return (yaml_utils, train_utils, path of CoAlign's yaml)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python -m oracle.build_ref` where /root/reference exists")
    for p in (os.path.join(REF_DST, "_stubs"), REF_DST):
        if p not in sys.path:
            sys.path.insert(0, p)
    from opencood.hypes_yaml import yaml_utils      # noqa: E402  (the unmodified reference)
    from opencood.tools import train_utils          # noqa: E402
    return yaml_utils, train_utils, os.path.join(REF_DST, YAML_REL)


def build_reference_model(args, sd):
    """The reference's own model object for `args` (our synthetic yaml-equivalent dict) through its own registry
    (train_utils.create_model, /root/reference/opencood/tools/train_utils.py:113-146), weights from `sd`, eval mode."""
    yaml_utils, train_utils, ypath = import_reference()
    hypes = yaml_utils.load_yaml(ypath)
    margs = hypes["model"]["args"]
    margs["lidar_range"] = args["lidar_range"]
    margs["voxel_size"] = args["voxel_size"]
    margs["point_pillar_scatter"]["grid_size"] = args["point_pillar_scatter"]["grid_size"]
    margs["fusion_method"] = args.get("fusion_method", "att")
    model = train_utils.create_model(hypes)
    model.load_state_dict(sd, strict=True)
    return model.eval()


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv) or "reference not present; nothing staged")
