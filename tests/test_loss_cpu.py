"""SURVEY 8f row 2 (first piece): PointPillarLoss - oracle restatement against the unmodified reference (CPU)."""
import os

import numpy as np
import pytest

from coalign_b200 import synth
from oracle import loss_oracle as LO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "typical": dict(seed=3, n=2, H=12, W=20),
    "empty_sample": dict(seed=4, n=3, H=10, W=16, empty_samples=(1,)),
    "float32_labels": dict(seed=5, n=1, H=8, W=12, dtype=np.float32),
}


@pytest.mark.parametrize("name", list(CASES))
def test_loss_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"loss_{name}.npz"))
    losses, grads = LO.loss_and_grads(synth.loss_args(), synth.loss_case(**CASES[name]))
    for k in ("total_loss", "reg_loss", "cls_loss", "dir_loss"):
        assert abs(losses[k] - float(g[k])) <= 1e-6 * abs(float(g[k])) + 1e-7, (k, losses[k], float(g[k]))
    for k in ("cls", "reg", "dir"):
        ref = g[f"g_{k}"]
        assert grads[k].shape == ref.shape
        np.testing.assert_allclose(grads[k], ref, rtol=1e-5, atol=1e-8, err_msg=k)
    assert np.abs(g["g_reg"]).sum() > 0 and np.abs(g["g_dir"]).sum() > 0


def test_loss_abi_symbols_and_argument_errors():
    """The library exports the loss entry points and rejects bad arguments before touching the device."""
    from coalign_b200 import _lib
    lib = _lib.load(check_device=False)
    assert lib.cb_pointpillar_loss_workspace_bytes(2, 12, 20, 2) >= 2 * 8 + 4 * 3 * 8
    assert lib.cb_pointpillar_loss_workspace_bytes(0, 12, 20, 2) == 0
    rc = lib.cb_pointpillar_loss(None, None, None, None, None, None, 1, 2, 12, 20, 2, 2, 2.0, 0.25, 2.0, 2.0, 3.0, 2.0, 0.2,
                                 0.7853, None, None, None, None, None, None, 0, None)
    assert rc == -1


def test_loss_plugin_resolves_like_the_reference_registry():
    """`loss.core_method: point_pillar_loss_b200` -> PointPillarLossB200 by the name match of train_utils.create_loss
    (train_utils.py:163-171); same constructor argument as the reference class."""
    import importlib
    import sys
    from coalign_b200 import LOSS_PLUGIN_DIR
    from coalign_b200.loss import PointPillarLossB200
    sys.path.insert(0, LOSS_PLUGIN_DIR)
    try:
        mod = importlib.import_module("point_pillar_loss_b200")
    finally:
        sys.path.remove(LOSS_PLUGIN_DIR)
    target = "point_pillar_loss_b200".replace("_", "").lower()
    assert [c for name, c in mod.__dict__.items() if name.lower() == target] == [PointPillarLossB200]
    crit = PointPillarLossB200(synth.loss_args())
    assert crit.loss_dict == {} and crit.dir["args"]["num_bins"] == 2
    with pytest.raises(NotImplementedError):
        PointPillarLossB200(dict(synth.loss_args(), iou={"sigma": 3.0, "weight": 1.0}))
