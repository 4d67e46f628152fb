"""coalign_b200 - B200-native (sm_100a) implementation of CoAlign's per-frame hot path.

    import coalign_b200; coalign_b200.register()      # then: model.core_method: point_pillar_coalign_b200

See DESIGN.md / INTEGRATION.md.  The CUDA library is mandatory; nothing here falls back to CPU.
"""
import os

__all__ = ["register", "PointPillarCoalignB200", "PointPillarB200", "PointPillarUncertaintyB200"]
PLUGIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencood_plugin")
LOSS_PLUGIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencood_plugin_loss")


def register():
    """Make `opencood.models.point_pillar_coalign_b200` importable for train_utils.create_model without
    touching the reference tree (extends the package search path of opencood.models)."""
    import opencood.models as m
    if PLUGIN_DIR not in list(m.__path__):
        m.__path__.append(PLUGIN_DIR)
    import opencood.loss as lo                        # train_utils.create_loss: loss.core_method: point_pillar_loss_b200
    if LOSS_PLUGIN_DIR not in list(lo.__path__):
        lo.__path__.append(LOSS_PLUGIN_DIR)


def __getattr__(name):
    if name in ("PointPillarCoalignB200", "PointPillarB200", "PointPillarUncertaintyB200"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
