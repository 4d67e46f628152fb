"""`opencood.loss.point_pillar_loss_b200` - the CoAlign loss (reference core_method `point_pillar_loss`) with the loss terms
and their gradients computed by `cb_pointpillar_loss`; found by the reference's loss registry
(/root/reference/opencood/tools/train_utils.py:149-181) once `coalign_b200.register()` has put this directory on
`opencood.loss.__path__`.  yaml: `loss.core_method: point_pillar_loss_b200`."""
from coalign_b200.loss import PointPillarLossB200  # noqa: F401  (class name == core_method sans '_')
