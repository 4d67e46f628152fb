"""`opencood.models.point_pillar_uncertainty_b200` - the stage-1 single-agent detector with the uncertainty head
(reference core_method `point_pillar_uncertainty`) on the B200 path; found by the reference registry
(/root/reference/opencood/tools/train_utils.py:127-146) once `coalign_b200.register()` has put this directory on
`opencood.models.__path__`.  yaml: `model.core_method: point_pillar_uncertainty_b200`."""
from coalign_b200.model import PointPillarUncertaintyB200  # noqa: F401  (class name == core_method sans '_')
