"""Namespace with the reference's public names, as `demo.py` imports them (demo.py:17-24):

    from boxfusion.instances import Instances3D
    from boxfusion.box_manager import BoxManager
    from boxfusion.box_fusion import BoxFusion

`dropin/boxfusion/` re-exports these under the reference's module paths; tests and the bench use this
namespace as the `impl` of boxfusion_b200.driver.FusionSession.
"""
from .boxes import GeneralInstance3DBoxes
from .box_fusion import BoxFusion
from .box_manager import BoxManager
from .instances import Instances3D, calculate_obb_iou, nms_3d

__all__ = ["GeneralInstance3DBoxes", "BoxFusion", "BoxManager", "Instances3D", "calculate_obb_iou", "nms_3d"]
