from boxfusion_b200.instances import Instances3D, calculate_obb_iou, nms_3d  # noqa: F401
