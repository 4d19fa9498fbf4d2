from boxfusion_b200.box_fusion import BoxFusion  # noqa: F401
