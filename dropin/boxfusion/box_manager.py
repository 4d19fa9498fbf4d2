from boxfusion_b200.box_manager import BoxManager  # noqa: F401
