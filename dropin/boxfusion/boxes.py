from boxfusion_b200.boxes import GeneralInstance3DBoxes  # noqa: F401
