"""Drop-in alias: put `<repo>/dropin` (and `<repo>`) on PYTHONPATH and the reference's import paths
(`boxfusion.box_fusion`, `boxfusion.box_manager`, `boxfusion.instances`, `boxfusion.boxes`) resolve to the
B200 implementation in `boxfusion_b200`."""
