from shapefromprojections_b200.nn.fp_opengl import FP, get_valid_mask  # noqa: F401
