from shapefromprojections_b200.nn.fp_vecs import FP, get_valid_mask  # noqa: F401
