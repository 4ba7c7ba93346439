from shapefromprojections_b200.utils.util_fun import compute_edge_length  # noqa: F401
