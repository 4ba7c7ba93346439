from shapefromprojections_b200.utils.projection import compute_P, ortho_Mat  # noqa: F401
