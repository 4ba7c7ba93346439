from shapefromprojections_b200.utils.util_vis import save_sino_as_img  # noqa: F401
