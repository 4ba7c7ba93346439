from shapefromprojections_b200.utils.util_mesh import *  # noqa: F401,F403
