from shapefromprojections_b200.utils.util_sino import *  # noqa: F401,F403
