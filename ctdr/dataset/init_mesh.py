from shapefromprojections_b200.dataset.init_mesh import *  # noqa: F401,F403
