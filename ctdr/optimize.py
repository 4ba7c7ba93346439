from shapefromprojections_b200.optimize import run, get_params, projection_step, GraphedIteration  # noqa: F401
