from shapefromprojections_b200.model.vanilla import Model  # noqa: F401
