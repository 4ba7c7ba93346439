from shapefromprojections_b200.nn.softras_loss import LaplacianLoss, FlattenLoss  # noqa: F401
