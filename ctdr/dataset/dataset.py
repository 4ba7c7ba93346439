from shapefromprojections_b200.dataset.dataset import SinoDataset  # noqa: F401
