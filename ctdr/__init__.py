"""Drop-in import surface of the reference package ``ctdr`` (jakeoung/ShapeFromProjections),
backed by ``shapefromprojections_b200``.  ``from ctdr.model.vanilla import Model`` etc. keep working."""
