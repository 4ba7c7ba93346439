from shapefromprojections_b200.function.rasterizer import *  # noqa: F401,F403
from shapefromprojections_b200.function.rasterizer import Rasterizer, dtype_int  # noqa: F401
