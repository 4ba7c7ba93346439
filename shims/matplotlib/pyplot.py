"""No-op pyplot (see shims/README.md): run/simulation_cone_beam.py:85-91, ctdr/utils/util_vis.py."""


def _noop(*args, **kwargs):
    return None


imshow = colorbar = show = plot = savefig = imsave = figure = close = clf = subplot = title = axis = _noop
