"""Sinogram snapshots (``ctdr/utils/util_vis.py:49-83``); needs matplotlib, silently skipped without it."""
import numpy as np


def save_sino_as_img(fname, sino, cmap="gray", nshow=4):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return False
    sino = sino.detach().cpu().numpy() if hasattr(sino, "detach") else np.asarray(sino)
    step = max(1, sino.shape[0] // nshow)
    for k, i in enumerate(range(0, sino.shape[0], step)[:nshow]):
        plt.imsave(fname.replace(".png", f"_{k}.png"), sino[i], cmap=cmap)
    return True
