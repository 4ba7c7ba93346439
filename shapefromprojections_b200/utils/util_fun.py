"""Edge-length regulariser (``ctdr/utils/util_fun.py:5-19``): mean squared edge length / 2."""
import torch


def compute_edge_length(vertices, faces):
    a, b, c = (vertices.index_select(0, faces[:, k]) for k in range(3))
    sq = lambda e: (e * e).sum(1).mean()
    return (sq(b - a) + sq(c - a) + sq(b - c)) / 6.0
