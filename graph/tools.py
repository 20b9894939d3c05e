"""Adjacency helpers for the skeleton graphs (API-compatible with the reference's graph/tools.py:4-27).

The adjacency is accepted by the model constructors but never used by any forward pass
(SURVEY.md App. E-1); it is kept so that ``Model(graph='graph.ntu_rgb_d.Graph', ...)`` keeps working.
"""
import numpy as np


def edge2mat(link, num_node):
    A = np.zeros((num_node, num_node))
    if len(link):
        src, dst = np.asarray(link, dtype=np.int64).T
        A[dst, src] = 1
    return A


def normalize_digraph(A):
    """column-normalise: divide every column by its sum (columns that sum to zero stay zero)"""
    col = A.sum(0)
    inv = np.divide(1.0, col, out=np.zeros_like(col, dtype=np.float64), where=col > 0)
    return A * inv[None, :]


def get_spatial_graph(num_node, self_link, inward, outward):
    return np.stack((edge2mat(self_link, num_node),
                     normalize_digraph(edge2mat(inward, num_node)),
                     normalize_digraph(edge2mat(outward, num_node))))
