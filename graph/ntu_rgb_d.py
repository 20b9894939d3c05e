"""NTU RGB+D 25-joint skeleton (reference: graph/ntu_rgb_d.py:6-33)."""
from ._base import SkeletonGraph

num_node = 25
# parent of joint j (1-based, as in the NTU documentation); joint 21 (spine shoulder) is the root
_PARENT_1BASED = {1: 2, 2: 21, 3: 21, 4: 3, 5: 21, 6: 5, 7: 6, 8: 7, 9: 21, 10: 9, 11: 10, 12: 11, 13: 1, 14: 13,
                  15: 14, 16: 15, 17: 1, 18: 17, 19: 18, 20: 19, 22: 23, 23: 8, 24: 25, 25: 12}
inward = [(c - 1, p - 1) for c, p in sorted(_PARENT_1BASED.items())]
outward = [(j, i) for (i, j) in inward]
self_link = [(i, i) for i in range(num_node)]
neighbor = inward + outward


class Graph(SkeletonGraph):
    num_node = num_node
    inward = tuple(inward)
