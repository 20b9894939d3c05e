"""MediaPipe Pose 33-landmark spanning tree rooted at the nose (reference: graph/mediapipe_pose.py:6-45)."""
from ._base import SkeletonGraph

num_node = 33
# (child, parent) pairs; the 0->11 and 0->9 bridges join MediaPipe's three disconnected components
_PARENT = {1: 0, 2: 1, 3: 2, 7: 3, 4: 0, 5: 4, 6: 5, 8: 6, 9: 0, 10: 9, 11: 0, 12: 11, 13: 11, 15: 13, 17: 15, 19: 15,
           21: 15, 14: 12, 16: 14, 18: 16, 20: 16, 22: 16, 23: 11, 24: 12, 25: 23, 27: 25, 29: 27, 31: 27, 26: 24,
           28: 26, 30: 28, 32: 28}
inward = list(_PARENT.items())
outward = [(j, i) for (i, j) in inward]
self_link = [(i, i) for i in range(num_node)]
neighbor = inward + outward


class Graph(SkeletonGraph):
    num_node = num_node
    inward = tuple(inward)
