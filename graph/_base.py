from . import tools


class SkeletonGraph:
    """Common behaviour of the reference's Graph classes (graph/ntu_rgb_d.py:17-33, graph/mediapipe_pose.py:29-45)."""

    num_node = 0
    inward = ()

    def __init__(self, labeling_mode='spatial'):
        self.self_link = [(i, i) for i in range(self.num_node)]
        self.inward = list(type(self).inward)
        self.outward = [(j, i) for (i, j) in self.inward]
        self.neighbor = self.inward + self.outward
        self.A = self.get_adjacency_matrix(labeling_mode)

    def get_adjacency_matrix(self, labeling_mode=None):
        if labeling_mode is None:
            return self.A
        if labeling_mode == 'spatial':
            return tools.get_spatial_graph(self.num_node, self.self_link, self.inward, self.outward)
        raise ValueError()
