"""Skeleton descriptor of the 17 canonical joints -- joint index j is heatmap channel j.

Mirrors `CanonicalSkeletonDesc` of /root/reference/src/margipose/data/skeleton.py:10-74 (names,
parent tree, horizontal-flip map); pure bookkeeping, bit-exact by construction and pinned by
tests/test_host_logic.py against the oracle's restatement.
"""


class SkeletonDesc:
    def __init__(self, joint_names, joint_tree, hflip_indices):
        self.joint_names = joint_names
        self.joint_tree = joint_tree
        self.hflip_indices = hflip_indices

    @property
    def n_joints(self):
        return len(self.joint_names)

    @property
    def canonical(self):
        return True

    @property
    def root_joint_id(self):
        return self.joint_names.index('pelvis')

    def to_dict(self):
        return {'joint_names': self.joint_names, 'joint_tree': self.joint_tree,
                'hflip_indices': self.hflip_indices}

    @classmethod
    def from_dict(cls, d):
        return cls(d['joint_names'], d['joint_tree'], d['hflip_indices'])


CanonicalSkeletonDesc = SkeletonDesc(
    joint_names=['head_top', 'neck', 'right_shoulder', 'right_elbow', 'right_wrist',
                 'left_shoulder', 'left_elbow', 'left_wrist', 'right_hip', 'right_knee',
                 'right_ankle', 'left_hip', 'left_knee', 'left_ankle', 'pelvis', 'spine', 'head'],
    joint_tree=[1, 15, 1, 2, 3, 1, 5, 6, 14, 8, 9, 14, 11, 12, 14, 14, 1],
    hflip_indices=[0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 14, 15, 16],
)
