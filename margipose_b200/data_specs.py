"""Input/output specification objects a MargiPose model carries (`model.data_specs`).

Mirrors the attribute surface of /root/reference/src/margipose/data_specs.py:26-64 that callers
of the model read (resolution, ImageNet mean / stddev, skeleton).  PIL image conversion is the
data pipeline's job and is out of the hot-path scope.
"""
from collections.abc import Sequence


class ImageSpecs:
    IMAGENET_MEAN = [0.485, 0.456, 0.406]
    IMAGENET_STDDEV = [0.229, 0.224, 0.225]

    def __init__(self, resolution, mean=None, stddev=None):
        if isinstance(resolution, Sequence):
            self.height, self.width = resolution
        else:
            self.height = self.width = resolution
        self.mean = mean
        self.stddev = stddev


class JointsSpecs:
    def __init__(self, skeleton_desc, n_dims=3):
        self.skeleton_desc = skeleton_desc
        self.n_dims = n_dims


class DataSpecs:
    def __init__(self, input_specs, output_specs):
        self._input_specs = input_specs
        self._output_specs = output_specs

    @property
    def input_specs(self):
        return self._input_specs

    @property
    def output_specs(self):
        return self._output_specs
