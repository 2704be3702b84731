"""Parameter initialisation of the heatmap columns / combiners (mirror of
/root/reference/src/margipose/nn_helpers.py:7-21): Kaiming-normal (fan_out) convolutions,
BatchNorm gamma = 1, beta = 0 (the reference's nn.Linear branch has no module to act on here).
Host-side, runs once at model construction."""
from torch import nn
from torch.nn import init
from torch.nn.modules.conv import _ConvNd


def init_parameters(net):
    for m in net.modules():
        if isinstance(m, _ConvNd):
            init.kaiming_normal_(m.weight, 0, 'fan_out')
            if m.bias is not None:
                init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            init.constant_(m.weight, 1)
            if m.bias is not None:
                init.constant_(m.bias, 0)
