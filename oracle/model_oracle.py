"""Oracle for the MargiPose network body (ResNet stem + stacked heatmap columns).

TEST INFRASTRUCTURE (see oracle/__init__.py).  A plain-PyTorch CPU restatement of
/root/reference/src/margipose/models/margipose_model.py with the SAME module
tree / state_dict key names, so `load_state_dict(reference.state_dict())` works
(pinned in tests/test_oracle_pin.py).  The truncated torchvision ResNet the
reference borrows (margipose_model.py:119-138) is restated here as well so the
oracle does not depend on torchvision.

`OracleMargiPose(..., emulate_bf16=True)` additionally rounds to bfloat16 at the
points where the CUDA path stores bf16 (DESIGN.md "Precision"): conv operands,
conv outputs, normalised activations.  Accumulation stays fp32.  That mode is the
tight-tolerance checker for the tensor-core path; the default fp32 mode is the
reference-exact one.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import dsnt_oracle as D

JOINT_NAMES = [  # data/skeleton.py:51-74 (CanonicalSkeletonDesc); index == heatmap channel
    'head_top', 'neck', 'right_shoulder', 'right_elbow', 'right_wrist',
    'left_shoulder', 'left_elbow', 'left_wrist', 'right_hip', 'right_knee',
    'right_ankle', 'left_hip', 'left_knee', 'left_ankle', 'pelvis',
    'spine', 'head',
]
JOINT_TREE = [1, 15, 1, 2, 3, 1, 5, 6, 14, 8, 9, 14, 11, 12, 14, 14, 1]
HFLIP_INDICES = [0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 14, 15, 16]
N_JOINTS = len(JOINT_NAMES)


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _Numerics:
    """Executes conv / batch-norm modules either exactly (fp32) or with bf16 rounding."""

    def __init__(self, emulate_bf16):
        self.emulate = emulate_bf16
        self.trace = None      # set to a list to record every block output (layer-wise parity)
        # inference path of the CUDA engine: eval-mode BatchNorm rides in the conv epilogue, so the conv
        # output is NOT rounded to bf16 before the affine (only after it) for the convs marked fold=True
        self.fold_eval = True

    def rec(self, x):
        if self.trace is not None:
            self.trace.append(x.detach())
        return x

    def q(self, x):
        return _bf16(x) if self.emulate else x

    def conv(self, m, x, fold=False):
        if not self.emulate:
            return m(x)
        w = _bf16(m.weight)
        if isinstance(m, nn.ConvTranspose2d):
            y = F.conv_transpose2d(x, w, m.bias, m.stride, m.padding, m.output_padding)
        else:
            y = F.conv2d(x, w, m.bias, m.stride, m.padding)
        if fold and self.fold_eval and not m.training:
            return y
        return _bf16(y)

    def bn(self, m, x):
        """Returns the normalised tensor WITHOUT rounding (callers round after fusing)."""
        if not self.emulate:
            return m(x)
        if m.training:
            mean = x.mean((0, 2, 3))
            var = x.var((0, 2, 3), unbiased=False)
            with torch.no_grad():
                n = x.numel() // x.shape[1]
                m.running_mean.mul_(1 - m.momentum).add_(mean * m.momentum)
                m.running_var.mul_(1 - m.momentum).add_(var * (n / max(n - 1, 1)) * m.momentum)
                m.num_batches_tracked += 1
        else:
            mean, var = m.running_mean, m.running_var
        scale = m.weight * torch.rsqrt(var + m.eps)
        shift = m.bias - mean * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def init_parameters(net):
    """Restates nn_helpers.py:7-21 (Kaiming-normal fan_out convs, BN gamma=1 beta=0)."""
    for m in net.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(m.weight, 0, 'fan_out')
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class ResidualBlock(nn.Module):
    """margipose_model.py:25-40: relu(bn(conv3x3(relu(bn(conv_in(x)))))) + bn(conv_sc(x))."""

    def __init__(self, chans, main_conv_in, shortcut_conv_in):
        super().__init__()
        self.module = nn.Sequential(
            main_conv_in, nn.BatchNorm2d(chans), nn.ReLU(inplace=True),
            nn.Conv2d(chans, chans, 3, padding=1, bias=False), nn.BatchNorm2d(chans),
            nn.ReLU(inplace=True))
        self.shortcut = nn.Sequential(shortcut_conv_in, nn.BatchNorm2d(chans))

    def run(self, nm, x, keep_fp32=False):
        # folded inference (engine.py residual_block): not for the logits block, nor for a 1x1 transposed shortcut
        fold = not keep_fp32 and not (isinstance(self.shortcut[0], nn.ConvTranspose2d))
        a = nm.q(F.relu(nm.bn(self.module[1], nm.conv(self.module[0], x, fold))))
        short = nm.bn(self.shortcut[1], nm.conv(self.shortcut[0], x, fold))
        if fold and nm.emulate and nm.fold_eval and not self.training:
            short = nm.q(short)        # the shortcut branch is stored before the main conv adds it
        main = F.relu(nm.bn(self.module[4], nm.conv(self.module[3], a, fold)))
        out = main + short
        return nm.rec(out if keep_fp32 else nm.q(out))


def _block(cin, cout, kind):
    if kind == 'regular':
        return ResidualBlock(cout, nn.Conv2d(cin, cout, 3, padding=1, bias=False),
                             nn.Conv2d(cin, cout, 1, bias=False))
    if kind == 'down':
        return ResidualBlock(cout, nn.Conv2d(cin, cout, 3, padding=1, stride=2, bias=False),
                             nn.Conv2d(cin, cout, 1, stride=2, bias=False))
    return ResidualBlock(
        cout,
        nn.ConvTranspose2d(cin, cout, 3, padding=1, stride=2, output_padding=1, bias=False),
        nn.ConvTranspose2d(cin, cout, 1, stride=2, output_padding=1, bias=False))


def permute_axes(mid, space):
    """The channel-group <-> spatial axis swap of margipose_model.py:86-99.

    With S = W = H of `mid` and channels viewed as (G, S):
      zy: out[b, g, w, h, c] = in[b, g, c, h, w]   (swap channel-in-group with W)
      xz: out[b, g, h, c, w] = in[b, g, c, h, w]   (swap channel-in-group with H)
    """
    if space == 'xy':
        return mid
    b, c, h, w = mid.shape
    s = w
    v = mid.reshape(b, c // s, s, h, w)
    if space == 'zy':
        v = v.permute(0, 1, 4, 3, 2)
    elif space == 'xz':
        v = v.permute(0, 1, 3, 2, 4)
    else:
        raise Exception()
    return v.reshape(b, c, h, w)


class HeatmapColumn(nn.Module):
    """margipose_model.py:43-100."""

    def __init__(self, n_joints, heatmap_space):
        super().__init__()
        self.n_joints = n_joints
        self.heatmap_space = heatmap_space
        self.down_layers = nn.Sequential(
            _block(128, 128, 'regular'), _block(128, 128, 'regular'), _block(128, 192, 'down'),
            _block(192, 192, 'regular'), _block(192, 192, 'regular'))
        self.up_layers = nn.Sequential(
            _block(192, 192, 'regular'), _block(192, 192, 'regular'), _block(192, 128, 'up'),
            _block(128, 128, 'regular'), _block(128, n_joints, 'regular'))
        init_parameters(self)

    def run(self, nm, x):
        for blk in self.down_layers:
            x = blk.run(nm, x)
        x = permute_axes(x, self.heatmap_space)
        n = len(self.up_layers)
        for i, blk in enumerate(self.up_layers):
            x = blk.run(nm, x, keep_fp32=(i == n - 1))   # logits stay fp32
        return x


class BasicBlock(nn.Module):
    """torchvision ResNet-18/34 block (used by margipose_model.py:120,130-135)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                            nn.BatchNorm2d(cout))
        else:
            self.downsample = None

    def run(self, nm, x):
        a = nm.q(F.relu(nm.bn(self.bn1, nm.conv(self.conv1, x, True))))
        main = nm.bn(self.bn2, nm.conv(self.conv2, a, True))
        if self.downsample is not None:
            ident = nm.bn(self.downsample[1], nm.conv(self.downsample[0], x, True))
            if nm.emulate and nm.fold_eval and not self.training:
                ident = nm.q(ident)
        else:
            ident = x
        return nm.rec(nm.q(F.relu(main + ident)))


class Bottleneck(nn.Module):
    """torchvision ResNet-50 block (v1.5: the stride sits on the 3x3 conv)."""

    def __init__(self, cin, width, stride):
        super().__init__()
        cout = width * 4
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                            nn.BatchNorm2d(cout))
        else:
            self.downsample = None

    def run(self, nm, x):
        a = nm.q(F.relu(nm.bn(self.bn1, nm.conv(self.conv1, x, True))))
        b = nm.q(F.relu(nm.bn(self.bn2, nm.conv(self.conv2, a, True))))
        main = nm.bn(self.bn3, nm.conv(self.conv3, b, True))
        if self.downsample is not None:
            ident = nm.bn(self.downsample[1], nm.conv(self.downsample[0], x, True))
            if nm.emulate and nm.fold_eval and not self.training:
                ident = nm.q(ident)
        else:
            ident = x
        return nm.rec(nm.q(F.relu(main + ident)))


_RESNETS = {  # name -> (block, layer1 count, layer2 count, layer2 output channels)
    'resnet18': (BasicBlock, 2, 2, 128),
    'resnet34': (BasicBlock, 3, 4, 128),
    'resnet50': (Bottleneck, 3, 4, 512),
}


def _resnet_init(net):
    # torchvision.models.resnet.ResNet.__init__: kaiming_normal_(fan_out, relu); BN 1/0
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def make_image_feature_extractor(model_name):
    """margipose_model.py:103-139, resnet branch (inceptionv4 needs `pretrainedmodels`,
    which is not part of the hot-path scope -- SURVEY.md section 8c)."""
    if model_name not in _RESNETS:
        raise Exception('unsupported image feature extractor model name: ' + model_name)
    block, n1, n2, out_chans = _RESNETS[model_name]
    if block is BasicBlock:
        layer1 = nn.Sequential(*[BasicBlock(64, 64, 1) for _ in range(n1)])
        layer2 = nn.Sequential(*[BasicBlock(64 if i == 0 else 128, 128, 2 if i == 0 else 1)
                                 for i in range(n2)])
    else:
        layer1 = nn.Sequential(*[Bottleneck(64 if i == 0 else 256, 64, 1) for i in range(n1)])
        layer2 = nn.Sequential(*[Bottleneck(256 if i == 0 else 512, 128, 2 if i == 0 else 1)
                                 for i in range(n2)])
    stem = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
            nn.MaxPool2d(3, 2, 1), layer1, layer2]
    net = nn.Sequential(*stem)
    _resnet_init(net)
    if out_chans != 128:
        net = nn.Sequential(*stem, nn.Conv2d(out_chans, 128, 1), nn.BatchNorm2d(128),
                            nn.ReLU(inplace=True))
    return net


def run_feature_extractor(nm, net, x):
    x = nm.q(x)
    x = nm.rec(nm.q(F.relu(nm.bn(net[1], nm.conv(net[0], x, True)))))
    x = nm.rec(net[3](x))
    for blk in net[4]:
        x = blk.run(nm, x)
    for blk in net[5]:
        x = blk.run(nm, x)
    if len(net) > 6:
        x = nm.rec(nm.q(F.relu(nm.bn(net[7], nm.conv(net[6], x, True)))))
    return x


class HeatmapCombiner(nn.Module):
    """margipose_model.py:142-150."""

    def __init__(self, n_joints):
        super().__init__()
        self.conv = nn.Conv2d(n_joints * 3, 128, 1, bias=False)
        init_parameters(self)

    def forward(self, xy, zy, xz):
        return self.conv(torch.cat([xy, zy, xz], -3))


class OracleInner(nn.Module):
    """margipose_model.py:153-200."""

    def __init__(self, n_joints, n_stages, axis_permutation, feature_extractor):
        super().__init__()
        self.n_stages = n_stages
        self.in_cnn = make_image_feature_extractor(feature_extractor)
        self.xy_hm_cnns = nn.ModuleList()
        self.zy_hm_cnns = nn.ModuleList()
        self.xz_hm_cnns = nn.ModuleList()
        self.hm_combiners = nn.ModuleList()
        zy, xz = ('zy', 'xz') if axis_permutation else ('xy', 'xy')
        for t in range(n_stages):
            if t > 0:
                self.hm_combiners.append(HeatmapCombiner(n_joints))
            self.xy_hm_cnns.append(HeatmapColumn(n_joints, 'xy'))
            self.zy_hm_cnns.append(HeatmapColumn(n_joints, zy))
            self.xz_hm_cnns.append(HeatmapColumn(n_joints, xz))


class OracleMargiPose(nn.Module):
    """margipose_model.py:203-267, same public surface (forward, heatmap lists, losses)."""

    def __init__(self, n_stages=4, axis_permutation=True, feature_extractor='resnet34',
                 pixelwise_loss='jsd', emulate_bf16=False):
        super().__init__()
        self.pixelwise_loss = pixelwise_loss
        self.inner = OracleInner(N_JOINTS, n_stages, axis_permutation, feature_extractor)
        self.xy_heatmaps = self.zy_heatmaps = self.xz_heatmaps = None
        self.logits = None
        self.nm = _Numerics(emulate_bf16)

    def forward(self, x):
        nm, inner = self.nm, self.inner
        feats = run_feature_extractor(nm, inner.in_cnn, x)
        xy_hms, zy_hms, xz_hms, logits = [], [], [], []
        inp = feats
        for t in range(inner.n_stages):
            if t > 0:
                comb = inner.hm_combiners[t - 1](xy_hms[-1], zy_hms[-1], xz_hms[-1])
                inp = nm.q(inp + comb)
            lz = [inner.xy_hm_cnns[t].run(nm, inp), inner.zy_hm_cnns[t].run(nm, inp),
                  inner.xz_hm_cnns[t].run(nm, inp)]
            logits.append(lz)
            xy_hms.append(D.flat_softmax(lz[0]))
            zy_hms.append(D.flat_softmax(lz[1]))
            xz_hms.append(D.flat_softmax(lz[2]))
        self.logits = logits
        self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps = xy_hms, zy_hms, xz_hms
        return D.heatmaps_to_coords(xy_hms[-1], zy_hms[-1], xz_hms[-1])

    heatmaps_to_coords = staticmethod(D.heatmaps_to_coords)

    def forward_3d_losses(self, out_var, target_var):
        return D.losses_3d(self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps, target_var,
                           self.pixelwise_loss)

    def forward_2d_losses(self, out_var, target_var):
        return D.losses_2d(self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps, target_var,
                           self.pixelwise_loss)


def create_oracle(model_desc, emulate_bf16=False):
    """Mirror of models/__init__.py:16-27 + margipose_model.py:270-284 for the oracle."""
    if model_desc['type'] != 'margipose' or not str(model_desc['version']).startswith('6.'):
        raise Exception('unrecognised model {} v{}'.format(model_desc['type'],
                                                           model_desc['version']))
    s = model_desc['settings']
    return OracleMargiPose(n_stages=s.get('n_stages', 4),
                           axis_permutation=s.get('axis_permutation', True),
                           feature_extractor=s.get('feature_extractor', 'inceptionv4'),
                           pixelwise_loss=s.get('pixelwise_loss', 'jsd'),
                           emulate_bf16=emulate_bf16)
