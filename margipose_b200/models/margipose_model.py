"""MargiPose model on the sm_100a engine -- drop-in for
/root/reference/src/margipose/models/margipose_model.py.

Same public surface: `MargiPoseModel.forward(x) -> (B, 17, 3)`, the
`xy_heatmaps / zy_heatmaps / xz_heatmaps` lists refreshed by every forward,
`forward_3d_losses`, `forward_2d_losses`, static `heatmaps_to_coords`, `data_specs`, and the
nn.Module protocol with the reference's state_dict keys and shapes (the module tree below only
HOLDS parameters; it is never executed with torch ops).  Compute goes through the C ABI
(include/margipose_b200.h): tcgen05 implicit-GEMM convolutions, fused BatchNorm/ReLU/residual
passes and the fused softmax / soft-argmax / JS tail.  There is no CPU or PyTorch fallback: CPU
inputs raise.

Differences a caller can observe, all forced by the environment or the precision contract:
  * conv operands are bf16 (fp32 accumulate, fp32 BatchNorm statistics, fp32 heatmaps/losses);
  * ImageNet-pretrained ResNet weights are not downloaded (no network) -- the stem starts from
    torchvision's own random initialisation; load a checkpoint with load_state_dict as usual;
  * the 'inceptionv4' stem (third-party `pretrainedmodels`) is outside the hot-path scope;
  * gradients are written straight into one flat buffer that every `param.grad` aliases.
"""
import torch
from torch import nn

from .. import dsntnn as K
from .._lib import require_cuda, MargiposeB200Error
from ..data_specs import DataSpecs, ImageSpecs, JointsSpecs
from ..engine import ParamBank, Engine, build_layers
from ..model_factory import ModelFactory
from ..nn_helpers import init_parameters
from ..skeleton import CanonicalSkeletonDesc

Default_MargiPose_Desc = {
    'type': 'margipose',
    'version': '6.0.1',
    'settings': {
        'n_stages': 4,
        'axis_permutation': True,
        'feature_extractor': 'inceptionv4',
        'pixelwise_loss': 'jsd',
    },
}


class _Holder(nn.Module):
    """A module that only owns parameters; the engine executes the network."""

    def forward(self, *inputs):
        raise MargiposeB200Error(
            '%s is a parameter holder; run the whole MargiPoseModel (the CUDA engine executes '
            'it) -- there is no per-module PyTorch path' % type(self).__name__)


class ResidualBlock(_Holder):
    """margipose_model.py:25-40: relu(bn(conv3x3(relu(bn(conv_in(x)))))) + bn(conv_sc(x))."""

    def __init__(self, chans, main_conv_in, shortcut_conv_in):
        super().__init__()
        assert main_conv_in.in_channels == shortcut_conv_in.in_channels
        self.module = nn.Sequential(
            main_conv_in,
            nn.BatchNorm2d(chans),
            nn.ReLU(inplace=True),
            nn.Conv2d(chans, chans, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(chans),
            nn.ReLU(inplace=True),
        )
        self.shortcut = nn.Sequential(shortcut_conv_in, nn.BatchNorm2d(chans))


class HeatmapColumn(_Holder):
    """margipose_model.py:43-100."""

    def __init__(self, n_joints, heatmap_space):
        super().__init__()
        if heatmap_space not in ('xy', 'zy', 'xz'):
            raise Exception()
        self.n_joints = n_joints
        self.heatmap_space = heatmap_space
        self.down_layers = nn.Sequential(
            self._regular_block(128, 128),
            self._regular_block(128, 128),
            self._down_stride_block(128, 192),
            self._regular_block(192, 192),
            self._regular_block(192, 192),
        )
        self.up_layers = nn.Sequential(
            self._regular_block(192, 192),
            self._regular_block(192, 192),
            self._up_stride_block(192, 128),
            self._regular_block(128, 128),
            self._regular_block(128, self.n_joints),
        )
        init_parameters(self)

    def _regular_block(self, in_chans, out_chans):
        return ResidualBlock(
            out_chans,
            nn.Conv2d(in_chans, out_chans, kernel_size=3, padding=1, bias=False),
            nn.Conv2d(in_chans, out_chans, kernel_size=1, bias=False))

    def _down_stride_block(self, in_chans, out_chans):
        return ResidualBlock(
            out_chans,
            nn.Conv2d(in_chans, out_chans, kernel_size=3, padding=1, stride=2, bias=False),
            nn.Conv2d(in_chans, out_chans, kernel_size=1, stride=2, bias=False))

    def _up_stride_block(self, in_chans, out_chans):
        return ResidualBlock(
            out_chans,
            nn.ConvTranspose2d(in_chans, out_chans, kernel_size=3, padding=1, stride=2,
                               output_padding=1, bias=False),
            nn.ConvTranspose2d(in_chans, out_chans, kernel_size=1, stride=2,
                               output_padding=1, bias=False))


class _BasicBlock(_Holder):
    """Parameter layout of torchvision.models.resnet.BasicBlock."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                            nn.BatchNorm2d(cout))


class _Bottleneck(_Holder):
    """Parameter layout of torchvision.models.resnet.Bottleneck (v1.5: stride on the 3x3)."""

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
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                            nn.BatchNorm2d(cout))


_RESNETS = {  # name -> (block, blocks in layer1, blocks in layer2, channels after layer2)
    'resnet18': (_BasicBlock, 2, 2, 128),
    'resnet34': (_BasicBlock, 3, 4, 128),
    'resnet50': (_Bottleneck, 3, 4, 512),
}


def make_image_feature_extractor(model_name):
    """margipose_model.py:103-139, ResNet branch: conv1/bn1/relu/maxpool/layer1/layer2 of a
    torchvision ResNet (+ 1x1 conv / BN / ReLU to 128 channels when the backbone is wider)."""
    if model_name == 'inceptionv4':
        raise Exception('unsupported image feature extractor model name: inceptionv4 (the third-party '
                        '`pretrainedmodels` stem is outside the B200 hot-path scope; use resnet18, '
                        'resnet34 or resnet50)')
    if model_name not in _RESNETS:
        raise Exception('unsupported image feature extractor model name: ' + model_name)
    block, n1, n2, out_chans = _RESNETS[model_name]
    if block is _BasicBlock:
        layer1 = nn.Sequential(*[_BasicBlock(64, 64, 1) for _ in range(n1)])
        layer2 = nn.Sequential(*[_BasicBlock(64 if i == 0 else 128, 128, 2 if i == 0 else 1)
                                 for i in range(n2)])
    else:
        layer1 = nn.Sequential(*[_Bottleneck(64 if i == 0 else 256, 64, 1) for i in range(n1)])
        layer2 = nn.Sequential(*[_Bottleneck(256 if i == 0 else 512, 128, 2 if i == 0 else 1)
                                 for i in range(n2)])
    stem = [nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64),
            nn.ReLU(inplace=True), nn.MaxPool2d(kernel_size=3, stride=2, padding=1), layer1, layer2]
    for m in nn.Sequential(*stem).modules():   # torchvision.models.resnet.ResNet.__init__
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
    extra = []
    if out_chans != 128:
        extra = [nn.Conv2d(out_chans, 128, 1), nn.BatchNorm2d(128), nn.ReLU(inplace=True)]
    return nn.Sequential(*stem, *extra)


class HeatmapCombiner(_Holder):
    """margipose_model.py:142-150."""

    def __init__(self, n_joints):
        super().__init__()
        self.conv = nn.Conv2d(n_joints * 3, 128, kernel_size=1, bias=False)
        init_parameters(self)


class MargiPoseModelInner(_Holder):
    """margipose_model.py:153-200 (structure only)."""

    def __init__(self, n_joints, n_stages, axis_permutation, feature_extractor):
        super().__init__()
        self.n_stages = n_stages
        self.in_cnn = make_image_feature_extractor(feature_extractor)
        self.xy_hm_cnns = nn.ModuleList()
        self.zy_hm_cnns = nn.ModuleList()
        self.xz_hm_cnns = nn.ModuleList()
        self.hm_combiners = nn.ModuleList()
        xy = 'xy'
        if axis_permutation:
            zy, xz = 'zy', 'xz'
        else:
            zy, xz = 'xy', 'xy'
        for t in range(self.n_stages):
            if t > 0:
                self.hm_combiners.append(HeatmapCombiner(n_joints))
            self.xy_hm_cnns.append(HeatmapColumn(n_joints, heatmap_space=xy))
            self.zy_hm_cnns.append(HeatmapColumn(n_joints, heatmap_space=zy))
            self.xz_hm_cnns.append(HeatmapColumn(n_joints, heatmap_space=xz))


def utils_default_deterministic():
    from .. import utils
    return utils.DETERMINISTIC


class _Body(torch.autograd.Function):
    """The whole network body as ONE autograd node: image -> 3 * n_stages probability heatmaps.
    Parameter gradients are accumulated by the engine straight into the flat gradient buffer
    that every `param.grad` aliases (they are not autograd inputs)."""

    @staticmethod
    def forward(ctx, eng, x, _anchor):
        probs = eng.forward(x)
        ctx.eng, ctx.generation = eng, eng.generation
        ctx.set_materialize_grads(False)
        return tuple(p.detach() for row in probs for p in row)

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.eng
        if eng.generation != ctx.generation:
            # one set of activation buffers per (batch, resolution, mode): a later forward of the same shape
            # has overwritten what this backward needs (and the heatmaps this forward returned)
            raise MargiposeB200Error(
                'backward through a MargiPose forward whose activations were overwritten by a later forward '
                'of the same shape; call backward before the next forward (gradient accumulation: one '
                'forward + backward per micro-batch)')
        eng.bank.attach_grads()
        eng.backward([[grads[3 * t + k] for k in range(3)] for t in range(len(grads) // 3)])
        return None, None, None


class MargiPoseModel(nn.Module):
    PRECISIONS = ('bf16', 'bf16x3')

    def __init__(self, skel_desc, n_stages, axis_permutation, feature_extractor, pixelwise_loss, precision='bf16'):
        super().__init__()
        if precision not in self.PRECISIONS:
            raise ValueError('precision must be one of %s' % (self.PRECISIONS,))
        # 'bf16'  : bf16 tensor-core operands and bf16 activation storage, fp32 accumulation (the fast path)
        # 'bf16x3': every activation / weight is a pair of bf16 values (hi + lo, ~16 significant bits) and every
        #           convolution is three tensor-core passes over one fp32 accumulator -- results agree with the
        #           reference's fp32 arithmetic to ~2e-4 (PARITY.md) at a bit under half the speed
        self.precision = precision
        # run-to-run bit-wise reproducible training steps (the reference's `deterministic` flag, utils.py:19-24):
        # set before the first forward, or call drop_engines() after changing it
        self.deterministic = utils_default_deterministic()
        self.data_specs = DataSpecs(
            ImageSpecs(256, mean=ImageSpecs.IMAGENET_MEAN, stddev=ImageSpecs.IMAGENET_STDDEV),
            JointsSpecs(skel_desc, n_dims=3),
        )
        self.pixelwise_loss = pixelwise_loss
        self.n_joints = skel_desc.n_joints
        self.inner = MargiPoseModelInner(skel_desc.n_joints, n_stages, axis_permutation,
                                         feature_extractor)
        self.xy_heatmaps = self.zy_heatmaps = self.xz_heatmaps = None
        self._bank = None
        self._layers = None
        self._engines = {}
        self._packed_version = None

    # ---- engine plumbing
    def set_precision(self, precision):
        """Switches the arithmetic of the CUDA engine ('bf16' | 'bf16x3'); parameters and buffers are kept."""
        if precision not in self.PRECISIONS:
            raise ValueError('precision must be one of %s' % (self.PRECISIONS,))
        self.precision = precision
        return self

    def _materialize(self, device):
        self._bank = ParamBank(split=self.precision == 'bf16x3')
        self._layers = build_layers(self, self._bank)
        self._bank.finalize(device)
        self._engines = {}
        self._packed_version = None

    def _ensure(self, device):
        if self._bank is None or self._bank.device != device or not self._bank.linked() or \
                self._bank.split != (self.precision == 'bf16x3'):
            self._materialize(device)

    def engine_for(self, n, h, w, training):
        key = (n, h, w, bool(training))
        eng = self._engines.get(key)
        if eng is None:
            if h % 16 != 0 or w % 16 != 0:
                raise MargiposeB200Error('input height and width must be multiples of 16 (got %dx%d)' % (h, w))
            eng = Engine(self, n, h, w, bool(training), self._bank.device)
            self._engines[key] = eng
        return eng

    def drop_engines(self):
        """Frees the activation buffers / launch programs of every (batch, resolution, mode) seen so far."""
        self._engines = {}

    def mark_params_dirty(self):
        """Call after writing parameters through raw device pointers (the flat optimiser does)."""
        self._packed_version = None

    def _refresh_packs(self):
        """bf16 GEMM operands follow the fp32 master weights: every training forward repacks (one launch);
        in eval mode the packs are reused until a parameter is written -- through the flat buffer (FlatSGD
        calls mark_params_dirty) or through any nn.Parameter (load_state_dict, torch.optim, p.add_)."""
        if self.training:
            self._bank.pack()
            self._packed_version = None
            return
        v = self._bank.param_version()
        if self._packed_version != v:
            self._bank.pack()
            self._packed_version = v

    @property
    def flat_params(self):
        return self._bank.flat

    @property
    def flat_grads(self):
        return self._bank.flat_grad

    def _inner_forward(self, x):
        require_cuda(x)
        if x.dtype == torch.uint8:
            # raw image batch (B, H, W, 3) as a decoder produces it: /255 and the ImageNet normalisation of
            # data_specs (ImageSpecs.convert, data_specs.py:38-39) are fused into the stem conv's gather
            if x.dim() != 4 or x.size(3) != 3:
                raise ValueError('expected a uint8 (B, H, W, 3) image batch, got %s' % (tuple(x.shape),))
            x = x.contiguous()
            n, h, w = x.size(0), x.size(1), x.size(2)
        else:
            if x.dim() != 4 or x.size(1) != 3:
                raise ValueError('expected a (B, 3, H, W) image batch, got %s' % (tuple(x.shape),))
            x = x.float().contiguous()
            n, h, w = x.size(0), x.size(2), x.size(3)
        self._ensure(x.device)
        eng = self.engine_for(n, h, w, self.training)
        self._refresh_packs()
        if self.training and torch.is_grad_enabled():
            # A fresh leaf ties the node into the autograd graph.  (Using a long-lived Parameter here
            # would pin the node's leaf stream to whichever stream first used that Parameter -- usually
            # the legacy default stream -- which autograd then touches at the end of backward and
            # thereby invalidates a CUDA-graph capture of the training step.)
            anchor = torch.empty(0, device=x.device, requires_grad=True)
            flat = _Body.apply(eng, x, anchor)
        else:
            flat = [p for row in eng.forward(x) for p in row]
        n = len(flat) // 3
        return ([flat[3 * t] for t in range(n)], [flat[3 * t + 1] for t in range(n)],
                [flat[3 * t + 2] for t in range(n)])

    # ---- the reference surface
    def _pixelwise_flag(self):
        if self.pixelwise_loss == 'jsd':
            return True
        elif self.pixelwise_loss is None:
            return False
        raise Exception('unrecognised pixelwise loss: {}'.format(self.pixelwise_loss))

    def forward_2d_losses(self, out_var, target_var):
        """margipose_model.py:223-234: per stage JS(xy) + |xy - target_xy|."""
        pixelwise = self._pixelwise_flag()
        target_xy = target_var.narrow(-1, 0, 2)
        target = torch.cat([target_xy, torch.zeros_like(target_xy[..., :1])], -1)
        flags = torch.zeros(target.size(0), dtype=torch.int32, device=target.device)
        losses = 0
        for xy_hm, zy_hm, xz_hm in zip(self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps):
            l, _ = K.fused_tail_losses(xy_hm, zy_hm, xz_hm, target, valid_depth=flags,
                                       pixelwise=pixelwise, sigma=1.0)
            losses = losses + l
        return losses

    def forward_3d_losses(self, out_var, target_var):
        """margipose_model.py:236-252: per stage JS(xy) + JS(zy) + JS(xz) + |xyz - target|."""
        pixelwise = self._pixelwise_flag()
        target_xyz = target_var.narrow(-1, 0, 3)
        losses = 0
        for xy_hm, zy_hm, xz_hm in zip(self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps):
            l, _ = K.fused_tail_losses(xy_hm, zy_hm, xz_hm, target_xyz, pixelwise=pixelwise, sigma=1.0)
            losses = losses + l
        return losses

    def forward_mixed_losses(self, out_var, target_var, valid_depth):
        """The mixed branch of bin/train_3d.py:134-140 without its per-sample Python loop: sample b gets the
        3D loss where valid_depth[b] == 1 and the 2D loss (xy plane and x, y only) where it is 0."""
        pixelwise = self._pixelwise_flag()
        target_xyz = target_var.narrow(-1, 0, 3)
        flags = torch.as_tensor(valid_depth).to(device=target_xyz.device, dtype=torch.int32)
        losses = 0
        for xy_hm, zy_hm, xz_hm in zip(self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps):
            l, _ = K.fused_tail_losses(xy_hm, zy_hm, xz_hm, target_xyz, valid_depth=flags,
                                       pixelwise=pixelwise, sigma=1.0)
            losses = losses + l
        return losses

    @staticmethod
    def heatmaps_to_coords(xy_hm, zy_hm, xz_hm):
        """margipose_model.py:254-261 in one launch."""
        return K.heatmaps_to_coords(xy_hm, zy_hm, xz_hm)

    def forward(self, *inputs):
        self.xy_heatmaps, self.zy_heatmaps, self.xz_heatmaps = self._inner_forward(inputs[0])
        xyz = self.heatmaps_to_coords(self.xy_heatmaps[-1], self.zy_heatmaps[-1],
                                      self.xz_heatmaps[-1])
        return xyz


class MargiPoseModelFactory(ModelFactory):
    def __init__(self):
        super().__init__('margipose', '^6.0.0')

    def create(self, model_desc):
        s = model_desc['settings']
        kwargs = dict(
            skel_desc=CanonicalSkeletonDesc,
            n_stages=s.get('n_stages', 4),
            axis_permutation=s.get('axis_permutation', True),
            feature_extractor=s.get('feature_extractor', 'inceptionv4'),
            pixelwise_loss=s.get('pixelwise_loss', 'jsd'),
            precision=s.get('precision', 'bf16'),      # extension of the settings dict: engine arithmetic
        )
        return MargiPoseModel(**kwargs)
