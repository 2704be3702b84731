"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the build container):

    python tests/golden/make_golden.py            # margipose_golden.pt (tail + small models)
    python tests/golden/make_golden.py large      # margipose_golden_large.pt (the benchmarked architectures)

The vectors are produced by /root/reference code (through oracle/ref_shim.py); inputs and
weights are reproducible from seeds with torch's CPU generator, so the GPU box -- which has
no /root/reference -- can regenerate the inputs, and compare both oracle/ and the CUDA path
with what the reference itself computed.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import load_reference  # noqa: E402
from oracle import model_oracle as M  # noqa: E402


def tail_inputs(seed, shape, scale):
    g = torch.Generator().manual_seed(seed)
    b, j, h, w = shape
    z = [torch.randn(b, j, h, w, generator=g) * scale for _ in range(3)]
    target = torch.rand(b, j, 3, generator=g) * 1.6 - 0.8
    mask = (torch.rand(b, j, generator=g) > 0.2).float()
    return z, target, mask


def model_inputs(seed, batch, res=256):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, res, res, generator=g)
    target = torch.rand(batch, 17, 3, generator=g) * 1.6 - 0.8
    mask = torch.ones(batch, 17)
    return x, target, mask


TAIL_CASES = [  # (seed, shape, logit scale)
    (11, (2, 17, 32, 32), 1.0),
    (12, (3, 17, 32, 32), 5.0),
    (13, (2, 17, 16, 16), 3.0),
    (14, (1, 17, 64, 64), 2.0),
    (15, (2, 5, 48, 48), 4.0),
]

MODEL_CASES = [  # (name, weight seed, input seed, batch, desc settings)
    ('r18x1', 21, 22, 1, dict(n_stages=1, feature_extractor='resnet18')),
    ('r18x2', 23, 24, 2, dict(n_stages=2, feature_extractor='resnet18')),
    ('r34x1', 25, 26, 1, dict(n_stages=1, feature_extractor='resnet34')),
]


# The architectures bench.py measures (BASELINE.json configs[1] and configs[4]) at a batch the CPU reference
# finishes in seconds: (name, weight seed, input seed, batch, resolution, desc settings)
LARGE_CASES = [
    ('r34x4', 51, 52, 4, 256, dict(n_stages=4, feature_extractor='resnet34')),
    ('r50x5@384', 53, 54, 2, 384, dict(n_stages=5, feature_extractor='resnet50')),
]


def desc_of(settings):
    s = dict(axis_permutation=True, pixelwise_loss='jsd')
    s.update(settings)
    return {'type': 'margipose', 'version': '6.0.1', 'settings': s}


def main():
    ref = load_reference()
    R = ref.dsntnn
    h2c = ref.model.MargiPoseModel.heatmaps_to_coords
    gold = {}

    # the reference's own known-answer test (tests/test_models.py:39-46)
    xy = R.make_gauss(torch.Tensor([[[-0.5, 0.5]]]), (32, 32), 1, normalize=True)
    zy = R.make_gauss(torch.Tensor([[[0.1, 0]]]), (32, 32), 1, normalize=True)
    xz = R.make_gauss(torch.Tensor([[[0, 0.2]]]), (32, 32), 1, normalize=True)
    gold['known_answer'] = dict(xy=xy, zy=zy, xz=xz, coords=h2c(xy, zy, xz),
                                expected=torch.Tensor([[[-0.5, 0.5, 0.15]]]))

    tails = []
    for seed, shape, scale in TAIL_CASES:
        z, target, mask = tail_inputs(seed, shape, scale)
        z = [t.requires_grad_() for t in z]
        p = [R.flat_softmax(t) for t in z]
        coords = h2c(*p)
        t_xy = target[..., [0, 1]]
        t_zy = target[..., [2, 1]]
        t_xz = target[..., [0, 2]]
        js = [R.js_reg_losses(p[0], t_xy, 1.0), R.js_reg_losses(p[1], t_zy, 1.0),
              R.js_reg_losses(p[2], t_xz, 1.0)]
        eu3 = R.euclidean_losses(coords, target)
        eu2 = R.euclidean_losses(coords[..., :2], t_xy)
        l3 = js[0] + js[1] + js[2] + eu3
        l2 = js[0] + eu2
        loss3 = R.average_loss(l3, mask)
        loss2 = R.average_loss(l2, mask)
        g3 = torch.autograd.grad(loss3, z, retain_graph=True)
        g2 = torch.autograd.grad(loss2, z, retain_graph=True, allow_unused=True)
        small = shape[-1] <= 16
        tails.append(dict(
            seed=seed, shape=shape, scale=scale,
            coords=coords.detach(), dsnt=[R.dsnt(t).detach() for t in p],
            js=[t.detach() for t in js], eu3=eu3.detach(), eu2=eu2.detach(),
            loss3=loss3.detach(), loss2=loss2.detach(),
            # full tensors only for the small case; row/col marginal sums otherwise
            probs=[t.detach() for t in p] if small else None,
            grad3=[t for t in g3] if small else None,
            prob_rowsum=[t.detach().sum(-1) for t in p], prob_colsum=[t.detach().sum(-2) for t in p],
            grad3_rowsum=[t.sum(-1) for t in g3], grad3_absmean=[t.abs().mean() for t in g3],
            grad2_rowsum=[(t.sum(-1) if t is not None else None) for t in g2],
            gauss_rowsum=R.make_gauss(t_xy, shape[-2:], 1.0).sum(-1),
        ))
    gold['tail'] = tails

    gold['model'] = model_cases(ref, [c[:4] + (256,) + c[4:] for c in MODEL_CASES])
    gold['joint_names'] = list(ref.CanonicalSkeletonDesc.joint_names)
    gold['joint_tree'] = list(ref.CanonicalSkeletonDesc.joint_tree)
    gold['hflip_indices'] = list(ref.CanonicalSkeletonDesc.hflip_indices)
    out_path = os.path.join(HERE, 'margipose_golden.pt')
    torch.save(gold, out_path)
    print('wrote', out_path, os.path.getsize(out_path), 'bytes')


def model_cases(ref, cases):
    R = ref.dsntnn
    models = []
    for name, wseed, iseed, batch, res, settings in cases:
        desc = desc_of(settings)
        torch.manual_seed(wseed)
        om = M.create_oracle(desc)          # weights reproducible from the seed
        rm = ref.models.create_model(desc)  # the reference executes them
        rm.load_state_dict(om.state_dict())
        x, target, mask = model_inputs(iseed, batch, res)
        rm.train()
        out = rm(x)
        l3 = R.average_loss(rm.forward_3d_losses(out, target), mask)
        l2 = R.average_loss(rm.forward_2d_losses(out, target), mask)
        l3.backward()
        grads = {k: p.grad for k, p in rm.named_parameters()}
        sd = rm.state_dict()
        probe = [k for k in grads if k.endswith('module.3.weight')][:4] + \
                ['inner.in_cnn.0.weight', 'inner.in_cnn.1.weight', 'inner.in_cnn.1.bias']
        train_hm = dict(
            xy_rowsum=[h.detach().sum(-1) for h in rm.xy_heatmaps],
            zy_rowsum=[h.detach().sum(-1) for h in rm.zy_heatmaps],
            xz_colsum=[h.detach().sum(-2) for h in rm.xz_heatmaps])
        rm.eval()
        with torch.no_grad():
            out_eval = rm(x)
        models.append(dict(
            name=name, desc=desc, weight_seed=wseed, input_seed=iseed, batch=batch, res=res,
            train_coords=out.detach(), loss3=l3.detach(), loss2=l2.detach(),
            **train_hm,
            grad_norms={k: g.norm() for k, g in grads.items()},
            grad_probe={k: grads[k].flatten()[:64].clone() for k in probe},
            running_mean_bn1=sd['inner.in_cnn.1.running_mean'].clone(),
            running_var_bn1=sd['inner.in_cnn.1.running_var'].clone(),
            eval_coords=out_eval,
            n_params=sum(p.numel() for p in rm.parameters()),
            state_keys=list(sd.keys()),
        ))
        print('case', name, 'loss3', l3.item())
    return models


def main_large():
    ref = load_reference()
    gold = {'model': model_cases(ref, LARGE_CASES)}
    out_path = os.path.join(HERE, 'margipose_golden_large.pt')
    torch.save(gold, out_path)
    print('wrote', out_path, os.path.getsize(out_path), 'bytes')


if __name__ == '__main__':
    if 'large' in sys.argv[1:]:
        main_large()
    else:
        main()
