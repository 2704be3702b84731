"""CPU checks of the host side: registry / factory behaviour (mirror of
/root/reference/src/margipose/models/__init__.py:16-27 and model_factory.py), state_dict surface,
joint bookkeeping, the C ABI exports, and that the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from oracle import model_oracle as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = torch.load(os.path.join(ROOT, 'tests', 'golden', 'margipose_golden.pt'), weights_only=False)


def desc(**settings):
    s = dict(n_stages=1, feature_extractor='resnet18', axis_permutation=True, pixelwise_loss='jsd')
    s.update(settings)
    return {'type': 'margipose', 'version': '6.0.1', 'settings': s}


def test_registry_matches_reference_behaviour():
    from margipose_b200.models import create_model, MODEL_FACTORIES
    from margipose_b200.model_factory import Version, Spec
    assert len(MODEL_FACTORIES) == 1 and MODEL_FACTORIES[0].is_for('margipose', Version('6.0.1'))
    assert Version('6.3.2') in Spec('^6.0.0') and Version('7.0.0') not in Spec('^6.0.0')
    assert Version('5.9.9') not in Spec('^6.0.0') and Version('6.0.0') in Spec('^6.0.0')
    with pytest.raises(Exception, match='unrecognised model chatterbox v1.0.0'):
        create_model({'type': 'chatterbox', 'version': '1.0.0', 'settings': {}})
    with pytest.raises(Exception, match='unrecognised model margipose v7.0.0'):
        create_model({'type': 'margipose', 'version': '7.0.0', 'settings': {}})
    with pytest.raises(Exception, match='unsupported image feature extractor'):
        create_model(desc(feature_extractor='vgg16'))
    with pytest.raises(Exception, match='unsupported image feature extractor'):
        create_model({'type': 'margipose', 'version': '6.0.1', 'settings': {}})    # default = inceptionv4


@pytest.mark.parametrize('case', GOLD['model'], ids=lambda c: c['name'])
def test_state_dict_surface_is_the_references(case):
    from margipose_b200.models import create_model
    model = create_model(case['desc'])
    sd = model.state_dict()
    assert list(sd.keys()) == case['state_keys']
    assert sum(p.numel() for p in model.parameters()) == case['n_params']
    torch.manual_seed(0)
    om = M.create_oracle(case['desc'])
    model.load_state_dict(om.state_dict())          # shapes agree
    om.load_state_dict(model.state_dict())
    assert model.data_specs.input_specs.height == 256 and model.data_specs.output_specs.n_dims == 3
    assert model.xy_heatmaps is None


def test_param_count_symmetry_of_columns():
    # /root/reference/tests/test_models.py:11-16
    from margipose_b200.models.margipose_model import HeatmapColumn
    n = [sum(p.numel() for p in HeatmapColumn(17, s).parameters()) for s in ('xy', 'zy', 'xz')]
    assert n[0] == n[1] == n[2] == 4739599


def test_joint_bookkeeping_is_bit_exact():
    from margipose_b200.skeleton import CanonicalSkeletonDesc as S
    assert S.joint_names == GOLD['joint_names'] == M.JOINT_NAMES
    assert S.joint_tree == GOLD['joint_tree'] and S.hflip_indices == GOLD['hflip_indices']
    assert S.n_joints == 17 and S.root_joint_id == 14


def test_flat_bank_layout_roundtrip_on_cpu():
    """Parameters become views of one flat buffer in channels-last (GEMM) order without changing
    what state_dict() / load_state_dict() see."""
    from margipose_b200.models import create_model
    from margipose_b200.engine import ParamBank, build_layers
    torch.manual_seed(3)
    model = create_model(desc(n_stages=2))
    before = {k: v.clone() for k, v in model.state_dict().items()}
    bank = ParamBank()
    layers = build_layers(model, bank)
    bank.finalize(torch.device('cpu'))
    after = model.state_dict()
    for k in before:
        assert torch.equal(before[k].float(), after[k].float()), k
    w = model.inner.xy_hm_cnns[0].down_layers[0].module[0].weight
    assert w.shape == (128, 128, 3, 3) and w.is_contiguous(memory_format=torch.channels_last)
    assert bank.linked()
    assert layers.n_bn == sum(isinstance(m, torch.nn.BatchNorm2d) for m in model.modules())
    # every pack-table entry stays inside the bf16 pack buffer and work ranges tile [0, n_work)
    pos = 0
    for e in bank.pack_entries:
        assert e['work_off'] == pos
        pos = e['work_end']
        last = e['dst_off'] + (e['rows_p'] - 1) * e['row_stride'] + e['taps'] * e['cols_p']
        assert last <= bank.n_pack
    assert pos == bank.n_work
    with torch.no_grad():
        w.mul_(2.0)                                  # writes through to the flat buffer
    assert torch.equal(model.state_dict()['inner.xy_hm_cnns.0.down_layers.0.module.0.weight'], before[
        'inner.xy_hm_cnns.0.down_layers.0.module.0.weight'] * 2)


def test_c_abi_exports_every_declared_symbol():
    from margipose_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'margipose_b200.h')).read()
    declared = set(re.findall(r'MP_API\s+[\w\s\*]+?\b(mp_\w+)\s*\(', header))
    assert declared, 'no declarations parsed'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert declared == set(_lib.exported_symbols())
    assert _lib.lib().mp_abi_version() >= 1


def test_no_cpu_fallback():
    from margipose_b200.models import create_model
    from margipose_b200 import dsntnn as K
    from margipose_b200._lib import MargiposeB200Error
    with pytest.raises(MargiposeB200Error):
        create_model(desc())(torch.randn(1, 3, 256, 256))
    with pytest.raises(MargiposeB200Error):
        K.flat_softmax(torch.randn(1, 17, 32, 32))
    with pytest.raises(MargiposeB200Error):
        K.dsnt(torch.rand(1, 17, 32, 32))


def test_engine_zips_the_three_column_programs_into_grouped_launches():
    """engine.Engine._merge_lanes (host logic, no GPU): the xy / zy / xz columns record structurally identical
    launch programs (models/margipose_model.py:196-198); ops with a grouped C-ABI variant become ONE launch over
    a contiguous array of the three argument structs, ops only some columns have (the axis permutations of the
    zy / xz columns) keep their own launch, and per-column order is preserved."""
    from margipose_b200 import engine as E
    from margipose_b200._lib import IgemmArgs, BnArgs

    def op(name, args, flops=0.0, **flags):
        f = lambda: None
        f.name, f.args, f.flops = name, args, flops
        for k, v in flags.items():
            setattr(f, k, v)
        return f

    def lane(k, permute):
        a, b, c = IgemmArgs(), BnArgs(), IgemmArgs()
        a.n_img, c.n_img, b.C = 10 + k, 20 + k, 30 + k
        ops = [op('mp_conv_igemm', a, flops=1.0), op('mp_bn_fwd', b, join_aux=True)]
        if permute:
            ops.append(op('mp_axis_permute', None))
        ops.append(op('mp_conv_igemm', c, flops=2.0, aux=True))
        return ops

    eng = E.Engine.__new__(E.Engine)
    eng.device = torch.device('cpu')
    lanes = [lane(0, False), lane(1, True), lane(2, True)]
    merged = eng._merge_lanes(lanes)
    assert [m.name for m in merged] == ['mp_conv_igemm', 'mp_bn_fwd', 'mp_axis_permute', 'mp_axis_permute',
                                        'mp_conv_igemm']
    first, bn, p1, p2, last = merged
    assert len(first.parts) == 3 and [first.args[i].n_img for i in range(3)] == [10, 11, 12]
    assert [bn.args[i].C for i in range(3)] == [30, 31, 32] and bn.join_aux
    assert p1 is lanes[1][2] and p2 is lanes[2][2]
    assert [last.args[i].n_img for i in range(3)] == [20, 21, 22] and last.aux and last.flops == 6.0
    assert first.flops == 3.0 and not getattr(first, 'aux', False)
    # the grouped argument array is contiguous memory of the C struct (what mp_*_grouped expects)
    assert ctypes.sizeof(first.args) == 3 * ctypes.sizeof(IgemmArgs)


def test_reference_checkpoint_loads_on_cpu_without_unpickling_code(tmp_path):
    """A checkpoint in the reference's wire format (bin/train_3d.py:374-382), written from the oracle -- whose module
    tree has the reference's key names -- with a torch.optim.SGD state as the reference saves it, loads through
    `load_model` (models/__init__.py:30-34) under `weights_only=True`: every tensor arrives under the same key.
    A file that needs arbitrary unpickling is refused unless the caller opts in."""
    import pickle
    from margipose_b200.models import load_model
    d = desc(n_stages=2)
    torch.manual_seed(3)
    om = M.create_oracle(d)
    opt = torch.optim.SGD(om.parameters(), lr=0.1, momentum=0.9)
    path = str(tmp_path / 'model-latest.pth')
    torch.save({'state_dict': om.state_dict(), 'model_desc': d, 'train_datasets': ['mpi3d-train', 'mpii-train'],
                'optimizer': opt.state_dict(), 'epoch': 7}, path)
    model = load_model(path)
    want, got = om.state_dict(), model.state_dict()
    assert list(want.keys()) == list(got.keys())
    for k in want:
        assert torch.equal(want[k], got[k]), k
    assert model.inner.n_stages == 2 and not any(p.is_cuda for p in model.parameters())

    class Evil:
        def __reduce__(self):
            return (print, ('code ran during unpickling',))
    bad = str(tmp_path / 'bad.pth')
    torch.save({'state_dict': om.state_dict(), 'model_desc': d, 'extra': Evil()}, bad)
    with pytest.raises(pickle.UnpicklingError):
        load_model(bad)


@pytest.mark.skipif(not os.path.exists('/root/reference/src/margipose/data_specs.py'), reason='reference tree not present')
def test_image_specs_convert_is_the_references_bit_for_bit():
    """bin/infer_single.py:62-66 goes through `model.data_specs.input_specs.convert / unconvert`
    (data_specs.py:6-42): same tensors, same image back; `convert_uint8` + the CUDA path's arithmetic
    ((x / 255 - mean) / stddev, csrc/misc.cu stem gather) is the same map."""
    import importlib.util
    import numpy as np
    import PIL.Image
    from margipose_b200.data_specs import ImageSpecs
    spec = importlib.util.spec_from_file_location('_ref_data_specs', '/root/reference/src/margipose/data_specs.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(5)
    img = PIL.Image.fromarray(rng.integers(0, 256, (48, 64, 3), dtype=np.uint8), 'RGB')
    for mean, std in [(ImageSpecs.IMAGENET_MEAN, ImageSpecs.IMAGENET_STDDEV), (None, None)]:
        ours, theirs = ImageSpecs(256, mean, std), ref.ImageSpecs(256, mean, std)
        a, b = ours.convert(img), theirs.convert(img)
        assert a.dtype == torch.float32 and a.shape == (3, 48, 64) and torch.equal(a, b)
        assert np.array_equal(np.array(ours.unconvert(a)), np.array(theirs.unconvert(b)))
        assert torch.equal(a, b), 'unconvert must not modify its argument'
    u8 = ImageSpecs(256, ImageSpecs.IMAGENET_MEAN, ImageSpecs.IMAGENET_STDDEV).convert_uint8(img)
    assert u8.dtype == torch.uint8 and u8.shape == (48, 64, 3)
    mean, std = torch.tensor(ImageSpecs.IMAGENET_MEAN), torch.tensor(ImageSpecs.IMAGENET_STDDEV)
    fused = ((u8.float() / 255 - mean) / std).permute(2, 0, 1)
    host = ImageSpecs(256, ImageSpecs.IMAGENET_MEAN, ImageSpecs.IMAGENET_STDDEV).convert(img)
    torch.testing.assert_close(fused, host, rtol=0, atol=1e-6)
    # greyscale input is promoted to RGB, as torchvision's to_tensor would not do silently: the model needs 3 channels
    grey = PIL.Image.fromarray(rng.integers(0, 256, (8, 8), dtype=np.uint8), 'L')
    assert ImageSpecs(256).convert(grey).shape == (3, 8, 8)
