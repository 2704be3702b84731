"""Model registry: the drop-in boundary of the reference
(/root/reference/src/margipose/models/__init__.py:10-34) -- same `create_model(model_desc)` /
`load_model(file)` / `MODEL_FACTORIES` surface, backed by the sm_100a engine.

Only the MargiPose model is on the hot path; a 'chatterbox' descriptor is
rejected with the registry's own 'unrecognised model' error.
"""
import torch

from ..model_factory import Version
from .margipose_model import MargiPoseModelFactory

MODEL_FACTORIES = [
    MargiPoseModelFactory(),
]


def create_model(model_desc):
    type_name = model_desc['type']
    version = Version(model_desc['version'])

    for factory in MODEL_FACTORIES:
        if factory.is_for(type_name, version):
            model = factory.create(model_desc)
            break
    else:
        raise Exception('unrecognised model {} v{}'.format(type_name, str(version)))

    return model


def load_model(model_file, allow_pickle=False):
    """models/__init__.py:30-34.  A reference checkpoint is a dict of plain containers and tensors
    ({'state_dict', 'model_desc', 'train_datasets', 'optimizer', 'epoch'}, bin/train_3d.py:374-382), so
    it loads under `weights_only=True`; full unpickling (arbitrary code execution) is an explicit opt-in."""
    details = torch.load(model_file, map_location='cpu', weights_only=not allow_pickle)
    model = create_model(details['model_desc'])
    model.load_state_dict(details['state_dict'])
    return model
