import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


# ---- observed-error log: tests call `parity_log(test, metric=value, ...)`; the session writes everything it saw to
# gpurun_out/parity_observed.json (MP_PARITY_OUT overrides), which is what profiles/r02_parity_observed.json and
# PARITY.md are made from -- the tolerances in the tests say what is allowed, this file says what was measured.
_OBSERVED = {}


def parity_log(test, **metrics):
    entry = _OBSERVED.setdefault(test, {})
    for k, v in metrics.items():
        entry[k] = float(v) if isinstance(v, (int, float)) or hasattr(v, 'item') else v


def pytest_sessionfinish(session, exitstatus):
    if not _OBSERVED:
        return
    import json
    path = os.environ.get('MP_PARITY_OUT', os.path.join(ROOT, 'gpurun_out', 'parity_observed.json'))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        old = {}
        if os.path.exists(path):
            with open(path) as f:
                old = json.load(f)
        old.update(_OBSERVED)
        with open(path, 'w') as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except OSError:
        pass
