import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shims
    have_ref = ref_shims.available()
    have_gpu = torch.cuda.is_available()
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def mini_state_dict():
    z = load_golden("mini_model.npz")
    sd = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("__")}
    cfg = json.loads(str(z["__config_json__"]))
    targs = json.loads(str(z["__training_args_json__"]))
    return sd, cfg, targs, str(z["__ckpt_name__"])


def write_model_dir(dirname, sd, cfg, targs, ckpt_name="epoch=19-step=1840.ckpt", mean_offset=None):
    """Lay out a training-output directory the way bin/train.py of the reference does."""
    os.makedirs(os.path.join(dirname, "models", "best_by_valid"), exist_ok=True)
    with open(os.path.join(dirname, "config.json"), "w") as f:
        json.dump(cfg, f)
    with open(os.path.join(dirname, "training_args.json"), "w") as f:
        json.dump(targs, f)
    torch.save({"epoch": 19, "global_step": 1840, "state_dict": sd},
               os.path.join(dirname, "models", "best_by_valid", ckpt_name))
    if mean_offset is not None:
        np.save(os.path.join(dirname, "training_mean_offset.npy"), mean_offset)
    return dirname


@pytest.fixture(scope="session")
def mini_dir(tmp_path_factory):
    sd, cfg, targs, ckpt = mini_state_dict()
    return write_model_dir(str(tmp_path_factory.mktemp("mini_model")), sd, cfg, targs, ckpt)


@pytest.fixture(scope="session")
def mini_oracle():
    from oracle import forward as ofwd
    sd, cfg, targs, _ = mini_state_dict()
    ocfg = ofwd.OracleConfig(**cfg)
    return ofwd.OracleModel(sd, ocfg, [True] * 6).eval(), sd, ocfg
