"""The reference's own callers resolve to the drop-in: with this repo ahead of /root/reference on
sys.path, `utils/core_utils_mtl_concat.py:8` and `utils/eval_utils_mtl_concat.py:6,15` import THIS repo's classes,
construct them the way the reference does (`TOAD_fc_mtl_concat(**model_dict)`, core_utils:113-116), print them
(`print_network`, utils.py:72-84) and build the optimizer over their parameters (`get_optim`, utils.py:63-70).

Runs in a subprocess (the import-path experiment must not leak into this pytest process); skipped where
/root/reference is absent (the GPU box).  The orchestration layers need stubs for packages this image lacks
(SURVEY.md F9): tensorboardX, h5py, openslide-free `datasets` namespace -- stubbed in the child, never in the reference.
"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

CHILD = textwrap.dedent('''
    import sys, types, importlib.machinery
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    stub("tensorboardX", SummaryWriter=object)
    stub("h5py", File=object)
    # the reference's `datasets/` directory is shadowed by the HuggingFace `datasets` package of this image
    ds = stub("datasets"); ds.__path__ = ["%(ref)s/datasets"]
    stub("datasets.dataset_mtl_concat", save_splits=lambda *a, **k: None)
    import utils.core_utils_mtl_concat as cu
    import utils.eval_utils_mtl_concat as eu
    import toad_b200.model_toad as ours
    import toad_b200.resnet_custom as ours_r
    assert cu.__file__.startswith("%(ref)s"), cu.__file__
    assert cu.TOAD_fc_mtl_concat is ours.TOAD_fc_mtl_concat, cu.TOAD_fc_mtl_concat.__module__
    assert eu.TOAD_fc_mtl_concat is ours.TOAD_fc_mtl_concat
    assert eu.resnet50_baseline is ours_r.resnet50_baseline
    # construction + print_network + get_optim, as train() / initiate_model() do (core_utils:113-124, eval_utils:19-27)
    args = types.SimpleNamespace(drop_out=True, n_classes=18, opt="adam", lr=1e-4, reg=1e-5)
    model = cu.TOAD_fc_mtl_concat(**{"dropout": args.drop_out, "n_classes": args.n_classes})
    cu.print_network(model)
    opt = cu.get_optim(model, args)
    n = sum(p.numel() for g in opt.param_groups for p in g["params"])
    assert n == 1192490, n
    sd_keys = list(model.state_dict().keys())
    assert sd_keys[0] == "attention_net.0.weight" and "attention_net.6.attention_c.bias" in sd_keys, sd_keys
    print("CALLERS-OK")
''')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils")), reason="/root/reference is not present on this machine")
def test_reference_callers_import_the_drop_in():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + REF, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", CHILD % {"ref": REF}], cwd="/tmp", env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "CALLERS-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_dataparallel_checkpoint_keys_load_both_ways():
    """A multi-GPU reference process saves `attention_net.module.*` keys (model_toad.py:79-82 wraps the trunk in
    nn.DataParallel) and loads with strict=False (eval_utils_mtl_concat.py:28-30): both key forms must load ALL tensors,
    and a partial load must not be silent."""
    import warnings

    import torch
    from models.model_toad import TOAD_fc_mtl_concat
    torch.manual_seed(0)
    src = TOAD_fc_mtl_concat(dropout=True, n_classes=18)
    dp = src.state_dict_dataparallel()
    assert sum(k.startswith("attention_net.module.") for k in dp) == 10 and len(dp) == 14
    for strict in (False, True):
        dst = TOAD_fc_mtl_concat(dropout=True, n_classes=18)
        res = dst.load_state_dict({k: v.clone() for k, v in dp.items()}, strict=strict)
        assert not res.missing_keys and not res.unexpected_keys
        for (k, a), b in zip(src.state_dict().items(), dst.state_dict().values()):
            assert torch.equal(a, b), k
    # what a reference nn.DataParallel module would accept: its own key set
    ref_like = torch.nn.ModuleDict({"attention_net": torch.nn.DataParallel(src.attention_net), "classifier": src.classifier,
                                    "site_classifier": src.site_classifier})
    assert sorted(ref_like.state_dict().keys()) == sorted(dp.keys())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        TOAD_fc_mtl_concat(n_classes=18).load_state_dict({"classifier.bias": torch.zeros(18)}, strict=False)
    assert any("untouched" in str(x.message) for x in w)
