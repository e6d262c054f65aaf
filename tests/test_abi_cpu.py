"""The C-ABI library loads without a GPU, exports every symbol include/toad_b200.h declares, and
its pure-host entry points (sizes, offsets, argument validation) behave.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from toad_b200 import _lib, build
    build.build()                     # nvcc cross-compiles for sm_100a without a GPU
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "toad_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(toad_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from toad_b200 import _lib
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names              # the binding covers exactly the header


def test_abi_version_and_error_strings(lib):
    from toad_b200 import _lib as L
    assert lib.toad_abi_version() == L.ABI_VERSION == 3
    assert lib.toad_error_string(0) == b"ok"
    assert b"workspace" in lib.toad_error_string(-2)
    assert b"not supported" in lib.toad_error_string(-3)


def test_param_offsets_match_state_dict_layout(lib):
    from toad_b200 import ops
    from models.model_toad import TOAD_fc_mtl_concat
    for size_arg, n_classes in (("big", 18), ("small", 2)):
        m = TOAD_fc_mtl_concat(size_arg=size_arg, n_classes=n_classes)
        off = ops.param_offsets(m._dims)
        sizes = [p.numel() for p in m._param_list()]
        assert [off[i + 1] - off[i] for i in range(14)] == sizes
        assert off[14] == sum(p.numel() for p in m.parameters())
        # parameter order == state_dict order (the checkpoint-compatibility contract)
        assert [p.data_ptr() for p in m._param_list()] == [v.data_ptr() for v in m.state_dict().values()]
    assert ops.param_offsets(TOAD_fc_mtl_concat(n_classes=18)._dims)[14] == 1192490


def test_state_dict_keys_are_the_references():
    from models.model_toad import TOAD_fc_mtl_concat
    from oracle.toad_oracle import PARAM_KEYS
    assert list(TOAD_fc_mtl_concat(n_classes=18).state_dict().keys()) == PARAM_KEYS
    keys_do = list(TOAD_fc_mtl_concat(dropout=True).state_dict().keys())
    assert "attention_net.3.weight" in keys_do and "attention_net.6.attention_c.bias" in keys_do   # indices shift, as in the reference
    with pytest.raises(NameError):
        TOAD_fc_mtl_concat(gate=False)                # the reference's own behaviour (model_toad.py:68)


def test_workspace_queries_and_argument_errors(lib):
    from toad_b200._lib import Dims
    d = Dims(1024, 512, 384, 2, 18)
    n = C.c_size_t()
    assert lib.toad_fwd_workspace_bytes(C.byref(d), 50000, 0, C.byref(n)) == 0
    tc_bytes = n.value
    assert lib.toad_fwd_workspace_bytes(C.byref(d), 50000, 2, C.byref(n)) == 0
    assert 0 < tc_bytes < (1 << 30) and n.value > 0
    assert lib.toad_bwd_workspace_bytes(C.byref(d), 50000, 0, C.byref(n)) == 0 and n.value > 0
    assert lib.toad_bwd_workspace_bytes(C.byref(d), 50000, 2, C.byref(n)) == 0 and n.value > 0
    assert lib.toad_fwd_workspace_bytes(C.byref(d), 0, 0, C.byref(n)) == -1           # empty bag
    assert lib.toad_fwd_workspace_bytes(None, 10, 0, C.byref(n)) == -1               # null dims
    bad = Dims(1000, 512, 384, 2, 18)                                                # in_dim % 64 != 0
    assert lib.toad_fwd_workspace_bytes(C.byref(bad), 10, 0, C.byref(n)) == -3
    bad = Dims(1024, 256, 384, 2, 18)
    assert lib.toad_fwd_workspace_bytes(C.byref(bad), 10, 0, C.byref(n)) == -3
    assert lib.toad_fwd(C.byref(d), None, None, 10, None, None, None, None, 0, 0, None) == -1
    assert lib.toad_topk(None, 10, 1, None, None, None, 0, None) == -1
    assert lib.toad_linear_workspace_bytes(128, 64, 64, C.byref(n)) == 0 and n.value > 0
    assert lib.toad_attn_gated_workspace_bytes(1024, 256, 1, 256, 0, C.byref(n)) == 0
    assert lib.toad_attn_gated_workspace_bytes(1000, 256, 1, 256, 0, C.byref(n)) == -3
    # batched forward: 1..16 slides per call, scratch grows with the slide count, offsets are validated before any launch
    assert lib.toad_fwd_batch_workspace_bytes(C.byref(d), 50000, 1, 0, C.byref(n)) == 0 and n.value == tc_bytes
    one = n.value
    assert lib.toad_fwd_batch_workspace_bytes(C.byref(d), 50000, 16, 0, C.byref(n)) == 0 and n.value > one
    assert lib.toad_fwd_batch_workspace_bytes(C.byref(d), 50000, 17, 0, C.byref(n)) == -1
    assert lib.toad_fwd_batch_workspace_bytes(C.byref(d), 50000, 0, 0, C.byref(n)) == -1
    offs = (C.c_int64 * 3)(0, 10, 10)                                                  # second slide is empty
    assert lib.toad_fwd_batch(C.byref(d), None, None, offs, 2, None, None, None, 0, 0, None) == -1
    offs = (C.c_int64 * 3)(5, 10, 20)                                                  # offsets[0] != 0
    assert lib.toad_fwd_batch(C.byref(d), None, None, offs, 2, None, None, None, 0, 0, None) == -1
    assert lib.toad_topk_workspace_bytes(200000, 1000, C.byref(n)) == 0 and 0 < n.value < (4 << 20)
    assert lib.toad_gather_rows(None, 10, 8, None, 1, None, None) == -1


def test_module_refuses_cpu_tensors():
    import torch
    from models.model_toad import TOAD_fc_mtl_concat
    m = TOAD_fc_mtl_concat(n_classes=18).eval()
    with pytest.raises(ValueError):
        m(torch.randn(4, 1024), torch.tensor([1.0]))       # no CPU fallback
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m.relocate()


def test_dropout_hash_matches_oracle(lib):
    """The mask hash is host-callable: pin the numpy restatement used by the GPU dropout tests."""
    import numpy as np
    from oracle import toad_oracle as O
    idx = np.array([0, 1, 2, 511, 512, 12345, 2 ** 33 + 5, 2 ** 40 + 7], dtype=np.uint64)
    for seed, layer in ((0, 1), (987654321, 3), (2 ** 62 - 1, 4)):
        ours = [lib.toad_dropout_hash(seed, layer, int(i)) for i in idx]
        assert O.dropout_hash(seed, layer, idx).tolist() == ours
    m = O.dropout_multipliers(42, 0.25, 64, 512, 256)
    assert abs((m[1] == 0).mean() - 0.25) < 0.02 and set(np.unique(m[3])) == {0.0, 1.0 / 0.75}


def test_python_flag_constants_match_the_header():
    """toad_b200/_lib.py restates the header's flag bits and ABI version: a drifted constant would silently select another
    code path (or none)."""
    import re
    from toad_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "toad_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(TOAD_\w+)\s+(\d+)u?\b", hdr)}
    assert defs["TOAD_ABI_VERSION"] == _lib.ABI_VERSION
    seen = 0
    for name, value in vars(_lib).items():
        if name.startswith("FLAG_"):
            assert defs["TOAD_" + name] == value, name
            seen += 1
    assert seen >= 8
    assert defs["TOAD_RESNET_FLAG_EXACT"] == _lib.RESNET_FLAG_EXACT
    fwd_flags = [v for k, v in defs.items() if k.startswith("TOAD_FLAG_")]
    assert len(set(fwd_flags)) == len(fwd_flags) and all(v & (v - 1) == 0 for v in fwd_flags)   # distinct single bits

