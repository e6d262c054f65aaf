"""Drop-in replacement for the reference's ``models/resnet_custom.py`` (mahmoodlab/TOAD).

``resnet50_baseline(pretrained=False)`` builds the same module tree as the reference
(``ResNet_Baseline(Bottleneck_Baseline, [3, 4, 6, 3])``, truncated after layer3, so the
``state_dict`` keys are torchvision's resnet50 names) but ``forward`` runs the whole trunk in
hand-written sm_100a kernels behind the C ABI (``toad_resnet_fwd``): BatchNorm folded into the
weights, NHWC activations, every convolution a tcgen05 (implicit) GEMM with bias / residual / ReLU
fused in the epilogue.  Only eval mode (running statistics) is implemented -- the reference's use is
offline feature extraction.

Two arithmetic modes (``model.precision``; env ``TOAD_B200_RESNET_EXACT=1`` selects the second as default):
``"f16x2"`` (default) keeps activations as one fp16 plane between layers and the weights as fp16 (hi, lo)
pairs -- activation rounding 2^-11, the input precision of the TF32 convolutions cuDNN runs for the
reference on a GPU; ``"bf16x3"`` keeps (hi, lo) bf16 plane pairs and three tensor passes (fp32-class).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn
import torch.utils.model_zoo as model_zoo

from . import _lib, ops

__all__ = ["ResNet_Baseline", "Bottleneck_Baseline", "resnet50_baseline", "load_pretrained_weights"]

model_urls = {  # resnet_custom.py:11-17
    "resnet18": "https://download.pytorch.org/models/resnet18-5c106cde.pth",
    "resnet34": "https://download.pytorch.org/models/resnet34-333f7ec4.pth",
    "resnet50": "https://download.pytorch.org/models/resnet50-19c8e357.pth",
    "resnet101": "https://download.pytorch.org/models/resnet101-5d3b4d8f.pth",
    "resnet152": "https://download.pytorch.org/models/resnet152-b121ed2d.pth",
}


def _conv_bn(owner: nn.Module, idx: int, cin: int, cout: int, k: int, stride: int) -> None:
    """Register `conv{idx}` (bias-free) and `bn{idx}` on `owner` -- the torchvision attribute names the
    reference's checkpoints use (resnet_custom.py:24-31)."""
    setattr(owner, "conv%d" % idx, nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False))
    setattr(owner, "bn%d" % idx, nn.BatchNorm2d(cout))


class Bottleneck_Baseline(nn.Module):
    """Parameter container for one bottleneck (1x1 -> 3x3/stride -> 1x1 x4, resnet_custom.py:19-34).
    It is never called: the whole trunk runs in toad_resnet_fwd."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        for idx, (cin, cout, k, s) in enumerate(((inplanes, planes, 1, 1), (planes, planes, 3, stride),
                                                 (planes, planes * self.expansion, 1, 1)), start=1):
            _conv_bn(self, idx, cin, cout, k, s)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class ResNet_Baseline(nn.Module):
    """resnet_custom.py:57-109.  `layers[3]` is ignored exactly as in the reference (no layer4 / fc)."""

    STAGES = ((64, 1), (128, 2), (256, 2))   # (planes, stride) of layer1..layer3

    def __init__(self, block, layers):
        super().__init__()
        if list(layers[:3]) != [3, 4, 6] or block is not Bottleneck_Baseline:
            raise NotImplementedError("toad_b200 implements the resnet50_baseline configuration ([3, 4, 6, 3] bottlenecks)")
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for i, ((planes, stride), blocks) in enumerate(zip(self.STAGES, layers), start=1):
            setattr(self, "layer%d" % i, self._make_layer(block, planes, blocks, stride))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.apply(self._init_module)          # kaiming-normal(fan_out) convs, BN weight 1 / bias 0 (:72-77)
        self._prepared = None
        self._prepared_key = None
        self._ws = ops.Workspace()
        self.precision = "bf16x3" if os.environ.get("TOAD_B200_RESNET_EXACT", "0") == "1" else "f16x2"

    def _flags(self) -> int:
        if self.precision not in ("f16x2", "bf16x3"):
            raise ValueError("precision must be 'f16x2' or 'bf16x3', got %r" % (self.precision,))
        return _lib.RESNET_FLAG_EXACT if self.precision == "bf16x3" else 0

    @staticmethod
    def _init_module(m: nn.Module) -> None:
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    def _make_layer(self, block, planes, blocks, stride=1):
        out_ch = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != out_ch:   # projection shortcut on the first block (:82-87)
            shortcut = nn.Sequential(nn.Conv2d(self.inplanes, out_ch, kernel_size=1, stride=stride, bias=False),
                                     nn.BatchNorm2d(out_ch))
        seq = [block(self.inplanes, planes, stride, shortcut)]
        self.inplanes = out_ch
        seq += [block(out_ch, planes) for _ in range(blocks - 1)]
        return nn.Sequential(*seq)

    # -- tensors in state_dict order without num_batches_tracked (the C ABI's `tensors` array)
    def _tensor_list(self):
        return [v for k, v in self.state_dict(keep_vars=True).items() if not k.endswith("num_batches_tracked")]

    def _prepare(self, device):
        lib = _lib.load()
        tensors = self._tensor_list()
        flags = self._flags()
        key = tuple((t.data_ptr(), t._version) for t in tensors) + (flags,)
        if self._prepared is not None and key == self._prepared_key and self._prepared.device == device:
            return
        for t in tensors:
            ops._check_dev_f32(t, "resnet parameter")
        nbytes = C.c_size_t()
        _lib.check(lib.toad_resnet_prepared_bytes(C.byref(nbytes)), "toad_resnet_prepared_bytes")
        buf = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device)
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        ptr = (buf.data_ptr() + 255) // 256 * 256
        _lib.check(lib.toad_resnet_prepare(arr, len(tensors), ptr, nbytes.value, flags, ops._stream()), "toad_resnet_prepare")
        self._prepared, self._prepared_key, self._prepared_ptr, self._prepared_bytes = buf, key, ptr, nbytes.value

    def invalidate_weight_cache(self) -> None:
        """Forget the folded-BN weight planes (keyed on each tensor's (data_ptr, autograd version)): call after writes
        that bypass the version counter (`p.data.copy_()`, raw-pointer writes); load_state_dict is tracked."""
        self._prepared = None
        self._prepared_key = None

    def forward(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """[B, 3, H, W] -> [B, 1024] (resnet_custom.py:96-109).  `out` (optional, beyond the reference's signature):
        a [B, 1024] fp32 CUDA buffer to write into, e.g. a slice of a slide's feature matrix."""
        if self.training:
            raise NotImplementedError("toad_b200: resnet50_baseline runs in eval mode only (BatchNorm running "
                                      "statistics are folded into the convolutions); call .eval() first")
        ops._check_dev_f32(x, "x")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("x must be [B, 3, H, W], got %s" % (tuple(x.shape),))
        lib = _lib.load()
        self._prepare(x.device)
        B, _, H, W = x.shape
        if out is None:
            out = torch.empty((B, 1024), dtype=torch.float32, device=x.device)
        else:
            ops._check_dev_f32(out, "out", (B, 1024))
        nbytes = C.c_size_t()
        flags = self._flags()
        _lib.check(lib.toad_resnet_workspace_bytes(B, H, W, flags, C.byref(nbytes)), "toad_resnet_workspace_bytes")
        wptr, wsize = self._ws.get(nbytes.value, x.device)
        _lib.check(lib.toad_resnet_fwd(self._prepared_ptr, x.data_ptr(), B, H, W, out.data_ptr(), wptr, wsize,
                                       flags, ops._stream()), "toad_resnet_fwd")
        return out


def resnet50_baseline(pretrained: bool = False) -> ResNet_Baseline:
    """Modified ResNet-50 truncated after layer3 (resnet_custom.py:111-119)."""
    model = ResNet_Baseline(Bottleneck_Baseline, [3, 4, 6, 3])
    if pretrained:
        model = load_pretrained_weights(model, "resnet50")
    return model


def load_pretrained_weights(model, name):
    """resnet_custom.py:121-124 (needs network access, like the reference)."""
    pretrained_dict = model_zoo.load_url(model_urls[name])
    model.load_state_dict(pretrained_dict, strict=False)
    return model
