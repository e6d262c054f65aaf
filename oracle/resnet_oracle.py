"""CPU oracle for the truncated ResNet-50 feature extractor -- TEST / BASELINE INFRASTRUCTURE ONLY.

Functional restatement of the reference's ``models/resnet_custom.py`` (mahmoodlab/TOAD) in eval
mode over a plain ``state_dict``: Bottleneck_Baseline.forward (:35-55), ResNet_Baseline.__init__
/_make_layer (:57-94, layers [3,4,6] -- layer4/fc are never built) and ResNet_Baseline.forward
(:96-109).  Same library calls as the reference (F.conv2d, batch_norm with running stats, relu,
max_pool2d, adaptive_avg_pool2d), so it is the reference's own CPU arithmetic without the module
tree.  Pinned against reference-generated golden vectors (tests/golden/resnet_*.npz, made by
tests/golden/make_golden_resnet.py); nothing in toad_b200/ or models/ imports it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

LAYERS = [("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 6, 2)]   # resnet_custom.py:67-69
EPS = 1e-5


def param_spec():
    """(key, shape) of every tensor of resnet50_baseline's state_dict, in order (torchvision names)."""
    spec = [("conv1.weight", (64, 3, 7, 7))]
    spec += _bn("bn1", 64)
    inplanes = 64
    for name, planes, blocks, stride in LAYERS:
        for i in range(blocks):
            p = "%s.%d" % (name, i)
            s = stride if i == 0 else 1
            spec.append((p + ".conv1.weight", (planes, inplanes, 1, 1)))
            spec += _bn(p + ".bn1", planes)
            spec.append((p + ".conv2.weight", (planes, planes, 3, 3)))
            spec += _bn(p + ".bn2", planes)
            spec.append((p + ".conv3.weight", (planes * 4, planes, 1, 1)))
            spec += _bn(p + ".bn3", planes * 4)
            if i == 0 and (s != 1 or inplanes != planes * 4):
                spec.append((p + ".downsample.0.weight", (planes * 4, inplanes, 1, 1)))
                spec += _bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    return spec


def _bn(prefix, c):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)), (prefix + ".running_mean", (c,)),
            (prefix + ".running_var", (c,)), (prefix + ".num_batches_tracked", ())]


def make_params(seed: int) -> dict:
    """Seeded synthetic weights: kaiming-normal(fan_out) convs (resnet_custom.py:72-74) and RANDOMISED
    BatchNorm affine/running stats (so that BN folding is really exercised)."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in param_spec():
        if k.endswith("num_batches_tracked"):
            out[k] = np.array(0, dtype=np.int64)
        elif len(shp) == 4:
            std = np.sqrt(2.0 / (shp[0] * shp[2] * shp[3]))
            out[k] = (rng.standard_normal(shp, dtype=np.float32) * np.float32(std)).astype(np.float32)
        elif k.endswith("running_var"):
            out[k] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        elif k.endswith("running_mean") or k.endswith(".bias"):
            out[k] = (rng.standard_normal(shp, dtype=np.float32) * np.float32(0.1)).astype(np.float32)
        else:  # bn weight
            out[k] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
    return out


def make_images(seed: int, batch: int, size: int = 256, width: int = None) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((batch, 3, size, width or size), dtype=np.float32)


def _bn_eval(x, p, prefix):
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"],
                        p[prefix + ".bias"], training=False, eps=EPS)


@torch.no_grad()
def resnet50_baseline_forward(x: torch.Tensor, params: dict) -> torch.Tensor:
    """x [B,3,H,W] -> [B,1024] (resnet_custom.py:96-109), eval mode."""
    p = {k: (torch.from_numpy(np.asarray(v)) if not isinstance(v, torch.Tensor) else v).to(x.dtype)
         if not k.endswith("num_batches_tracked") else v for k, v in params.items()}
    x = F.conv2d(x, p["conv1.weight"], stride=2, padding=3)                 # :97
    x = F.relu(_bn_eval(x, p, "bn1"))                                       # :98-99
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)                 # :100
    for name, planes, blocks, stride in LAYERS:                             # :102-104
        for i in range(blocks):
            pre = "%s.%d" % (name, i)
            s = stride if i == 0 else 1
            residual = x
            out = F.relu(_bn_eval(F.conv2d(x, p[pre + ".conv1.weight"]), p, pre + ".bn1"))            # :38-40
            out = F.relu(_bn_eval(F.conv2d(out, p[pre + ".conv2.weight"], stride=s, padding=1), p, pre + ".bn2"))  # :42-44
            out = _bn_eval(F.conv2d(out, p[pre + ".conv3.weight"]), p, pre + ".bn3")                  # :46-47
            if (pre + ".downsample.0.weight") in p:
                residual = _bn_eval(F.conv2d(x, p[pre + ".downsample.0.weight"], stride=s), p, pre + ".downsample.1")  # :49-50
            x = F.relu(out + residual)                                      # :52-53
    x = F.adaptive_avg_pool2d(x, 1)                                         # :106
    return x.view(x.size(0), -1)                                            # :107


@torch.no_grad()
def resnet50_baseline_forward_f16act(x: torch.Tensor, params: dict) -> torch.Tensor:
    """Restatement of the CUDA trunk's default arithmetic ("f16x2" mode, include/toad_b200.h): the same network with
    BatchNorm folded into the convolution weights (hi + lo fp16 pairs ~ fp32 weights), fp32 accumulation, and every
    activation a layer STORES (stem output, each conv's output, each block's output) rounded to fp16.  Used by the CPU
    tests to show that this mode meets the 1e-3 parity bar against the reference's goldens; the product never calls it."""
    def q(t):
        return t.to(torch.float16).to(t.dtype)

    p = {k: (torch.from_numpy(np.asarray(v)) if not isinstance(v, torch.Tensor) else v).to(x.dtype)
         if not k.endswith("num_batches_tracked") else v for k, v in params.items()}

    def conv_bn(t, conv, bn, **kw):
        scale = p[bn + ".weight"] / torch.sqrt(p[bn + ".running_var"] + EPS)
        w = p[conv + ".weight"] * scale.view(-1, 1, 1, 1)
        w_hi = q(w)
        w = w_hi + q(w - w_hi)                                   # what the (hi, lo) fp16 weight planes hold
        return F.conv2d(t, w, **kw) + (p[bn + ".bias"] - p[bn + ".running_mean"] * scale).view(1, -1, 1, 1)

    x = q(x)                                                     # the stem's im2col plane is fp16
    x = q(F.relu(conv_bn(x, "conv1", "bn1", stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for name, planes, blocks, stride in LAYERS:
        for i in range(blocks):
            pre = "%s.%d" % (name, i)
            s = stride if i == 0 else 1
            residual = x
            out = q(F.relu(conv_bn(x, pre + ".conv1", pre + ".bn1")))
            out = q(F.relu(conv_bn(out, pre + ".conv2", pre + ".bn2", stride=s, padding=1)))
            out = conv_bn(out, pre + ".conv3", pre + ".bn3")
            if (pre + ".downsample.0.weight") in p:
                residual = q(conv_bn(x, pre + ".downsample.0", pre + ".downsample.1", stride=s))
            x = q(F.relu(out + residual))
    return F.adaptive_avg_pool2d(x, 1).view(x.size(0), -1)
