"""Same import path as the reference's models/resnet_custom.py; implementation in toad_b200."""
from toad_b200.resnet_custom import (Bottleneck_Baseline, ResNet_Baseline, load_pretrained_weights,  # noqa: F401
                                     resnet50_baseline)
