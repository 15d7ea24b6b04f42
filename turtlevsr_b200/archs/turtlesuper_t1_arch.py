"""Drop-in for ``basicsr/models/archs/turtlesuper_t1_arch.py`` (SR_MVSR yml).

t1 plus a bilinear x4 upsample of the selected low-resolution frame before padding
(TS:975-978, 1049-1071); the output is cropped to 4H x 4W (TS:1051, 1137-1139).
"""
import torch.nn as nn

from ._common import TurtleNet, model_kwargs_from_opt


class TurtleSuper_t1(TurtleNet):   # reference class: TS:932
    variant = "super"

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.upsample_4x = nn.Upsample(scale_factor=4, mode="bilinear")   # parameter-free, API parity


def make_model(opt):
    return TurtleSuper_t1(**model_kwargs_from_opt(opt))


def create_video_model(opt):
    return make_model(opt)
