"""Drop-in for ``basicsr/models/archs/turtle_t1_arch.py`` (used by the Gopro / Davis ymls).

Same module-level API (``make_model``, ``create_video_model``, class ``Turtle_t1``), same
state-dict schema, same forward signature; the computation runs on the sm_100a kernels.
"""
from ._common import TurtleNet, model_kwargs_from_opt


class Turtle_t1(TurtleNet):        # reference class: T1:932
    variant = "t1"


def make_model(opt):               # T1:10-53
    return Turtle_t1(**model_kwargs_from_opt(opt))


def create_video_model(opt):       # T1:56-59
    return make_model(opt)
