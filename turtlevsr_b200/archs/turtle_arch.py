"""Drop-in for ``basicsr/models/archs/turtle_arch.py`` ("t0"; Derain / Desnow ymls).

Identical parameters to t1; the effective StateAlignBlock forward differs (T0:459-533):
sinusoidal position code added before q/k, q/k taken as dilated patches (k2/q2 unused),
``(local + topk)/2`` logits, and the aggregated result replaced by ``v`` (T0:521-523).
"""
from ._common import TurtleNet, model_kwargs_from_opt


class Turtle(TurtleNet):           # reference class: T0:855
    variant = "t0"


def make_model(opt):
    return Turtle(**model_kwargs_from_opt(opt))


def create_video_model(opt):
    return make_model(opt)
