"""Arch registry: the reference selects an arch by module file name (VRM:18-21) or by the
``model_type`` switch of the inference scripts (INF:372-385)."""
from importlib import import_module

_BY_MODEL_TYPE = {"t0": "turtle_arch", "t1": "turtle_t1_arch", "SR": "turtlesuper_t1_arch"}


def create_video_model(opt, model_type=None):
    """``create_video_model(opt)`` -> VRM:18-21; ``create_video_model(opt, model_type)`` -> INF:372-385
    (the reference's ``turtle_super_t1_arch`` typo, INF:380, is accepted as an alias)."""
    if model_type is not None:
        name = _BY_MODEL_TYPE.get(model_type)
        if name is None:
            print("Model type not defined")
            raise SystemExit()
    else:
        name = str(opt["model"]).lower()
        if name == "turtle_super_t1_arch":
            name = "turtlesuper_t1_arch"
    return import_module(f"{__name__}.{name}").make_model(opt)
