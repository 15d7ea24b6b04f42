"""Arch hyper-parameters of the reference's shipped options/*.yml, restated so that the bench,
smoke test and GPU tests do not need /root/reference at run time.  All six ymls share the same
network (options/Turtle_Deblur_Gopro.yml:8-58); they differ in the arch file (``model``) and, for
Davis, in the attention aliases MEST/CTS (Turtle_Denoise_Davis.yml:42-55).  A real yml parsed with
``yaml.safe_load`` can be passed to ``create_video_model`` just the same.
"""
import copy

_BASE = dict(
    n_colors=3, dim=64, Enc_blocks=[2, 6, 10], Middle_blocks=11, Dec_blocks=[10, 6, 2], num_refinement_blocks=2,
    use_both_input=False, num_heads=[1, 2, 4, 8], num_frames_tocache=3, ffn_expansion_factor=2.5,
    encoder1_attn_type1="ReducedAttn", encoder1_attn_type2="ReducedAttn", encoder1_ffw_type="FFW",
    encoder2_attn_type1="ReducedAttn", encoder2_attn_type2="ReducedAttn", encoder2_ffw_type="FFW",
    encoder3_attn_type1="Channel", encoder3_attn_type2="Channel", encoder3_ffw_type="GFFW",
    decoder1_attn_type1="Channel", decoder1_attn_type2="CHM", decoder1_ffw_type="GFFW",
    decoder2_attn_type1="Channel", decoder2_attn_type2="CHM", decoder2_ffw_type="GFFW",
    decoder3_attn_type1="Channel", decoder3_attn_type2="CHM", decoder3_ffw_type="GFFW",
    latent_attn_type1="FHR", latent_attn_type2="Channel", latent_attn_type3="FHR", latent_ffw_type="GFFW",
    refinement_attn_type1="ReducedAttn", refinement_attn_type2="ReducedAttn", refinement_ffw_type="GFFW",
    manual_seed=10,
)

_MODEL = {
    "Turtle_Deblur_Gopro": "Turtle_t1_arch", "Turtle_Denoise_Davis": "Turtle_t1_arch",
    "Turtle_Derain": "Turtle_arch", "Turtle_Derain_VRDS": "Turtle_t1_arch", "Turtle_Desnow": "Turtle_arch",
    "Turtle_SR_MVSR": "Turtlesuper_t1_arch",
}


def shipped(name: str) -> dict:
    """e.g. ``shipped("Turtle_Deblur_Gopro")`` -> flat opt dict accepted by make_model."""
    name = name[:-4] if name.endswith(".yml") else name
    o = copy.deepcopy(_BASE)
    o["model"] = _MODEL[name]
    if name == "Turtle_Denoise_Davis":
        for k in ("decoder1_attn_type2", "decoder2_attn_type2", "decoder3_attn_type2"):
            o[k] = "MEST"
        o["latent_attn_type1"] = o["latent_attn_type3"] = "CTS"
    return o
