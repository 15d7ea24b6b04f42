"""CPU oracle for the Turtle inference hot path -- TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``turtlevsr_b200``) never routes through it.

It is a *functional* restatement (state-dict in, tensors out; no nn.Module tree) of what
the reference computes, written from the semantics in SURVEY.md Appendix A.  Every function
cites the reference lines it follows (T1 = basicsr/models/archs/turtle_t1_arch.py,
T0 = turtle_arch.py, TS = turtlesuper_t1_arch.py).

Pinning: the reference ships no tests/golden vectors, so this oracle is
pinned against *outputs of the reference itself*: ``oracle/make_golden.py`` imports the
reference arch files by path (in the build container, where /root/reference exists), checks
this oracle against them on identical weights/inputs and writes the fixtures under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle against those
fixtures everywhere (no /root/reference needed at run time).

All arithmetic is fp32 torch CPU ops (layout NCHW like the reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

ATTN_ALIASES = {"MEST": "CHM", "CTS": "FHR"}  # Turtle_Denoise_Davis.yml names


# ----------------------------------------------------------------------------------------
# architecture description
# ----------------------------------------------------------------------------------------
@dataclass
class LevelSpec:
    name: str
    dim: int
    heads: int
    attn_types: List[str]          # one per block
    ffw_type: str
    scale_patchsize: int = 1
    frames_tocache: int = 1


@dataclass
class ArchSpec:
    variant: str                   # 't1' | 't0' | 'super'
    n_colors: int
    dim: int
    use_both_input: bool
    ffn_expansion_factor: float
    levels: Dict[str, LevelSpec] = field(default_factory=dict)

    @staticmethod
    def from_opt(opt: dict, variant: Optional[str] = None) -> "ArchSpec":
        """Same key handling as make_model (T1:10-53): required keys have no defaults."""
        if variant is None:
            m = str(opt.get("model", "turtle_t1_arch")).lower()
            variant = {"turtle_arch": "t0", "turtle_t1_arch": "t1",
                       "turtlesuper_t1_arch": "super"}[m]
        dim = opt["dim"]
        heads = opt.get("num_heads", [1, 1, 1, 1])
        K = opt.get("num_frames_tocache", 1)
        enc, mid, dec = opt["Enc_blocks"], opt["Middle_blocks"], opt["Dec_blocks"]
        nref = opt.get("num_refinement_blocks", 1)

        def al(t):
            return ATTN_ALIASES.get(t, t)

        def lvl(name, d, h, n, t1, t2, ffw, sp=1, k=K):
            return LevelSpec(name, d, h, [al(t1)] * (n - 1) + [al(t2)], ffw, sp, k)

        spec = ArchSpec(variant, opt["n_colors"], dim, bool(opt["use_both_input"]),
                        opt.get("ffn_expansion_factor", 1))
        L = spec.levels
        L["encoder_level1"] = lvl("encoder_level1", dim, heads[0], enc[0],
                                  opt["encoder1_attn_type1"], opt["encoder1_attn_type2"], opt["encoder1_ffw_type"])
        L["encoder_level2"] = lvl("encoder_level2", dim * 2, heads[1], enc[1],
                                  opt["encoder2_attn_type1"], opt["encoder2_attn_type2"], opt["encoder2_ffw_type"])
        L["encoder_level3"] = lvl("encoder_level3", dim * 4, heads[2], enc[2],
                                  opt["encoder3_attn_type1"], opt["encoder3_attn_type2"], opt["encoder3_ffw_type"])
        lat_types = [al(opt["latent_attn_type1"])] + [al(opt["latent_attn_type2"])] * (mid - 2) + \
                    [al(opt["latent_attn_type3"])]
        L["latent"] = LevelSpec("latent", dim * 8, heads[3], lat_types, opt["latent_ffw_type"], 1, K)
        # NB the yml's "decoder1_*" keys configure decoder_level3 (T1:1009-1012) and so on.
        L["decoder_level3"] = lvl("decoder_level3", dim * 4, heads[2], dec[0],
                                  opt["decoder1_attn_type1"], opt["decoder1_attn_type2"], opt["decoder1_ffw_type"], 2)
        L["decoder_level2"] = lvl("decoder_level2", dim * 2, heads[1], dec[1],
                                  opt["decoder2_attn_type1"], opt["decoder2_attn_type2"], opt["decoder2_ffw_type"], 4)
        L["decoder_level1"] = lvl("decoder_level1", dim, heads[0], dec[2],
                                  opt["decoder3_attn_type1"], opt["decoder3_attn_type2"], opt["decoder3_ffw_type"], 8,
                                  2)  # hard-coded K=2, T1:1027
        L["refinement"] = lvl("refinement", dim, heads[0], nref,
                              opt["refinement_attn_type1"], opt["refinement_attn_type2"], opt["refinement_ffw_type"])
        return spec


# ----------------------------------------------------------------------------------------
# elementary ops
# ----------------------------------------------------------------------------------------
def channel_layernorm(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """Per-pixel LayerNorm over channels, biased variance, eps 1e-5 (T1:83-112)."""
    mu = x.mean(dim=1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=1, keepdim=True)
    if b is None:   # BiasFree variant does not subtract the mean in the numerator (T1:79-81)
        return x / torch.sqrt(var + 1e-5) * w.view(1, -1, 1, 1)
    return (x - mu) / torch.sqrt(var + 1e-5) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def conv1x1(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return F.conv2d(x, w, b)


def dwconv3x3(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return F.conv2d(x, w, b, padding=1, groups=x.shape[1])


def l2norm_rows(x: Tensor) -> Tensor:
    """F.normalize(dim=-1): x / max(||x||, 1e-12)."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def clipped_softmax_rows(z: Tensor) -> Tensor:
    """T1:115-132: zeros are excluded, softmax over the rest, then re-normalised by the sum."""
    dead = z == 0
    p = torch.softmax(z.masked_fill(dead, float("-inf")), dim=-1).masked_fill(dead, 0)
    return p / p.sum(dim=-1, keepdim=True)


def local_l1_mask(Hg: int, Wg: int, radius: int = 4) -> Tensor:
    """T1:448-464: [N,N] bool, true where the L1 grid distance is <= radius."""
    yy, xx = torch.meshgrid(torch.arange(Hg), torch.arange(Wg), indexing="ij")
    yy, xx = yy.reshape(-1), xx.reshape(-1)
    d = (yy[:, None] - yy[None, :]).abs() + (xx[:, None] - xx[None, :]).abs()
    return d <= radius


def sincos_posenc_2d(c: int, h: int, w: int) -> Tensor:
    """T0:412-439 sinusoidal 2-D encoding [c,h,w]."""
    if c % 4 != 0:
        raise ValueError("Cannot use sin/cos positional encoding with odd dimension (got dim={:d})".format(c))
    pe = torch.zeros(c, h, w)
    half = c // 2
    div = torch.exp(torch.arange(0., half, 2) * -(math.log(10000.0) / half))
    pw = torch.arange(0., w).unsqueeze(1) * div      # [w, half/2]
    ph = torch.arange(0., h).unsqueeze(1) * div
    pe[0:half:2] = torch.sin(pw).t().unsqueeze(1).expand(-1, h, -1)
    pe[1:half:2] = torch.cos(pw).t().unsqueeze(1).expand(-1, h, -1)
    pe[half::2] = torch.sin(ph).t().unsqueeze(2).expand(-1, -1, w)
    pe[half + 1::2] = torch.cos(ph).t().unsqueeze(2).expand(-1, -1, w)
    return pe


def to_dilated_patches(v: Tensor, ws: int) -> Tensor:
    """'b d (p1 h) (p2 w) -> b (h w) (p1 p2 d)' (T1:573): patch (i,j) samples rows p1*Hg+i."""
    b, d, h, w = v.shape
    Hg, Wg = h // ws, w // ws
    t = v.view(b, d, ws, Hg, ws, Wg).permute(0, 3, 5, 2, 4, 1)     # b Hg Wg p1 p2 d
    return t.reshape(b, Hg * Wg, ws * ws * d)


def from_dilated_patches(o: Tensor, ws: int, d: int, h: int, w: int) -> Tensor:
    """inverse of to_dilated_patches for [..., N, ws*ws*d] -> [..., d, h, w] (T1:602-604)."""
    lead = o.shape[:-2]
    Hg, Wg = h // ws, w // ws
    t = o.reshape(*lead, Hg, Wg, ws, ws, d)
    n = len(lead)
    t = t.permute(*range(n), n + 4, n + 2, n + 0, n + 3, n + 1)    # d p1 Hg p2 Wg
    return t.reshape(*lead, d, h, w)


# ----------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------
class Oracle:
    """Functional Turtle forward over a plain ``state_dict`` (name -> fp32 CPU tensor)."""

    def __init__(self, spec: ArchSpec, state_dict: Dict[str, Tensor]):
        self.spec = spec
        self.sd = {k[7:] if k.startswith("module.") else k: v.detach().float().cpu()
                   for k, v in state_dict.items()}
        self.trace: Optional[dict] = None      # when a dict, SAB records top-k indices etc.

    # -- helpers ---------------------------------------------------------------------
    def p(self, name: str) -> Optional[Tensor]:
        return self.sd.get(name)

    # -- feed-forwards ----------------------------------------------------------------
    def gated_ffw(self, pre: str, x: Tensor) -> Tensor:
        """T1:173-178."""
        u = dwconv3x3(conv1x1(x, self.sd[pre + "project_in.weight"], self.p(pre + "project_in.bias")),
                      self.sd[pre + "dwconv.weight"], self.p(pre + "dwconv.bias"))
        a, g = u.chunk(2, dim=1)
        return conv1x1(F.gelu(a) * g, self.sd[pre + "project_out.weight"], self.p(pre + "project_out.bias"))

    def plain_ffw(self, pre: str, x: Tensor) -> Tensor:
        """T1:204-210."""
        h = F.gelu(conv1x1(x, self.sd[pre + "conv4.weight"], self.sd[pre + "conv4.bias"]))
        return conv1x1(h, self.sd[pre + "conv5.weight"], self.sd[pre + "conv5.bias"]) * self.sd[pre + "gamma"]

    # -- attentions ---------------------------------------------------------------------
    def reduced_attn(self, pre: str, x: Tensor) -> Tensor:
        """T1:736-742."""
        h = conv1x1(x, self.sd[pre + "conv1.weight"], self.sd[pre + "conv1.bias"])
        h = F.gelu(dwconv3x3(h, self.sd[pre + "conv2.weight"], self.sd[pre + "conv2.bias"]))
        return conv1x1(h, self.sd[pre + "conv3.weight"], self.sd[pre + "conv3.bias"]) * self.sd[pre + "beta"]

    def channel_attn(self, pre: str, x: Tensor, heads: int, k_hist: Optional[Tensor] = None,
                     v_hist: Optional[Tensor] = None, keep_frames: Optional[int] = None):
        """ChannelAttention (T1:680-702) and, with history rows, FrameHistoryRouter (T1:243-286)."""
        b, c, h, w = x.shape
        qkv = dwconv3x3(conv1x1(x, self.sd[pre + "qkv.weight"], self.p(pre + "qkv.bias")),
                        self.sd[pre + "qkv_dwconv.weight"], self.p(pre + "qkv_dwconv.bias"))
        q, k, v = (t.reshape(b, heads, c // heads, h * w) for t in qkv.chunk(3, dim=1))
        q, k = l2norm_rows(q), l2norm_rows(k)
        if k_hist is not None and v_hist is not None:
            k = torch.cat([k_hist, k], dim=2)
            v = torch.cat([v_hist, v], dim=2)
        attn = torch.softmax((q @ k.transpose(-1, -2)) * self.sd[pre + "temperature"], dim=-1)
        out = (attn @ v).reshape(b, c, h, w)
        out = conv1x1(out, self.sd[pre + "project_out.weight"], self.p(pre + "project_out.bias"))
        if keep_frames is None:
            return out, None, None
        keep = int(keep_frames * c / heads)
        return out, k[:, :, -keep:], v[:, :, -keep:]

    def state_align(self, pre: str, x: Tensor, ws: int, keep: int, k_hist, v_hist):
        """StateAlignBlock effective forward: T1:548-610 (t1/super) or T0:459-533 (t0)."""
        b, c, h, w = x.shape
        t0 = self.spec.variant == "t0"
        x_qk = x + sincos_posenc_2d(c, h, w) if t0 else x
        # (every conv of the block takes the arch's ``bias`` option, T1:298-310; the shipped ymls leave it False)
        qk = dwconv3x3(conv1x1(x_qk, self.sd[pre + "qk.weight"], self.p(pre + "qk.bias")),
                       self.sd[pre + "qk_dwconv.weight"], self.p(pre + "qk_dwconv.bias"))
        q, k = qk.chunk(2, dim=1)
        v = dwconv3x3(conv1x1(x, self.sd[pre + "v.weight"], self.p(pre + "v.bias")),
                      self.sd[pre + "v_dwconv.weight"], self.p(pre + "v_dwconv.bias"))
        Hg, Wg = h // ws, w // ws
        if t0:
            q, k = to_dilated_patches(q, ws), to_dilated_patches(k, ws)
        else:
            k = F.conv2d(conv1x1(k, self.sd[pre + "k2.weight"], self.p(pre + "k2.bias")), self.sd[pre + "k2_dwconv.weight"],
                         self.p(pre + "k2_dwconv.bias"), stride=ws, padding=1, groups=2 * c)
            q = F.conv2d(conv1x1(q, self.sd[pre + "q2.weight"], self.p(pre + "q2.bias")), self.sd[pre + "q2_dwconv.weight"],
                         self.p(pre + "q2_dwconv.bias"), stride=ws, padding=1, groups=2 * c)
            assert q.shape[-2:] == (Hg, Wg)
            q = q.flatten(2).transpose(1, 2)          # b N 2c
            k = k.flatten(2).transpose(1, 2)
        v = to_dilated_patches(v, ws)                 # b N ws*ws*c
        q = l2norm_rows(q)[:, None, None]             # b 1 1 N D
        k = l2norm_rows(k)[:, None, None]
        v = v[:, None, None]
        if k_hist is not None and v_hist is not None:
            k = torch.cat([k_hist, k], dim=1)
            v = torch.cat([v_hist, v], dim=1)
        Fr = k.shape[1]
        S = (q @ k.transpose(-1, -2)) * self.sd[pre + "temperature"]          # b F 1 N N
        top_idx = torch.topk(S, k=5, dim=-1).indices
        keep_top = torch.zeros_like(S).scatter_(-1, top_idx, 1.0)
        local = local_l1_mask(Hg, Wg, 4).to(S.dtype)
        Z = S * keep_top + S * local
        if t0:
            Z = Z / 2
        Wt = clipped_softmax_rows(Z)
        if self.trace is not None:
            self.trace.setdefault(pre, []).append(
                {"topk": top_idx.clone(), "scores": S.clone(), "weights": Wt.clone()})
        O = v if t0 else Wt @ v                        # T0:521-523 discards the aggregation
        O = from_dilated_patches(O[:, :, 0], ws, c, h, w)                    # b F c h w
        O = conv1x1(O.reshape(b * Fr, c, h, w), self.sd[pre + "project_out.weight"],
                    self.p(pre + "project_out.bias")).reshape(b, Fr, c, h, w)
        return O, k[:, -keep:], v[:, -keep:]

    def causal_history(self, pre: str, x: Tensor, heads: int, scale_patch: int, keep: int, k_hist, v_hist):
        """CausalHistoryModel T1:627-662."""
        b, c, h, w = x.shape
        xs, k_new, v_new = self.state_align(pre + "spatial_aligner.", x, 2 * scale_patch, keep, k_hist, v_hist)
        Fr = xs.shape[1]
        kv = dwconv3x3(conv1x1(xs.reshape(b * Fr, c, h, w), self.sd[pre + "kv.weight"], self.p(pre + "kv.bias")),
                       self.sd[pre + "kv_dwconv.weight"], self.p(pre + "kv_dwconv.bias"))
        k, v = kv.chunk(2, dim=1)

        def rows(t):   # '(b f) (head c) h w -> b head (f c) (h w)'
            return t.reshape(b, Fr, heads, c // heads, h * w).permute(0, 2, 1, 3, 4).reshape(
                b, heads, Fr * (c // heads), h * w)
        k, v = l2norm_rows(rows(k)), rows(v)
        out, _, _ = self.channel_attn(pre + "ChanAttn.", x, heads, k, v, keep_frames=1)
        return out, k_new, v_new

    # -- block / level ---------------------------------------------------------------------
    def block(self, pre: str, lv: LevelSpec, attn_type: str, x: Tensor, k_hist=None, v_hist=None):
        """TurtleAttnBlock.forward T1:804-811."""
        kc = vc = None
        if attn_type != "NoAttn":
            y = channel_layernorm(x, self.sd[pre + "norm1.body.weight"], self.p(pre + "norm1.body.bias"))
            a = pre + "attn."
            if attn_type == "Channel":
                o, _, _ = self.channel_attn(a, y, lv.heads)
            elif attn_type == "ReducedAttn":
                o = self.reduced_attn(a, y)
            elif attn_type == "FHR":
                o, kc, vc = self.channel_attn(a, y, lv.heads, k_hist, v_hist, keep_frames=lv.frames_tocache)
            elif attn_type == "CHM":
                o, kc, vc = self.causal_history(a, y, lv.heads, lv.scale_patchsize, lv.frames_tocache,
                                                k_hist, v_hist)
            else:
                raise SystemExit(f"{attn_type}  Not defined")      # reference print+exit (T1:790-792)
            x = x + o
        y = channel_layernorm(x, self.sd[pre + "norm2.body.weight"], self.p(pre + "norm2.body.bias"))
        f = pre + "ffn."
        if lv.ffw_type == "GFFW":
            x = x + self.gated_ffw(f, y)
        elif lv.ffw_type == "FFW":
            x = x + self.plain_ffw(f, y)
        else:
            raise SystemExit(f"{lv.ffw_type}  Not defined")
        return x, kc, vc

    def level(self, name: str, x: Tensor, k_hist=None, v_hist=None):
        """LevelBlock.forward T1:856-865: only the last block sees the history."""
        lv = self.spec.levels[name]
        n = len(lv.attn_types)
        kc = vc = None
        for i, t in enumerate(lv.attn_types):
            last = i == n - 1
            x, kc, vc = self.block(f"{name}.transformer_blocks.{i}.", lv, t, x,
                                   k_hist if last else None, v_hist if last else None)
        return x, kc, vc

    def latent(self, x: Tensor, k1, v1, k2, v2):
        """LatentCacheBlock.forward T1:919-928: first and last block carry history."""
        lv = self.spec.levels["latent"]
        n = len(lv.attn_types)
        out = [None] * 4
        for i, t in enumerate(lv.attn_types):
            pre = f"latent.transformer_blocks.{i}."
            if i == 0:
                x, out[0], out[1] = self.block(pre, lv, t, x, k1, v1)
            elif i == n - 1:
                x, out[2], out[3] = self.block(pre, lv, t, x, k2, v2)
            else:
                x, _, _ = self.block(pre, lv, t, x)
        return (x, *out)

    # -- whole frame ---------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, pair: Tensor, k_cached: Optional[Sequence] = None, v_cached: Optional[Sequence] = None):
        """Turtle_t1.forward T1:1045-1132 / TurtleSuper_t1.forward TS:1049-1132 / Turtle.forward T0:968."""
        sd = self.sd
        B, _, C, H, W = pair.shape
        pair = pair.float()
        if k_cached is None:
            k_cached, v_cached = [None] * 8, [None] * 8
        if self.spec.use_both_input:
            img = torch.cat([pair[:, 0], pair[:, 1]], dim=1)
        else:
            img = pair[:, 1]
        if self.spec.variant == "super":
            img = F.interpolate(img, scale_factor=4, mode="bilinear")       # TS:975-978
            H, W = 4 * H, 4 * W
        ph, pw = (-img.shape[-2]) % 32, (-img.shape[-1]) % 32
        img = F.pad(img, (0, pw, 0, ph))                                    # T1:1134-1140
        current = img if not self.spec.use_both_input else img[:, C:]
        ks: List[Optional[Tensor]] = []
        vs: List[Optional[Tensor]] = []

        x = F.conv2d(img, sd["input_projection.weight"], self.p("input_projection.bias"), padding=1)
        e1, kc, vc = self.level("encoder_level1", x, k_cached[0], v_cached[0]); ks.append(kc); vs.append(vc)
        x = F.pixel_unshuffle(F.conv2d(e1, sd["down1_2.body.0.weight"], padding=1), 2)
        e2, kc, vc = self.level("encoder_level2", x, k_cached[1], v_cached[1]); ks.append(kc); vs.append(vc)
        x = F.pixel_unshuffle(F.conv2d(e2, sd["down2_3.body.0.weight"], padding=1), 2)
        e3, kc, vc = self.level("encoder_level3", x, k_cached[2], v_cached[2]); ks.append(kc); vs.append(vc)
        x = F.pixel_unshuffle(F.conv2d(e3, sd["down3_4.body.0.weight"], padding=1), 2)
        x, k4, v4, k5, v5 = self.latent(x, k_cached[3], v_cached[3], k_cached[4], v_cached[4])
        ks += [k4, k5]; vs += [v4, v5]

        x = F.pixel_shuffle(F.conv2d(x, sd["up4_3.body.0.weight"], padding=1), 2)
        x = conv1x1(torch.cat([x, e3], 1), sd["reduce_chan_level3.weight"], self.p("reduce_chan_level3.bias"))
        x, kc, vc = self.level("decoder_level3", x, k_cached[5], v_cached[5]); ks.append(kc); vs.append(vc)
        x = F.pixel_shuffle(F.conv2d(x, sd["up3_2.body.0.weight"], padding=1), 2)
        x = conv1x1(torch.cat([x, e2], 1), sd["reduce_chan_level2.weight"], self.p("reduce_chan_level2.bias"))
        x, kc, vc = self.level("decoder_level2", x, k_cached[6], v_cached[6]); ks.append(kc); vs.append(vc)
        x = F.pixel_shuffle(F.conv2d(x, sd["up2_1.body.0.weight"], padding=1), 2)
        x = conv1x1(torch.cat([x, e1], 1), sd["reduce_chan_level1.weight"], self.p("reduce_chan_level1.bias"))
        x, kc, vc = self.level("decoder_level1", x, k_cached[7], v_cached[7]); ks.append(kc); vs.append(vc)
        x, _, _ = self.level("refinement", x)
        out = F.conv2d(x, sd["ending.weight"], sd["ending.bias"], padding=1) + current
        return out[:, :, :H, :W], ks, vs

    @torch.no_grad()
    def run_clip(self, clip: Tensor):
        """Cached frame loop, VRM:110-129: pre = frame[j or j-1]; x = stack([pre, cur])."""
        outs, k, v = [], None, None
        for j in range(clip.shape[1]):
            pre = clip[:, j if j == 0 else j - 1]
            o, k, v = self.forward(torch.stack([pre, clip[:, j]], dim=1), k, v)
            outs.append(o)
        return torch.stack(outs, dim=1), k, v


# ----------------------------------------------------------------------------------------
# sparse restatement of the SAB selection (used to check the CUDA select/aggregate kernels)
# ----------------------------------------------------------------------------------------
def sab_select_sparse(qn: Tensor, kn: Tensor, tau: float, Hg: int, Wg: int, halve: bool = False):
    """Given normalised q [N,D] and keys [F,N,D]: per (f,i) the <=46 surviving (index, weight)
    pairs of T1:585-596, as dense [F,N,N] weights (small N only) plus the top-5 indices."""
    S = torch.einsum("nd,fmd->fnm", qn, kn) * tau
    top = torch.topk(S, k=5, dim=-1).indices
    Z = S * torch.zeros_like(S).scatter_(-1, top, 1.0) + S * local_l1_mask(Hg, Wg, 4).to(S.dtype)
    if halve:
        Z = Z / 2
    return clipped_softmax_rows(Z), top, S


def psnr(a: Tensor, b: Tensor) -> float:
    """metrics/psnr_ssim.py:63-67 on [0,1] data."""
    mse = (a.double() - b.double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 20.0 * math.log10(1.0 / math.sqrt(mse))


def randomize_gates(sd: Dict[str, Tensor], seed: int = 1234) -> Dict[str, Tensor]:
    """'live gates' weight variant: beta,gamma ~ N(0,0.1), temperatures ~ U(0.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith(".beta") or k.endswith(".gamma"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith(".temperature"):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        else:
            out[k] = v.clone()
    return out
