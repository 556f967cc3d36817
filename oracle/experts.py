"""Oracle: expert denoisers as pure functions of (state_dict, inputs).

Test infrastructure only (see oracle/__init__.py).  Rows a3-a8 of SURVEY.md
section 8.  Every function takes a plain ``dict[str, Tensor]`` keyed exactly
like the reference module's ``state_dict()`` so that the checkpoint key set is
part of what is checked.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# state_dict layouts ("the checkpoint contract")
# ---------------------------------------------------------------------------
def _resblock_spec(prefix, cin, cout, tdim, spec):
    # reference: mnist/models/unet_small.py:22-37
    spec[f"{prefix}.block1.0.weight"] = (cin,)
    spec[f"{prefix}.block1.0.bias"] = (cin,)
    spec[f"{prefix}.block1.2.weight"] = (cout, cin, 3, 3)
    spec[f"{prefix}.block1.2.bias"] = (cout,)
    spec[f"{prefix}.time_mlp.1.weight"] = (cout, tdim)
    spec[f"{prefix}.time_mlp.1.bias"] = (cout,)
    spec[f"{prefix}.block2.0.weight"] = (cout,)
    spec[f"{prefix}.block2.0.bias"] = (cout,)
    spec[f"{prefix}.block2.3.weight"] = (cout, cout, 3, 3)
    spec[f"{prefix}.block2.3.bias"] = (cout,)
    if cin != cout:
        spec[f"{prefix}.res_conv.weight"] = (cout, cin, 1, 1)
        spec[f"{prefix}.res_conv.bias"] = (cout,)


def unet_small_spec(in_channels=1, base_dim=64, time_emb_dim=256, num_classes=None):
    """Key -> shape, in the reference's registration order.

    reference: mnist/models/unet_small.py:47-73, shapes/models/unet_small.py:57-90.
    """
    d, td = base_dim, time_emb_dim
    spec = OrderedDict()
    spec["time_mlp.1.weight"] = (td, d)
    spec["time_mlp.1.bias"] = (td,)
    spec["time_mlp.3.weight"] = (td, td)
    spec["time_mlp.3.bias"] = (td,)
    if num_classes is not None:
        spec["label_emb.weight"] = (num_classes, td)
    spec["init_conv.weight"] = (d, in_channels, 3, 3)
    spec["init_conv.bias"] = (d,)
    _resblock_spec("down1", d, d, td, spec)
    _resblock_spec("down2", d, 2 * d, td, spec)
    _resblock_spec("bot1", 2 * d, 4 * d, td, spec)
    _resblock_spec("up1", 6 * d, 2 * d, td, spec)
    _resblock_spec("up2", 3 * d, d, td, spec)
    spec["out_conv.weight"] = (in_channels, d, 1, 1)
    spec["out_conv.bias"] = (in_channels,)
    return spec


def mlp_2d_spec(num_hid=256, num_out=2):
    """reference: mnist/models/mlp_2d.py:6-16."""
    spec = OrderedDict()
    dims = [(num_hid, 1 + num_out), (num_hid, num_hid), (num_hid, num_hid), (num_out, num_hid)]
    for i, (o, k) in zip((0, 2, 4, 6), dims):
        spec[f"main.{i}.weight"] = (o, k)
        spec[f"main.{i}.bias"] = (o,)
    return spec


def _guided_block_spec(prefix, cin, cout, tdim, cdim, spec):
    # reference: src/compositional_diffusion_with_cross_attention.py:105-117, 86-97
    spec[f"{prefix}.time_mlp.weight"] = (cout, tdim)
    spec[f"{prefix}.time_mlp.bias"] = (cout,)
    spec[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3)
    spec[f"{prefix}.conv1.bias"] = (cout,)
    spec[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3)
    spec[f"{prefix}.conv2.bias"] = (cout,)
    spec[f"{prefix}.norm1.weight"] = (cout,)
    spec[f"{prefix}.norm1.bias"] = (cout,)
    spec[f"{prefix}.norm2.weight"] = (cout,)
    spec[f"{prefix}.norm2.bias"] = (cout,)
    # nn.MultiheadAttention registers separate q/k/v projections when kdim/vdim != embed_dim
    # and one packed in_proj_weight when they are equal (cout == 256 here: down2, bot2).
    if cout == cdim:
        spec[f"{prefix}.attn.attention.in_proj_weight"] = (3 * cout, cout)
    else:
        spec[f"{prefix}.attn.attention.q_proj_weight"] = (cout, cout)
        spec[f"{prefix}.attn.attention.k_proj_weight"] = (cout, cdim)
        spec[f"{prefix}.attn.attention.v_proj_weight"] = (cout, cdim)
    spec[f"{prefix}.attn.attention.in_proj_bias"] = (3 * cout,)
    spec[f"{prefix}.attn.attention.out_proj.weight"] = (cout, cout)
    spec[f"{prefix}.attn.attention.out_proj.bias"] = (cout,)
    spec[f"{prefix}.attn_norm.weight"] = (cout,)
    spec[f"{prefix}.attn_norm.bias"] = (cout,)


def guided_unet_spec(num_digits=10, num_colors=3, embed_dim=128):
    """reference: src/compositional_diffusion_with_cross_attention.py:144-181."""
    e = embed_dim
    c = 2 * e
    spec = OrderedDict()
    spec["digit_embedding.weight"] = (num_digits + 1, e)
    spec["color_embedding.weight"] = (num_colors + 1, e)
    spec["time_mlp.1.weight"] = (e, e)
    spec["time_mlp.1.bias"] = (e,)
    spec["init_conv.weight"] = (64, 3, 3, 3)
    spec["init_conv.bias"] = (64,)
    _guided_block_spec("down1", 64, 128, e, c, spec)
    _guided_block_spec("down2", 128, 256, e, c, spec)
    _guided_block_spec("bot1", 256, 512, e, c, spec)
    _guided_block_spec("bot2", 512, 256, e, c, spec)
    spec["up1.weight"] = (256, 128, 2, 2)
    spec["up1.bias"] = (128,)
    _guided_block_spec("up2", 384, 128, e, c, spec)
    spec["up3.weight"] = (128, 64, 2, 2)
    spec["up3.bias"] = (64,)
    _guided_block_spec("up4", 192, 64, e, c, spec)
    spec["out_conv.weight"] = (3, 128, 1, 1)
    spec["out_conv.bias"] = (3,)
    return spec


def _bn_spec(prefix, c, spec):
    spec[f"{prefix}.weight"] = (c,)
    spec[f"{prefix}.bias"] = (c,)
    spec[f"{prefix}.running_mean"] = (c,)
    spec[f"{prefix}.running_var"] = (c,)
    spec[f"{prefix}.num_batches_tracked"] = ()


def _score_block_spec(prefix, cin, cout, tdim, up, transform, spec):
    # reference: src/models/compose_grayscale_object_and_color.py:35-76
    spec[f"{prefix}.time_mlp.weight"] = (cout, tdim)
    spec[f"{prefix}.time_mlp.bias"] = (cout,)
    spec[f"{prefix}.conv1.weight"] = (cout, 2 * cin if up else cin, 3, 3)
    spec[f"{prefix}.conv1.bias"] = (cout,)
    if transform:
        spec[f"{prefix}.transform.weight"] = (cout, cout, 4, 4)
        spec[f"{prefix}.transform.bias"] = (cout,)
    spec[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3)
    spec[f"{prefix}.conv2.bias"] = (cout,)
    _bn_spec(f"{prefix}.bnorm1", cout, spec)
    _bn_spec(f"{prefix}.bnorm2", cout, spec)


def score_model_spec(in_channels=3, time_emb_dim=32):
    """ColoredMNISTScoreModel; reference: src/models/compose_grayscale_object_and_color.py:80-96."""
    td = time_emb_dim
    spec = OrderedDict()
    spec["time_mlp.1.weight"] = (4 * td, td)
    spec["time_mlp.1.bias"] = (4 * td,)
    spec["time_mlp.3.weight"] = (td, 4 * td)
    spec["time_mlp.3.bias"] = (td,)
    spec["initial_conv.weight"] = (32, in_channels, 3, 3)
    spec["initial_conv.bias"] = (32,)
    _score_block_spec("down1", 32, 64, td, False, True, spec)
    _score_block_spec("down2", 64, 128, td, False, True, spec)
    _score_block_spec("bot1", 128, 256, td, False, True, spec)
    spec["up_transpose_1.weight"] = (256, 128, 4, 4)
    spec["up_transpose_1.bias"] = (128,)
    _score_block_spec("up_block_1", 256, 128, td, False, False, spec)
    spec["up_transpose_2.weight"] = (128, 64, 4, 4)
    spec["up_transpose_2.bias"] = (64,)
    _score_block_spec("up_block_2", 128, 64, td, False, False, spec)
    spec["up_transpose_3.weight"] = (64, 32, 4, 4)
    spec["up_transpose_3.bias"] = (32,)
    _score_block_spec("up_block_3", 64, 32, td, False, False, spec)
    spec["output.weight"] = (in_channels, 32, 1, 1)
    spec["output.bias"] = (in_channels,)
    return spec


def beta_vae_spec(latent_dims=10):
    """BetaVAE; reference: src/4.3 best_of_both_worlds_3.py:95-114 (registration order of its state_dict)."""
    spec = OrderedDict()
    spec["encoder.0.weight"] = (32, 3, 4, 4)
    spec["encoder.0.bias"] = (32,)
    spec["encoder.2.weight"] = (64, 32, 4, 4)
    spec["encoder.2.bias"] = (64,)
    spec["encoder.4.weight"] = (128, 64, 4, 4)
    spec["encoder.4.bias"] = (128,)
    spec["encoder.7.weight"] = (256, 128 * 4 * 4)
    spec["encoder.7.bias"] = (256,)
    spec["fc_mu.weight"] = (latent_dims, 256)
    spec["fc_mu.bias"] = (latent_dims,)
    spec["fc_log_var.weight"] = (latent_dims, 256)
    spec["fc_log_var.bias"] = (latent_dims,)
    spec["decoder_input.weight"] = (256, latent_dims)
    spec["decoder_input.bias"] = (256,)
    spec["decoder.0.weight"] = (128 * 4 * 4, 256)
    spec["decoder.0.bias"] = (128 * 4 * 4,)
    spec["decoder.3.weight"] = (128, 64, 4, 4)
    spec["decoder.3.bias"] = (64,)
    spec["decoder.5.weight"] = (64, 32, 4, 4)
    spec["decoder.5.bias"] = (32,)
    spec["decoder.7.weight"] = (32, 3, 4, 4)
    spec["decoder.7.bias"] = (3,)
    return spec


def simple_unet_spec(num_classes=3):
    """SimpleUnet; reference: src/composing_conditional_diffusion_on_shape_and_color_6.py:161-206 (registration order)."""
    td = 32
    down, up = (64, 128, 256, 512, 1024), (1024, 512, 256, 128, 64)
    spec = OrderedDict()
    spec["time_mlp.1.weight"] = (td, td)
    spec["time_mlp.1.bias"] = (td,)
    spec["label_emb.weight"] = (num_classes + 1, td)
    spec["conv0.weight"] = (down[0], 3, 3, 3)
    spec["conv0.bias"] = (down[0],)
    for name, chans, is_up in (("downs", down, False), ("ups", up, True)):
        for i in range(4):
            ci, co, p = chans[i], chans[i + 1], f"{name}.{i}"
            spec[f"{p}.time_mlp.weight"] = (co, td)
            spec[f"{p}.time_mlp.bias"] = (co,)
            spec[f"{p}.conv1.weight"] = (co, 2 * ci if is_up else ci, 3, 3)
            spec[f"{p}.conv1.bias"] = (co,)
            spec[f"{p}.transform.weight"] = (co, co, 4, 4)
            spec[f"{p}.transform.bias"] = (co,)
            spec[f"{p}.conv2.weight"] = (co, co, 3, 3)
            spec[f"{p}.conv2.bias"] = (co,)
            for gn in ("gn1", "gn2"):
                spec[f"{p}.{gn}.weight"] = (co,)
                spec[f"{p}.{gn}.bias"] = (co,)
    spec["output.weight"] = (3, up[-1], 1, 1)
    spec["output.bias"] = (3,)
    return spec


def synth_state_dict(spec, seed):
    """Deterministic synthetic weights for a key->shape spec.

    Uses an explicit CPU generator and one ``randn`` per key in spec order, so
    the values do not depend on nn.Module default-init code paths.  Scales are
    chosen to keep activations O(1): conv/linear weights ~ N(0, 1/fan_in),
    biases ~ 0.1 N(0,1), norm gains 1 + 0.1 N(0,1), running_var in [0.5, 1.5].
    """
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()
    for name, shape in spec.items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        if name.endswith("running_var"):
            v = 1.0 + 0.5 * torch.tanh(r)
        elif name.endswith("running_mean"):
            v = 0.1 * r
        elif "embedding" in name or name.startswith("label_emb"):
            v = r
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if name.endswith("transform.weight") and False:
                pass
            v = r / math.sqrt(fan_in)
        elif leaf == "weight" or leaf.endswith("_weight"):
            v = 1.0 + 0.1 * r       # 1-D "weight" == a norm gain
        else:
            v = 0.1 * r             # biases
        sd[name] = v.contiguous()
    return sd


# ---------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------
def sinusoidal_pos_emb(t, dim):
    """reference: mnist/models/unet_small.py:12-19."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -k).to(device=t.device, dtype=t.dtype)
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def _resblock(sd, p, x, t_emb):
    """reference: mnist/models/unet_small.py:39-44 (dropout is identity in eval)."""
    h = F.group_norm(x, 8, sd[f"{p}.block1.0.weight"], sd[f"{p}.block1.0.bias"], eps=1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[f"{p}.block1.2.weight"], sd[f"{p}.block1.2.bias"], padding=1)
    te = F.linear(F.silu(t_emb), sd[f"{p}.time_mlp.1.weight"], sd[f"{p}.time_mlp.1.bias"])
    h = h + te[:, :, None, None]
    h = F.group_norm(h, 8, sd[f"{p}.block2.0.weight"], sd[f"{p}.block2.0.bias"], eps=1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[f"{p}.block2.3.weight"], sd[f"{p}.block2.3.bias"], padding=1)
    if f"{p}.res_conv.weight" in sd:
        r = F.conv2d(x, sd[f"{p}.res_conv.weight"], sd[f"{p}.res_conv.bias"])
    else:
        r = x
    return h + r


def unet_small_forward(sd, x, t, y=None, return_intermediates=False):
    """reference: mnist/models/unet_small.py:75-92, shapes/models/unet_small.py:92-120."""
    base_dim = sd["init_conv.weight"].shape[0]
    t_emb = sinusoidal_pos_emb(t, base_dim)
    t_emb = F.linear(t_emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    t_emb = F.silu(t_emb)
    t_emb = F.linear(t_emb, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
    if "label_emb.weight" in sd:
        if y is None:
            raise ValueError("Class labels `y` must be provided for a conditional UNet.")
        t_emb = t_emb + sd["label_emb.weight"][y]
    x0 = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=1)
    d1 = _resblock(sd, "down1", x0, t_emb)
    d2 = _resblock(sd, "down2", F.max_pool2d(d1, 2), t_emb)
    b1 = _resblock(sd, "bot1", F.max_pool2d(d2, 2), t_emb)
    u1 = F.interpolate(b1, scale_factor=2, mode="bilinear", align_corners=True)
    u1 = _resblock(sd, "up1", torch.cat([u1, d2], dim=1), t_emb)
    u2 = F.interpolate(u1, scale_factor=2, mode="bilinear", align_corners=True)
    u2 = _resblock(sd, "up2", torch.cat([u2, d1], dim=1), t_emb)
    out = F.conv2d(u2, sd["out_conv.weight"], sd["out_conv.bias"])
    if return_intermediates:
        return out, dict(t_emb=t_emb, x0=x0, d1=d1, d2=d2, b1=b1, u1=u1, u2=u2)
    return out


def mlp_2d_forward(sd, t, x):
    """reference: mnist/models/mlp_2d.py:17-20 (note the (t, x) argument order)."""
    h = torch.cat([t.view(-1, 1), x], dim=1)
    for i in (0, 2, 4):
        h = F.silu(F.linear(h, sd[f"main.{i}.weight"], sd[f"main.{i}.bias"]))
    return F.linear(h, sd["main.6.weight"], sd["main.6.bias"])


def _guided_block(sd, p, x, t_emb, context):
    """reference: src/compositional_diffusion_with_cross_attention.py:119-141.

    The attention has ONE key/value token, so the softmax is identically 1 and
    the attention output is out_proj(v_proj(context)) for every query pixel; it
    is written out in full here anyway (q/k projections, scaled softmax, 4
    heads) so that the degeneracy is checked rather than assumed.
    """
    cout = sd[f"{p}.conv1.weight"].shape[0]
    h = F.conv2d(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.group_norm(h, 8, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps=1e-5)
    h = h + F.linear(t_emb, sd[f"{p}.time_mlp.weight"], sd[f"{p}.time_mlp.bias"])[:, :, None, None]
    h = F.silu(h)
    b, c, hh, ww = h.shape
    seq = h.view(b, c, -1).permute(0, 2, 1)                      # (B, HW, C)
    bq, bk, bv = sd[f"{p}.attn.attention.in_proj_bias"].split(cout)
    if f"{p}.attn.attention.in_proj_weight" in sd:
        wq, wk, wv = sd[f"{p}.attn.attention.in_proj_weight"].split(cout)
    else:
        wq, wk, wv = (sd[f"{p}.attn.attention.{n}_proj_weight"] for n in "qkv")
    q = F.linear(seq, wq, bq)
    k = F.linear(context, wk, bk)
    v = F.linear(context, wv, bv)
    nh = 4
    hd = cout // nh
    qh = q.view(b, -1, nh, hd).transpose(1, 2)
    kh = k.view(b, -1, nh, hd).transpose(1, 2)
    vh = v.view(b, -1, nh, hd).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (att @ vh).transpose(1, 2).reshape(b, -1, cout)
    o = F.linear(o, sd[f"{p}.attn.attention.out_proj.weight"], sd[f"{p}.attn.attention.out_proj.bias"])
    seq = F.layer_norm(seq + o, (cout,), sd[f"{p}.attn_norm.weight"], sd[f"{p}.attn_norm.bias"], eps=1e-5)
    h = seq.permute(0, 2, 1).reshape(b, c, hh, ww)
    h = F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    h = F.group_norm(h, 8, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps=1e-5)
    return F.silu(h)


def guided_unet_forward(sd, x, t, digit_labels, color_labels):
    """reference: src/compositional_diffusion_with_cross_attention.py:183-208."""
    e = sd["digit_embedding.weight"].shape[1]
    t_emb = sinusoidal_pos_emb(t.to(torch.float32), e)
    t_emb = F.silu(F.linear(t_emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    ctx = torch.cat([sd["digit_embedding.weight"][digit_labels],
                     sd["color_embedding.weight"][color_labels]], dim=1).unsqueeze(1)
    x0 = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=1)
    d1 = _guided_block(sd, "down1", x0, t_emb, ctx)
    d2 = _guided_block(sd, "down2", F.max_pool2d(d1, 2), t_emb, ctx)
    b1 = _guided_block(sd, "bot1", F.max_pool2d(d2, 2), t_emb, ctx)
    b2 = _guided_block(sd, "bot2", b1, t_emb, ctx)
    u1 = F.conv_transpose2d(b2, sd["up1.weight"], sd["up1.bias"], stride=2)
    u2 = _guided_block(sd, "up2", torch.cat([u1, d2], dim=1), t_emb, ctx)
    u3 = F.conv_transpose2d(u2, sd["up3.weight"], sd["up3.bias"], stride=2)
    u4 = _guided_block(sd, "up4", torch.cat([u3, d1], dim=1), t_emb, ctx)
    return F.conv2d(torch.cat([u4, x0], dim=1), sd["out_conv.weight"], sd["out_conv.bias"])


def _bn_eval(sd, p, x):
    return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"],
                        sd[f"{p}.weight"], sd[f"{p}.bias"], training=False, eps=1e-5)


def _score_block(sd, p, x, t_emb, kind):
    """reference: src/models/compose_grayscale_object_and_color.py:53-60, 72-77."""
    h = _bn_eval(sd, f"{p}.bnorm1", F.relu(F.conv2d(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)))
    te = F.relu(F.linear(t_emb, sd[f"{p}.time_mlp.weight"], sd[f"{p}.time_mlp.bias"]))
    h = h + te[:, :, None, None]
    h = _bn_eval(sd, f"{p}.bnorm2", F.relu(F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)))
    if kind == "down":
        return F.conv2d(h, sd[f"{p}.transform.weight"], sd[f"{p}.transform.bias"], stride=2, padding=1)
    return h


def score_model_forward(sd, x, t):
    """ColoredMNISTScoreModel.forward; reference: src/models/compose_grayscale_object_and_color.py:98-112."""
    td = sd["time_mlp.3.weight"].shape[0]
    t_emb = sinusoidal_pos_emb(t, td)
    t_emb = F.relu(F.linear(t_emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    t_emb = F.linear(t_emb, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
    x1 = F.conv2d(x, sd["initial_conv.weight"], sd["initial_conv.bias"], padding=1)
    x2 = _score_block(sd, "down1", x1, t_emb, "down")
    x3 = _score_block(sd, "down2", x2, t_emb, "down")
    xb = _score_block(sd, "bot1", x3, t_emb, "down")
    u1 = F.conv_transpose2d(xb, sd["up_transpose_1.weight"], sd["up_transpose_1.bias"], stride=2, padding=1)
    u1 = _score_block(sd, "up_block_1", torch.cat([u1, x3], dim=1), t_emb, "conv")
    u2 = F.conv_transpose2d(u1, sd["up_transpose_2.weight"], sd["up_transpose_2.bias"], stride=2, padding=1)
    u2 = _score_block(sd, "up_block_2", torch.cat([u2, x2], dim=1), t_emb, "conv")
    u3 = F.conv_transpose2d(u2, sd["up_transpose_3.weight"], sd["up_transpose_3.bias"], stride=2, padding=1)
    u3 = _score_block(sd, "up_block_3", torch.cat([u3, x1], dim=1), t_emb, "conv")
    return F.conv2d(u3, sd["output.weight"], sd["output.bias"])


def _simple_block(sd, p, x, t_emb, up):
    """Block.forward; reference: src/composing_conditional_diffusion_on_shape_and_color_6.py:175-181."""
    h = F.group_norm(F.relu(F.conv2d(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)), 8,
                     sd[f"{p}.gn1.weight"], sd[f"{p}.gn1.bias"])
    te = F.relu(F.linear(t_emb, sd[f"{p}.time_mlp.weight"], sd[f"{p}.time_mlp.bias"]))
    h = h + te[:, :, None, None]
    h = F.group_norm(F.relu(F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)), 8,
                     sd[f"{p}.gn2.weight"], sd[f"{p}.gn2.bias"])
    if up:
        return F.conv_transpose2d(h, sd[f"{p}.transform.weight"], sd[f"{p}.transform.bias"], stride=2, padding=1)
    return F.conv2d(h, sd[f"{p}.transform.weight"], sd[f"{p}.transform.bias"], stride=2, padding=1)


def simple_unet_forward(sd, x, timestep, y):
    """SimpleUnet.forward; reference: src/composing_conditional_diffusion_on_shape_and_color_6.py:208-221."""
    t_emb = F.relu(F.linear(sinusoidal_pos_emb(timestep.float(), 32), sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    emb = t_emb + F.embedding(y, sd["label_emb.weight"])
    x = F.conv2d(x, sd["conv0.weight"], sd["conv0.bias"], padding=1)
    residuals = []
    for i in range(4):
        x = _simple_block(sd, f"downs.{i}", x, emb, False)
        residuals.append(x)
    for i in range(4):
        x = _simple_block(sd, f"ups.{i}", torch.cat((x, residuals.pop()), dim=1), emb, True)
    return F.conv2d(x, sd["output.weight"], sd["output.bias"])


def beta_vae_decode(sd, z):
    """BetaVAE.decode; reference: src/4.3 best_of_both_worlds_3.py:107-126 (decoder_input, decoder, decode)."""
    h = F.linear(z, sd["decoder_input.weight"], sd["decoder_input.bias"])
    h = F.relu(F.linear(h, sd["decoder.0.weight"], sd["decoder.0.bias"])).unflatten(1, (128, 4, 4))
    h = F.relu(F.conv_transpose2d(h, sd["decoder.3.weight"], sd["decoder.3.bias"], stride=2, padding=1))
    h = F.relu(F.conv_transpose2d(h, sd["decoder.5.weight"], sd["decoder.5.bias"], stride=2, padding=1))
    return torch.sigmoid(F.conv_transpose2d(h, sd["decoder.7.weight"], sd["decoder.7.bias"], stride=2, padding=1))


def save_image_quantize(x):
    """torchvision.utils.save_image's quantisation (torchvision/utils.py: mul(255).add_(0.5).clamp_(0, 255).to(uint8))."""
    return x.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)


# ---------------------------------------------------------------------------
# Hutchinson divergence (row a11)
# ---------------------------------------------------------------------------
def hutchinson_vjp_div(fn, x, probe):
    """eps_hat and v^T J v through reverse mode, as the reference does.

    reference: shapes/compose_images_ito.py:46-63 (``vector_field``): the VJP
    v^T J is taken with ``torch.autograd.grad(eps_hat, x, grad_outputs=v)`` and
    dotted with the same probe v.
    """
    with torch.enable_grad():
        xc = x.clone().requires_grad_(True)
        eps = fn(xc)
        vj = torch.autograd.grad(eps, xc, grad_outputs=probe, create_graph=False)[0]
    div = (vj * probe).flatten(1).sum(dim=1)
    return eps.detach(), div.detach()
