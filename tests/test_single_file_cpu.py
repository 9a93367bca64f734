"""Single-file (LDM layout) -> diffusers-layout key conversion (trainer/single_file.py, the from_single_file step of
reference trainer/models.py:15-28).  The LDM-side parameter list is enumerated HERE from the published LDM UNetModel
constructor (channel bookkeeping restated independently of the converter), with every tensor tagged by a unique value; the
converted dict must have exactly the oracle UNet's (diffusers-layout) keys and shapes, and tagged tensors must land on the
diffusers names their LDM position implies.  No real checkpoint is available offline: parity unpinned."""
import pytest
import torch

P = "model.diffusion_model."


def _ldm_unet(mc, channel_mult, depth, mid_depth, ctx, linear_proj, adm=None, nres=2):
    sd, tag = {}, [0]

    def put(name, *shape):
        tag[0] += 1
        sd[P + name] = torch.full(shape, float(tag[0]))

    def res(p, cin, cout):
        put(f"{p}.in_layers.0.weight", cin); put(f"{p}.in_layers.0.bias", cin)
        put(f"{p}.in_layers.2.weight", cout, cin, 3, 3); put(f"{p}.in_layers.2.bias", cout)
        put(f"{p}.emb_layers.1.weight", cout, 4 * mc); put(f"{p}.emb_layers.1.bias", cout)
        put(f"{p}.out_layers.0.weight", cout); put(f"{p}.out_layers.0.bias", cout)
        put(f"{p}.out_layers.3.weight", cout, cout, 3, 3); put(f"{p}.out_layers.3.bias", cout)
        if cin != cout:
            put(f"{p}.skip_connection.weight", cout, cin, 1, 1); put(f"{p}.skip_connection.bias", cout)

    def st(p, ch, d):
        put(f"{p}.norm.weight", ch); put(f"{p}.norm.bias", ch)
        for nm in ("proj_in", "proj_out"):
            if linear_proj:
                put(f"{p}.{nm}.weight", ch, ch)
            else:
                put(f"{p}.{nm}.weight", ch, ch, 1, 1)
            put(f"{p}.{nm}.bias", ch)
        for j in range(d):
            b = f"{p}.transformer_blocks.{j}"
            for n in ("norm1", "norm2", "norm3"):
                put(f"{b}.{n}.weight", ch); put(f"{b}.{n}.bias", ch)
            for a, kdim in (("attn1", ch), ("attn2", ctx)):
                put(f"{b}.{a}.to_q.weight", ch, ch); put(f"{b}.{a}.to_k.weight", ch, kdim); put(f"{b}.{a}.to_v.weight", ch, kdim)
                put(f"{b}.{a}.to_out.0.weight", ch, ch); put(f"{b}.{a}.to_out.0.bias", ch)
            put(f"{b}.ff.net.0.proj.weight", 8 * ch, ch); put(f"{b}.ff.net.0.proj.bias", 8 * ch)
            put(f"{b}.ff.net.2.weight", ch, 4 * ch); put(f"{b}.ff.net.2.bias", ch)

    put("time_embed.0.weight", 4 * mc, mc); put("time_embed.0.bias", 4 * mc)
    put("time_embed.2.weight", 4 * mc, 4 * mc); put("time_embed.2.bias", 4 * mc)
    if adm:
        put("label_emb.0.0.weight", 4 * mc, adm); put("label_emb.0.0.bias", 4 * mc)
        put("label_emb.0.2.weight", 4 * mc, 4 * mc); put("label_emb.0.2.bias", 4 * mc)
    put("input_blocks.0.0.weight", mc, 4, 3, 3); put("input_blocks.0.0.bias", mc)
    ch, chans, i = mc, [mc], 1
    for level, mult in enumerate(channel_mult):
        for _ in range(nres):
            res(f"input_blocks.{i}.0", ch, mult * mc)
            ch = mult * mc
            if depth[level]:
                st(f"input_blocks.{i}.1", ch, depth[level])
            chans.append(ch)
            i += 1
        if level != len(channel_mult) - 1:
            put(f"input_blocks.{i}.0.op.weight", ch, ch, 3, 3); put(f"input_blocks.{i}.0.op.bias", ch)
            chans.append(ch)
            i += 1
    res("middle_block.0", ch, ch); st("middle_block.1", ch, mid_depth); res("middle_block.2", ch, ch)
    i = 0
    for level, mult in reversed(list(enumerate(channel_mult))):
        for k in range(nres + 1):
            res(f"output_blocks.{i}.0", ch + chans.pop(), mult * mc)
            ch = mult * mc
            sub = 1
            if depth[level]:
                st(f"output_blocks.{i}.1", ch, depth[level])
                sub = 2
            if level and k == nres:
                put(f"output_blocks.{i}.{sub}.conv.weight", ch, ch, 3, 3); put(f"output_blocks.{i}.{sub}.conv.bias", ch)
            i += 1
    put("out.0.weight", ch); put("out.0.bias", ch)
    put("out.2.weight", 4, ch, 3, 3); put("out.2.bias", 4)
    return sd


@pytest.mark.parametrize("family", ["sd15", "sdxl"])
def test_ldm_unet_keys_map_onto_the_diffusers_layout(family):
    from oracle.unet import UNet2DConditionModel, UNetConfig
    from sd_lora_trainer_b200.trainer.single_file import convert_ldm_unet, is_single_file
    if family == "sd15":
        ldm = _ldm_unet(320, (1, 2, 4, 4), (1, 1, 1, 0), 1, 768, linear_proj=False)
        cfg = UNetConfig.sd15()
    else:
        ldm = _ldm_unet(320, (1, 2, 4), (0, 2, 10), 10, 2048, linear_proj=True, adm=2816)
        cfg = UNetConfig.sdxl()
    assert is_single_file(ldm.keys())
    with torch.device("meta"):
        ref = UNet2DConditionModel(cfg).state_dict()
    out = convert_ldm_unet(ldm)
    assert set(out) == set(ref), (sorted(set(out) - set(ref))[:5], sorted(set(ref) - set(out))[:5])
    for k, v in out.items():
        assert tuple(v.shape) == tuple(ref[k].shape), (k, v.shape, ref[k].shape)
    assert len(out) == len(ldm)                                   # one-to-one
    # spot checks of positions whose meaning is fixed by the LDM constructor order
    val = lambda name: float(ldm[P + name].flatten()[0])
    got = lambda name: float(out[name].flatten()[0])
    assert got("conv_in.weight") == val("input_blocks.0.0.weight")
    assert got("down_blocks.0.resnets.1.conv2.weight") == val("input_blocks.2.0.out_layers.3.weight")
    assert got("down_blocks.0.downsamplers.0.conv.weight") == val("input_blocks.3.0.op.weight")
    assert got("down_blocks.1.resnets.0.conv_shortcut.weight") == val("input_blocks.4.0.skip_connection.weight")
    assert got("mid_block.resnets.1.time_emb_proj.bias") == val("middle_block.2.emb_layers.1.bias")
    assert got("up_blocks.0.resnets.2.norm1.weight") == val("output_blocks.2.0.in_layers.0.weight")
    assert got("conv_norm_out.bias") == val("out.0.bias") and got("conv_out.weight") == val("out.2.weight")
    if family == "sd15":
        assert got("up_blocks.0.upsamplers.0.conv.weight") == val("output_blocks.2.1.conv.weight")          # no attention: sub 1
        assert got("up_blocks.1.upsamplers.0.conv.weight") == val("output_blocks.5.2.conv.weight")
        assert got("down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight") == \
            val("input_blocks.2.1.transformer_blocks.0.attn2.to_k.weight")
    else:
        assert got("add_embedding.linear_2.weight") == val("label_emb.0.2.weight")
        assert got("up_blocks.0.upsamplers.0.conv.weight") == val("output_blocks.2.2.conv.weight")
        assert got("up_blocks.1.attentions.2.transformer_blocks.1.ff.net.2.weight") == \
            val("output_blocks.5.1.transformer_blocks.1.ff.net.2.weight")
        assert got("mid_block.attentions.0.transformer_blocks.9.attn1.to_q.weight") == \
            val("middle_block.1.transformer_blocks.9.attn1.to_q.weight")


def test_open_clip_text_tower_maps_onto_clip_text_model_with_projection():
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    from sd_lora_trainer_b200.trainer.single_file import convert_open_clip
    D, layers, pre = 64, 2, "conditioner.embedders.1.model."
    sd, t = {}, [0]

    def put(name, *shape):
        t[0] += 1
        sd[pre + name] = torch.full(shape, float(t[0]))
    put("token_embedding.weight", 100, D); put("positional_embedding", 77, D); put("text_projection", D, D); put("logit_scale")
    put("ln_final.weight", D); put("ln_final.bias", D)
    for n in range(layers):
        b = f"transformer.resblocks.{n}."
        put(b + "ln_1.weight", D); put(b + "ln_1.bias", D); put(b + "ln_2.weight", D); put(b + "ln_2.bias", D)
        sd[pre + b + "attn.in_proj_weight"] = torch.arange(3 * D).float()[:, None].repeat(1, D)
        sd[pre + b + "attn.in_proj_bias"] = torch.arange(3 * D).float()
        put(b + "attn.out_proj.weight", D, D); put(b + "attn.out_proj.bias", D)
        put(b + "mlp.c_fc.weight", 4 * D, D); put(b + "mlp.c_fc.bias", 4 * D); put(b + "mlp.c_proj.weight", D, 4 * D); put(b + "mlp.c_proj.bias", D)
    out = convert_open_clip(sd)
    cfg = CLIPTextConfig(vocab_size=100, hidden_size=D, intermediate_size=4 * D, num_hidden_layers=layers, num_attention_heads=4,
                         max_position_embeddings=77, projection_dim=D, hidden_act="gelu")
    ref = {k: v for k, v in CLIPTextModelWithProjection(cfg).state_dict().items() if "position_ids" not in k}
    assert set(out) == set(ref), (sorted(set(out) - set(ref))[:5], sorted(set(ref) - set(out))[:5])
    assert all(tuple(out[k].shape) == tuple(ref[k].shape) for k in ref)
    k_w = out["text_model.encoder.layers.1.self_attn.k_proj.weight"]
    assert float(k_w[0, 0]) == D and float(k_w[-1, 0]) == 2 * D - 1          # the middle third of in_proj
    assert float(out["text_model.encoder.layers.0.self_attn.v_proj.bias"][0]) == 2 * D


def test_vae_encoder_keys_map_onto_the_diffusers_layout():
    """LDM first_stage_model encoder enumerated from its constructor (ch 128, ch_mult 1-2-4-4, two ResnetBlocks per level,
    nin_shortcut where the width changes, a single-head mid attention with 1x1-conv projections) vs oracle/vae.py's
    diffusers-layout AutoencoderKL encoder."""
    from oracle.vae import build_vae, state_dict_of
    from sd_lora_trainer_b200.trainer.single_file import convert_ldm_vae_encoder
    pre, sd = "first_stage_model.", {}

    def put(name, *shape):
        sd[pre + name + ".weight"] = torch.zeros(*shape)
        sd[pre + name + ".bias"] = torch.zeros(shape[0])

    def res(p, cin, cout):
        put(p + ".norm1", cin); put(p + ".conv1", cout, cin, 3, 3); put(p + ".norm2", cout); put(p + ".conv2", cout, cout, 3, 3)
        if cin != cout:
            put(p + ".nin_shortcut", cout, cin, 1, 1)
    put("encoder.conv_in", 128, 3, 3, 3)
    ch = 128
    for lvl, mult in enumerate((1, 2, 4, 4)):
        for n in range(2):
            res(f"encoder.down.{lvl}.block.{n}", ch, 128 * mult)
            ch = 128 * mult
        if lvl != 3:
            put(f"encoder.down.{lvl}.downsample.conv", ch, ch, 3, 3)
    res("encoder.mid.block_1", ch, ch)
    put("encoder.mid.attn_1.norm", ch)
    for n in ("q", "k", "v", "proj_out"):
        put(f"encoder.mid.attn_1.{n}", ch, ch, 1, 1)
    res("encoder.mid.block_2", ch, ch)
    put("encoder.norm_out", ch); put("encoder.conv_out", 8, ch, 3, 3); put("quant_conv", 8, 8, 1, 1)
    sd[pre + "decoder.conv_in.weight"] = torch.zeros(1)                      # ignored: not on the training path
    out = convert_ldm_vae_encoder(sd)
    ref = state_dict_of(build_vae())
    assert set(out) == set(ref), (sorted(set(out) - set(ref))[:5], sorted(set(ref) - set(out))[:5])
    for k in ref:
        assert tuple(out[k].shape) == tuple(ref[k].shape), (k, out[k].shape, ref[k].shape)
