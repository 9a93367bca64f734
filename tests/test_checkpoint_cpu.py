"""The checkpoint FILE surface (reference trainer/checkpoint.py:84-221, embedding_handler.py:401-457):
file names, kohya / WebUI key naming, tensor layouts, and a save -> load round trip of the flat LoRA buffer."""
import json
import os
import re

import pytest
import torch

BF = torch.bfloat16


def _trainer(family="sdxl", rank=4):
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.unet import UNetB200
    arch = by_name(f"tiny_{family}")
    unet = UNetB200(arch, random_state_dict(arch, seed=0, device="cpu"), rank, device="cpu")
    g = torch.Generator().manual_seed(5)
    for s in unet.store.slots:
        s.B()[:, :s.r].copy_((torch.randn(s.fan_out, s.r, generator=g) * 0.05).to(BF))
    return unet


class _Handler:
    """Stand-in with the reference's save_embeddings contract (keys clip_l / clip_g)."""

    def __init__(self):
        self.rows = [torch.randn(3, 64).to(BF), torch.randn(3, 64).to(BF)]

    def save_embeddings(self, path):
        from safetensors.torch import save_file
        save_file({"clip_l": self.rows[0], "clip_g": self.rows[1]}, path)


def test_remove_delimiter_characters():
    from sd_lora_trainer_b200.trainer.checkpoint import remove_delimiter_characters
    assert remove_delimiter_characters("my face! (v2)") == "my_face_v2"
    assert remove_delimiter_characters("..a__b..") == "a_b"
    with pytest.raises(ValueError):
        remove_delimiter_characters("!!!")


@pytest.mark.parametrize("family", ["sdxl", "sd15"])
def test_save_checkpoint_files_keys_and_round_trip(tmp_path, family):
    from safetensors.torch import load_file
    from sd_lora_trainer_b200.trainer.checkpoint import load_lora_weights, save_checkpoint
    unet = _trainer(family)
    before = unet.store.params.clone()
    token_dict = {"TOK": "<s0><s1><s2>"}
    save_checkpoint(str(tmp_path), 100, unet, _Handler(), token_dict, True, [unet.store.params], family, name="eden concept!")
    names = sorted(os.listdir(tmp_path))
    assert names == sorted(["adapter_config.json", f"eden_concept_{family}_embeddings.safetensors",
                            f"eden_concept_{family}_lora.safetensors", "special_params.json"])
    assert json.load(open(tmp_path / "special_params.json")) == token_dict
    cfg = json.load(open(tmp_path / "adapter_config.json"))
    assert cfg["peft_type"] == "LORA" and cfg["r"] == 4 and cfg["lora_alpha"] == 4 and cfg["init_lora_weights"] == "gaussian"
    assert cfg["target_modules"] == ["to_k", "to_q", "to_v", "to_out.0", "conv2"]
    emb = load_file(str(tmp_path / f"eden_concept_{family}_embeddings.safetensors"))
    assert sorted(emb) == ["clip_g", "clip_l"]
    sd = load_file(str(tmp_path / f"eden_concept_{family}_lora.safetensors"))
    pat = re.compile(r"^lora_unet_[A-Za-z0-9_]+\.(lora_down\.weight|lora_up\.weight|alpha)$")
    assert all(pat.match(k) for k in sd), [k for k in sd if not pat.match(k)][:3]
    assert not any("base_model" in k for k in sd)
    assert len(sd) == 3 * len(unet.store.slots)
    # a linear and a conv module, by their reference-visible names
    lin = next(s for s in unet.store.slots if s.kind == "linear" and s.name.endswith("to_out.0"))
    key = "lora_unet_" + lin.name.replace(".", "_")
    assert sd[key + ".lora_down.weight"].shape == (4, lin.fan_in) and sd[key + ".lora_up.weight"].shape == (lin.fan_out, 4)
    assert int(sd[key + ".alpha"]) == 4
    conv = next(s for s in unet.store.slots if s.kind == "conv")
    key = "lora_unet_" + conv.name.replace(".", "_")
    assert sd[key + ".lora_down.weight"].shape == (4, conv.fan_in, 3, 3)
    assert sd[key + ".lora_up.weight"].shape == (conv.fan_out, 4, 1, 1)
    # round trip into a fresh executor: identical flat parameter buffer
    unet2 = _trainer(family)
    unet2.store.params.zero_()
    load_lora_weights(str(tmp_path / f"eden_concept_{family}_lora.safetensors"), unet2)
    assert torch.equal(unet2.store.params, before)


def test_save_checkpoint_rejects_what_the_step_does_not_train(tmp_path):
    from sd_lora_trainer_b200.trainer.checkpoint import save_checkpoint
    unet = _trainer()
    with pytest.raises(ValueError):
        save_checkpoint(str(tmp_path), 0, unet, None, {}, True, [1], "sd21", name="x")
    with pytest.raises(NotImplementedError):           # is_lora=False on an executor that was not built for it
        save_checkpoint(str(tmp_path), 0, unet, None, {}, False, [1], "sdxl", name="x")


def test_full_finetune_checkpoint_is_a_diffusers_unet_folder(tmp_path):
    """checkpoint.py:211-213: ``unet.save_pretrained``: config.json + diffusion_pytorch_model.safetensors, diffusers names."""
    import json
    from safetensors.torch import load_file
    from oracle.unet import UNet2DConditionModel, UNetConfig
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.trainer.checkpoint import save_checkpoint
    from sd_lora_trainer_b200.unet import UNetB200
    arch = by_name("tiny_sdxl")
    sd0 = random_state_dict(arch, seed=0, device="cpu")
    unet = UNetB200(arch, sd0, 0, device="cpu", dense=True)
    save_checkpoint(str(tmp_path), 7, unet, None, {"TOK": "<s0>"}, False, None, "sdxl", name="ft run")
    assert sorted(os.listdir(tmp_path)) == ["config.json", "diffusion_pytorch_model.safetensors", "special_params.json"]
    sd = load_file(str(tmp_path / "diffusion_pytorch_model.safetensors"))
    with torch.device("meta"):
        ref = UNet2DConditionModel(UNetConfig.tiny_sdxl())
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    for k, v in sd0.items():
        assert torch.equal(sd[k], v), k                    # native layouts (conv taps, padded channels) round-trip bit-exactly
    cfg = json.load(open(tmp_path / "config.json"))
    assert cfg["block_out_channels"] == [64, 128, 256] and cfg["down_block_types"][0] == "DownBlock2D"
