"""Generates tests/golden/*.pt from the ORACLE (oracle/) - the reference itself cannot run here (diffusers / peft are
absent and there are no weights), so these vectors pin the oracle against regressions, not against diffusers.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.step import OracleTrainer, StepConfig, make_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(family: str, rank: int, batch: int, hw: int):
    torch.set_num_threads(4)
    cfg = StepConfig(family=family, tiny=True, resolution=hw * 8, lora_rank=rank)
    orc = OracleTrainer(cfg, device="cpu")
    g = torch.Generator().manual_seed(7)
    for n, p in orc.unet.named_parameters():
        if "lora_B" in n:
            p.data.copy_((torch.randn(p.shape, generator=g) * 0.05).to(torch.bfloat16))
    inputs = make_inputs(cfg, batch=batch, latent_hw=hw, face_mask=True, train_ids=orc.train_ids)
    out = orc.step(inputs, do_optimizer=False)
    grads = {n: p.grad.float() for n, p in orc.unet.named_parameters() if p.grad is not None}
    pick = sorted(grads)[::max(1, len(grads) // 6)][:6]
    fix = {
        "family": family, "rank": rank, "batch": batch, "hw": hw,
        "img_loss": float(out["img_loss"]), "token_attention_loss": float(out["token_attention_loss"]),
        "token_std_loss": float(out["token_std_loss"]), "tot_loss": float(out["tot_loss"]),
        "model_pred": out["model_pred"].detach().float(), "noisy_latent": out["noisy_latent"].float(),
        "score0": out["attention_scores"][0].float(),
        "grad_norms": {n: float(grads[n].norm()) for n in pick},
        "grad_sample": {n: grads[n].flatten()[:64].clone() for n in pick[:2]},
    }
    torch.save(fix, os.path.join(HERE, f"step_{family}_r{rank}_b{batch}.pt"))
    print(family, {k: v for k, v in fix.items() if isinstance(v, float)})


def make_dense(family: str, batch: int, hw: int):
    """Full-UNet fine-tune (is_lora=False, disable_ti=True; BASELINE config 5) on the tiny nets."""
    torch.set_num_threads(4)
    cfg = StepConfig(family=family, tiny=True, resolution=hw * 8, is_lora=False, disable_ti=True)
    orc = OracleTrainer(cfg, device="cpu")
    inputs = make_inputs(cfg, batch=batch, latent_hw=hw, face_mask=True)
    out = orc.step(inputs, do_optimizer=False)
    grads = {n: p.grad.float() for n, p in orc.unet.named_parameters()}
    pick = sorted(grads)[::max(1, len(grads) // 12)][:12]
    fix = {"family": family, "batch": batch, "hw": hw, "img_loss": float(out["img_loss"]), "tot_loss": float(out["tot_loss"]),
           "model_pred": out["model_pred"].detach().float(), "n_params": len(grads),
           "grad_norms": {n: float(grads[n].norm()) for n in pick},
           "grad_sample": {n: grads[n].flatten()[:64].clone() for n in pick[:3]}}
    torch.save(fix, os.path.join(HERE, f"dense_step_{family}_b{batch}.pt"))
    print("dense", family, fix["img_loss"], fix["n_params"])


def make_vae():
    """AutoencoderKL.encode restatement (oracle/vae.py) on the tiny graph: image -> posterior parameters."""
    from oracle.vae import VAEConfig, build_vae
    torch.set_num_threads(4)
    orc = build_vae(VAEConfig.tiny(), seed=3)
    img = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(11)) * 2 - 1
    torch.save({"image": img, "moments": orc.encode_moments(img)}, os.path.join(HERE, "vae_tiny_moments.pt"))
    print("vae", float(orc.encode_moments(img).abs().mean()))


if __name__ == "__main__":
    make("sdxl", 8, 2, 8)
    make("sd15", 4, 1, 8)
    make_dense("sd15", 1, 8)
    make_vae()
