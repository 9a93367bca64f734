"""VAE-encode prologue timing on one B200 (CUDA events, after warm-up): published SD / SDXL VAE graph, random weights."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200.init import random_vae_encoder_state_dict  # noqa: E402
from sd_lora_trainer_b200.vae import VAEEncoderB200                   # noqa: E402

enc = VAEEncoderB200(random_vae_encoder_state_dict(seed=0, device="cuda"), device="cuda:0")
for side in (512, 1024):
    img = torch.rand(1, 3, side, side, device="cuda") * 2 - 1
    for _ in range(2):
        enc.encode_moments(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        enc.encode_moments(img)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"VAE encode {side}x{side}: {ms:.2f} ms / image  (peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)")
