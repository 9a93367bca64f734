"""2-rank smoke of the data-parallel path on real GPUs with a tiny model: NCCL init, eager step, graph-captured step.
Prints a marker after every stage and dumps all Python stacks if a stage stalls (faulthandler)."""
import faulthandler
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("DP_DEBUG_STALL_S", "50")), repeat=False, file=sys.stderr)


def mark(msg):
    print(f"[rank {os.environ.get('RANK')}] {time.time():.1f} {msg}", file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
    mark("process group up")
    t = torch.ones(1024, device=dev)
    torch.distributed.all_reduce(t)
    torch.cuda.synchronize()
    mark(f"all_reduce ok {float(t[0])}")
    from sd_lora_trainer_b200.data import synthetic_inputs
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.step import StepConfig, TrainerB200
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection

    def build_text_encoders(seed):              # tiny SDXL-shaped CLIP pair, random init
        kw = dict(vocab_size=128, max_position_embeddings=77, bos_token_id=126, eos_token_id=127, pad_token_id=127)
        torch.manual_seed(seed)
        c1 = CLIPTextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                            hidden_act="quick_gelu", **kw)
        c2 = CLIPTextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                            hidden_act="gelu", projection_dim=64, **kw)
        return CLIPTextModel(c1), CLIPTextModelWithProjection(c2)
    use_graph = os.environ.get("DP_DEBUG_GRAPH", "1") == "1"
    cfg = StepConfig(family="sdxl", tiny=True, resolution=128, lora_rank=8)
    sd = random_state_dict(cfg.arch(), seed=0, device=dev)
    tes = build_text_encoders(seed=1)
    tr = TrainerB200(cfg, sd, tes, device=dev, process_group=torch.distributed.group.WORLD, use_cuda_graph=use_graph)
    mark("trainer built")
    batches = [synthetic_inputs("sdxl", 2, 128, cfg.n_tokens, seed=10 + rank * 7 + i, face_mask=True,
                                vae_scaling_factor=0.13025, tiny=True, pin=True) for i in range(2)]
    for i in range(4):
        out = tr.step(batches[i % 2])
        loss = float(out["tot_loss"])
        torch.cuda.synchronize()
        mark(f"step {i} done, loss {loss:.5f} (graph={use_graph})")
    # replicas must hold identical parameters after identical all-reduced updates
    p = tr.store.params.float()
    ref = p.clone()
    torch.distributed.broadcast(ref, src=0)
    mark(f"max |param - rank0 param| = {float((p - ref).abs().max()):.3e}")
    mark("done (leaving without destroy_process_group: it blocks under live CUDA graphs that captured NCCL kernels)")
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
