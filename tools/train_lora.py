#!/usr/bin/env python
"""LoRA fine-tuning driver over chatts_b200.train (the loop of the external ChatTS-Training recipe, README.md:216-218, for
the records of chatts/align/uts_template_qa.py:127-131: one JSON object per line with "input", "output", "timeseries").

    # one GPU
    python tools/train_lora.py --model /path/to/ChatTS-8B --data train.jsonl --out adapter_dir --epochs 1
    # data parallel, one process per GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
        tools/train_lora.py --model /path/to/ChatTS-8B --data train.jsonl --out adapter_dir

Without --model the ChatTS-8B shape is instantiated with synthetic weights and the byte-level stand-in tokenizer (there is no
checkpoint offline): useful as an end-to-end check of the training path, not to learn anything.  The adapter is written in
peft's layout (adapter_model.safetensors + adapter_config.json): PeftModel.from_pretrained(...).merge_and_unload()
(demo/demo_lora.ipynb cells 3-4) and ChatTSForCausalLM.merge_lora() both load it.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default=None, help="checkpoint directory (config.json + safetensors + tokenizer); default: synthetic ChatTS-8B")
    ap.add_argument("--data", required=True, help="jsonl of {input, output, timeseries}")
    ap.add_argument("--out", required=True, help="adapter output directory")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--samples-per-step", type=int, default=8, help="records per rank and optimisation step")
    ap.add_argument("--micro-batch", type=int, default=4, help="records per forward/backward (gradient accumulation)")
    ap.add_argument("--r", type=int, default=16)
    ap.add_argument("--alpha", type=float, default=32.0)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--warmup-steps", type=int, default=0)
    ap.add_argument("--schedule", default="cosine", choices=["cosine", "linear", "constant"])
    ap.add_argument("--weight-decay", type=float, default=0.0)
    ap.add_argument("--max-grad-norm", type=float, default=1.0)
    ap.add_argument("--max-length", type=int, default=2048)
    ap.add_argument("--checkpoint", default=None, help="resume file (written every --checkpoint-every steps)")
    ap.add_argument("--checkpoint-every", type=int, default=0)
    ap.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float16"])
    ap.add_argument("--layers", type=int, default=0, help="synthetic model only: fewer decoder layers (smoke runs)")
    args = ap.parse_args()

    import torch.distributed as dist
    from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer
    from chatts_b200.model import ChatTSForCausalLM
    from chatts_b200.train import LoraTrainer, load_jsonl

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dt = torch.bfloat16 if args.dtype == "bfloat16" else torch.float16
    if args.model:
        model = ChatTSForCausalLM.from_pretrained(args.model, device=f"cuda:{local}", torch_dtype=dt, max_batch=1, max_seq_len=args.max_length)
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(args.model, trust_remote_code=True)
        tok.padding_side = "left"
    else:
        cfg = ChatTSConfig.chatts_8b()
        if args.layers:
            cfg.num_hidden_layers = args.layers
        model = ChatTSForCausalLM.from_synthetic(cfg, seed=1234, dtype=dt, max_batch=1, max_seq_len=args.max_length, use_cuda_graph=False)
        tok = SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id, cfg.eos_token_id)
    cfg = model.config
    proc = ChatTSProcessor(tok, cfg)
    tr = LoraTrainer(model, r=args.r, lora_alpha=args.alpha, lr=args.lr, weight_decay=args.weight_decay, max_grad_norm=args.max_grad_norm)
    records = load_jsonl(args.data)
    t0 = time.time()

    def on_step(step, loss, trainer):
        if rank == 0 and (step % 10 == 0):
            print(f"[train_lora] step {step} loss {float(loss[0]):.4f} lr {trainer.lr:.3e} grad_norm {float(trainer.norm_out[0]):.3f} "
                  f"({time.time() - t0:.0f} s)", flush=True)

    losses = tr.fit(proc, records, epochs=args.epochs, samples_per_step=args.samples_per_step, micro_batch=args.micro_batch,
                    lr_schedule=args.schedule, warmup_steps=args.warmup_steps, eos_token_id=cfg.eos_token_id, max_length=args.max_length,
                    on_step=on_step, checkpoint=args.checkpoint, checkpoint_every=args.checkpoint_every)
    if rank == 0:
        tr.save_adapter(args.out)
        print(f"[train_lora] {len(losses)} steps, loss {losses[0]:.4f} -> {losses[-1]:.4f}; adapter written to {args.out}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
