import torch, numpy as np, sys
import torch.nn.functional as F
sys.path.insert(0,'.')
from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer
from chatts_b200.weights import synthetic_state_dict
from oracle import decoder as od, merge as om, ts_encoder as ote
from tests.test_gpu_model import _demo_series, _oracle_cfg

def fwd_fused(x, w, cfg, dt):
    """materialise only: xn, q/k/v, P, ao, h, act (dtype); everything between in fp32 (fp32-accumulate matmuls on dtype inputs)"""
    nh, nkv = int(cfg["num_attention_heads"]), int(cfg["num_key_value_heads"])
    d = int(cfg.get("head_dim") or cfg["hidden_size"] // nh); eps=float(cfg.get("rms_norm_eps",1e-6))
    T=x.shape[0]
    cos,sin = od.rope_tables(cfg, T, dt); cos,sin=cos.float(),sin.float()
    lin=lambda a,W,b=None: F.linear(a.float(), W.float(), None if b is None else b.float())
    def norm(h,wn):
        hf=h.float(); return (wn.float()*hf*torch.rsqrt(hf.pow(2).mean(-1,keepdim=True)+eps)).to(dt)
    h=x
    for l in range(int(cfg["num_hidden_layers"])):
        pre=f"model.layers.{l}."
        a=norm(h,w[pre+"input_layernorm.weight"])
        q=lin(a,w[pre+"self_attn.q_proj.weight"],w.get(pre+"self_attn.q_proj.bias")).view(T,nh,d)
        k=lin(a,w[pre+"self_attn.k_proj.weight"],w.get(pre+"self_attn.k_proj.bias")).view(T,nkv,d)
        v=lin(a,w[pre+"self_attn.v_proj.weight"],w.get(pre+"self_attn.v_proj.bias")).view(T,nkv,d).to(dt)
        q,k=od.apply_rope(q,k,cos,sin); q,k=q.to(dt),k.to(dt)
        o=od.attention(q,k,v,nh//nkv,0)      # P rounded to dt inside, out dt
        h=(h.float()+lin(o,w[pre+"self_attn.o_proj.weight"])).to(dt)
        a=norm(h,w[pre+"post_attention_layernorm.weight"])
        g=(F.silu(lin(a,w[pre+"mlp.gate_proj.weight"]))*lin(a,w[pre+"mlp.up_proj.weight"])).to(dt)
        h=(h.float()+lin(g,w[pre+"mlp.down_proj.weight"])).to(dt)
    return norm(h,w["model.norm.weight"])

for dt in (torch.float16, torch.bfloat16):
  for seed in (99, 1, 7):
    cfg = ChatTSConfig.tiny()
    sd = synthetic_state_dict(cfg, seed=seed, device="cpu", dtype=dt, std=0.05)
    proc = ChatTSProcessor(SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id, cfg.eos_token_id), cfg)
    ts1,_ = _demo_series()
    enc = proc(text=["Describe <ts><ts/> please"], timeseries=[ts1], return_tensors="pt")
    sd32 = {k: v.float() for k, v in sd.items()}
    ts_w32 = {k[len("ts_encoder."):]: v.float() for k, v in sd.items() if k.startswith("ts_encoder.")}
    ts_w = {k[len("ts_encoder."):]: v for k, v in sd.items() if k.startswith("ts_encoder.")}
    feats32, pc = ote.forward(enc["timeseries"].to(dt).float(), cfg.ts, ts_w32)
    emb32 = om.hf_merge(enc["input_ids"], enc["attention_mask"], sd32["model.embed_tokens.weight"], feats32, pc.tolist(), cfg.ts_token_start_index)[0]
    ref32 = od.logits(od.forward_hidden(emb32, sd32, _oracle_cfg(cfg), od.State(cfg.num_hidden_layers))[-1:], sd32)
    feats, pc = ote.forward(enc["timeseries"].to(dt), cfg.ts, ts_w)
    emb = om.hf_merge(enc["input_ids"], enc["attention_mask"], sd["model.embed_tokens.weight"], feats, pc.tolist(), cfg.ts_token_start_index)[0]
    ref16 = od.logits(od.forward_hidden(emb, sd, _oracle_cfg(cfg), od.State(cfg.num_hidden_layers))[-1:], sd)
    # fused rounding with dtype TS features (same input embedding as HF-dtype run)
    hf = fwd_fused(emb, sd, _oracle_cfg(cfg), dt)
    lgf = F.linear(hf[-1:].float(), sd["lm_head.weight"].float()).to(dt)
    # fused rounding + fp32 ts feats rounded once
    emb_b = om.hf_merge(enc["input_ids"], enc["attention_mask"], sd["model.embed_tokens.weight"], feats32.to(dt), pc.tolist(), cfg.ts_token_start_index)[0]
    hf2 = fwd_fused(emb_b, sd, _oracle_cfg(cfg), dt)
    lgf2 = F.linear(hf2[-1:].float(), sd["lm_head.weight"].float()).to(dt)
    e=lambda a: float((a.float()-ref32).abs().max()/ref32.abs().max())
    print(str(dt), seed, "hf-rounding:", round(e(ref16),5), " fused-rounding:", round(e(lgf),5), " fused + precise TS:", round(e(lgf2),5))
