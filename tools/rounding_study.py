"""CPU study: which bf16 rounding points of the native DiT path cost how much parity (development tool).

Runs the oracle (fp32) at BASELINE configs[0] (1.3B, L=3200, 30 blocks) once as the reference, then re-runs it with
bf16 rounding inserted at selectable points that mirror the native kernels' storage formats:

  lin     every Linear input rounded to bf16 (GEMM A operand)            -- the survey's "emulated" design (97.5 % inside)
          (= head + blk + emb: the head GEMM, the block GEMMs, the time / text embedding MLPs separately)
  patch   latents rounded to bf16 before the patch embedding (its GEMM A operand)
  qkv     q, k, v rounded to bf16 where the QKV GEMM stores them (before the q/k RMSNorm + RoPE)
  qk2     q, k rounded to bf16 after RMSNorm + RoPE (what attention consumes)
  v       v rounded to bf16 (attention operand)
  p       softmax numerators exp(s - m) rounded to bf16 before P.V (row sum taken from the unrounded values)

usage: python tools/rounding_study.py [variant ...]     (default: all)
Each variant is a '+'-joined set of the points above, e.g. lin+qk2+v+p.
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wan_dit_oracle as O  # noqa: E402
from tools import synth  # noqa: E402

POINTS = set()


def r(x):
    return x.to(torch.bfloat16).to(x.dtype)


_orig_lin = O._lin


def lin(sd, name, x):
    cls = "head" if name.startswith("head") else ("blk" if name.startswith("blocks") else "emb")
    if "lin" in POINTS or cls in POINTS:
        x = r(x)
    return _orig_lin(sd, name, x)


_orig_patchify = O.patchify


def patchify(sd, cfg, x):
    if "patch" in POINTS:
        x = r(x)
    return _orig_patchify(sd, cfg, x)


def attention(q, k, v, H):
    if "p" not in POINTS:
        return O_attention(q, k, v, H)
    B, Lq, D = q.shape
    hd = D // H
    qh = q.view(B, Lq, H, hd).transpose(1, 2)
    kh = k.view(B, -1, H, hd).transpose(1, 2)
    vh = v.view(B, -1, H, hd).transpose(1, 2)
    out = torch.empty_like(qh)
    for h in range(H):      # per head to bound memory
        s = (qh[:, h] @ kh[:, h].transpose(1, 2)) * hd ** -0.5
        m = s.amax(dim=-1, keepdim=True)
        e = torch.exp(s - m)
        out[:, h] = (r(e) @ vh[:, h]) / e.sum(dim=-1, keepdim=True)
    return out.transpose(1, 2).reshape(B, Lq, D)


O_attention = O.attention


def self_attention(sd, pre, x, angles, H, eps):
    q, k, v = lin(sd, pre + ".q", x), lin(sd, pre + ".k", x), lin(sd, pre + ".v", x)
    ssq, ssk = q.pow(2).mean(-1, keepdim=True), k.pow(2).mean(-1, keepdim=True)     # epilogue sums: before rounding
    if "qkv" in POINTS:
        q, k, v = r(q), r(k), r(v)
    q = q * torch.rsqrt(ssq + eps) * sd[pre + ".norm_q.weight"]
    k = k * torch.rsqrt(ssk + eps) * sd[pre + ".norm_k.weight"]
    q, k = O.rope_apply(q, angles, H), O.rope_apply(k, angles, H)
    if "qk2" in POINTS:
        q, k = r(q), r(k)
    if "v" in POINTS:
        v = r(v)
    return lin(sd, pre + ".o", attention(q, k, v, H))


def cross_attention(sd, pre, x, ctx, H, eps, has_image_input):
    q, k, v = lin(sd, pre + ".q", x), lin(sd, pre + ".k", ctx), lin(sd, pre + ".v", ctx)
    ssq, ssk = q.pow(2).mean(-1, keepdim=True), k.pow(2).mean(-1, keepdim=True)
    if "qkv" in POINTS:
        q, k, v = r(q), r(k), r(v)
    q = q * torch.rsqrt(ssq + eps) * sd[pre + ".norm_q.weight"]
    k = k * torch.rsqrt(ssk + eps) * sd[pre + ".norm_k.weight"]
    if "qk2" in POINTS:
        q, k = r(q), r(k)
    if "v" in POINTS:
        v = r(v)
    return lin(sd, pre + ".o", attention(q, k, v, H))


def stats(out, ref):
    err = (out - ref).abs()
    inside = (err <= 1e-3 + 1e-2 * ref.abs()).float().mean().item()
    return inside, err.max().item(), err.mean().item() / ref.std().item()


def main():
    variants = sys.argv[1:] or ["lin", "lin+qk2+v", "lin+qkv+qk2+v", "lin+qk2+v+p", "lin+qkv+qk2+v+p"]
    torch.set_num_threads(os.cpu_count())
    cfg = synth.CFG_T2V_1_3B
    layers = int(os.environ.get("LAYERS", cfg["num_layers"]))
    cfg = dict(cfg, num_layers=layers)
    init = os.environ.get("INIT", "normal")      # normal: synth's N(0, 1/fan_in); torch_default: nn.Linear's U(+-1/sqrt(fan_in))
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=0, init=init).items()}
    print(f"init={init} layers={layers}", flush=True)
    inp = synth.make_dit_inputs(cfg, 5, 40, 64, seed=0, ctx_len=512)
    ts = torch.tensor([1000.0])
    with torch.no_grad():
        t0 = time.time()
        ref = inp["x"] - O.dit_forward(sd, cfg, inp["x"], ts, inp["context"])
        print(f"fp32 reference: {time.time() - t0:.1f} s", flush=True)
        O._lin, O.self_attention, O.cross_attention, O.patchify = lin, self_attention, cross_attention, patchify
        for var in variants:
            POINTS.clear()
            POINTS.update(var.split("+"))
            t0 = time.time()
            out = inp["x"] - O.dit_forward(sd, cfg, inp["x"], ts, inp["context"])
            ins, mx, rel = stats(out, ref)
            print(f"{var:24s} inside={ins:.4f} max={mx:.4e} mean/std={rel:.4e}   ({time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
