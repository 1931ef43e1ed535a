"""GPU: SVI-Talk — the enable_multitalk DiT forward (AudioProjModel + per-frame audio cross-attention) and the three-way
guidance loop on the native kernels against the CPU oracle (pinned to the reference model by
tests/golden/dit_tiny_talk.npz)."""
import pytest
import torch

from tools import synth

pytestmark = pytest.mark.gpu


def _stats(out, ref):
    err = (out - ref).abs()
    return (err <= 1e-3 + 1e-2 * ref.abs()).float().mean().item(), err.max().item(), err.mean().item() / ref.std().item()


def _setup(seed, f=3, h=8, w=8):
    from diffsynth.models.wan_video_dit import WanModel
    from oracle import wan_dit_oracle as O
    cfg = synth.CFG_TINY_TALK
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth.make_dit_state_dict(cfg, seed=seed).items()}
    m = WanModel(**cfg).eval()
    m.load_state_dict(sd)
    m.to(device="cuda", dtype=torch.bfloat16)
    inp = synth.make_dit_inputs(cfg, f, h, w, seed=seed, ctx_len=24)
    audio = synth.make_audio_embed(4 * (f - 1) + 1, seed=seed).to(torch.bfloat16).float()
    return cfg, sd, m, inp, audio, O


def test_talk_forward_matches_oracle():
    from diffsynth.pipelines.svi_video_talk import model_fn_wan_talk_video, preprocess_audio
    cfg, sd, m, inp, audio, O = _setup(5)
    tup = preprocess_audio(audio)
    ts = torch.tensor([600.0])
    out = model_fn_wan_talk_video(m, inp["x"].cuda(), ts, inp["context"].cuda(), inp["clip_feature"].cuda(), inp["y"].cuda(),
                                  audio_embed_tuple=tuple(t.cuda() for t in tup)).float().cpu()
    ref = O.dit_forward(sd, cfg, inp["x"], ts, inp["context"], inp["clip_feature"], inp["y"], audio_embed_tuple=tup)
    base = O.dit_forward(sd, cfg, inp["x"], ts, inp["context"], inp["clip_feature"], inp["y"])
    inside, mx, rel = _stats(out, ref)
    print(f"talk forward: inside={inside:.4f} max={mx:.4e} mean/std={rel:.4e}; audio moves the output by {(ref - base).abs().max():.3f}")
    assert inside > 0.80 and rel < 4e-3 and (ref - base).abs().max() > 0.1
    # the audio tokens themselves (AudioProjModel on the GEMM kernel with the ReLU epilogue)
    eng = m.engine("cuda")
    st = eng.audio_state(tuple(t.cuda() for t in tup))
    assert st.n_frames == 3 and st.tokens == 32 and len(st.kv) == cfg["num_layers"]
    tok = O.audio_proj(sd, *tup).reshape(-1, 768)
    kv0 = (tok @ sd["blocks.0.audio_cross_attn.kv_linear.weight"].T + sd["blocks.0.audio_cross_attn.kv_linear.bias"])
    assert (st.kv[0].float().cpu() - kv0).abs().max().item() < 3e-2 * kv0.abs().max().item()


def test_talk_three_way_guidance_matches_oracle():
    from diffsynth import SVITalkVideoPipeline
    cfg, sd, m, inp, audio, O = _setup(6)
    b = synth.make_dit_inputs(cfg, 3, 8, 8, seed=7, ctx_len=24)
    pipe = SVITalkVideoPipeline(device="cuda", torch_dtype=torch.bfloat16)
    pipe.dit = m
    steps, st, sa = 2, 5.0, 4.0
    pipe.scheduler.set_timesteps(steps, shift=5.0)
    tup = pipe.preprocess_audio(audio)
    null = pipe.preprocess_audio(torch.zeros_like(audio))
    lat = pipe.denoise_latents_talk(inp["x"].cuda().float().clone(), inp["context"].cuda(), b["context"].cuda(),
                                    inp["clip_feature"].cuda(), inp["y"].cuda(), tup, null, {"text": st, "audio": sa}).cpu()
    sig = O.flow_match_sigmas(steps, 5.0)
    ref = inp["x"].clone()
    tup_c = tuple(t.float().cpu() for t in tup)
    null_c = tuple(t.float().cpu() for t in null)
    for i in range(steps):
        ts = (sig[i] * 1000).reshape(1)
        vc = O.dit_forward(sd, cfg, ref, ts, inp["context"], inp["clip_feature"], inp["y"], audio_embed_tuple=tup_c)
        vu = O.dit_forward(sd, cfg, ref, ts, b["context"], inp["clip_feature"], inp["y"], audio_embed_tuple=null_c)
        vd = O.dit_forward(sd, cfg, ref, ts, b["context"], inp["clip_feature"], inp["y"], audio_embed_tuple=tup_c)
        ref = O.flow_match_step(sig, i, vu + st * (vc - vd) + sa * (vd - vu), ref)          # svi_video_talk.py:460-462
    inside, mx, rel = _stats(lat, ref)
    print(f"talk 2-step 3-way CFG: inside={inside:.4f} max={mx:.4e} mean/std={rel:.4e}")
    assert inside > 0.5 and rel < 2e-2
