"""bench.py — denoised latent frames/sec of the SVI clip-denoising hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|cfg1|tiny]

A "step" is ONE flow-matching denoise step of the clip: two Wan-DiT forwards (classifier-free guidance:
conditional + unconditional) + the fused CFG/Euler update, on synthetic inputs of BASELINE.json configs[1]
(Wan2.1-T2V-1.3B, 81 frames x 480 x 832 -> latents [1,16,21,60,104], L = 32760 tokens, 512 text tokens,
random-init weights).  The metric is latent frames per second of the full 50-step clip:
value = 21 / (50 * seconds_per_step).  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  The working set of one step (2.8 GB of weights + ~1.5 GB of activations) is far
larger than the 126 MB L2, so no explicit L2 flush is needed between steps ("inputs larger than L2").

N > 1 (launched by torchrun): ranks split as CFG-parallel x token-axis sequence-parallel
(distributed/sequence_parallel.py); total work is fixed -> "scaling": "strong".

--impl reference: the reference's own CPU path for the same step, timed on the host cores.  The reference
package itself cannot be installed offline (missing diffusers / xfuser / xformers / imageio and no GPU
attention library for CPU), so the arm runs the oracle port (oracle/wan_dit_oracle.py, pinned to the
reference by tests/golden) on a bounded sample: one DiT block at the workload's L, extrapolated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stable-video-infinity_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (cfg name, latent f, h, w, ctx_len, description)
    "cfg2": ("CFG_T2V_1_3B", 21, 60, 104, 512, "Wan2.1-T2V-1.3B 81fx480x832, 50-step CFG denoise (BASELINE configs[1])"),
    "cfg1": ("CFG_T2V_1_3B", 5, 40, 64, 512, "Wan2.1-T2V-1.3B 17fx320x512 (BASELINE configs[0])"),
    "tiny": ("CFG_TINY_T2V", 3, 16, 16, 64, "tiny 2-layer debug model"),
}
CLIP_STEPS = 50
CFG_SCALE = 5.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def build_model(cfg, device):
    from diffsynth.models.wan_video_dit import WanModel, precompute_freqs_cis_3d
    from tools import synth
    with torch.device("meta"):
        model = WanModel(**cfg)
    sd = synth.make_dit_state_dict_fast(cfg, seed=0, device=device, dtype=torch.bfloat16)
    model.load_state_dict(sd, assign=True)
    model.freqs = precompute_freqs_cis_3d(128)  # the meta-device construction above produced meta tables
    return model.eval()


def cpu_block_baseline(cfg, L, ctx_len, max_seconds=45.0):
    """Reference CPU path (oracle port) on the host cores: ONE DiT block forward at the workload's token count
    (1/(30*100) of a clip), fp32, all cores.  Returns (seconds per block, cores)."""
    from oracle import wan_dit_oracle as O
    from tools import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    one = dict(cfg, num_layers=1)
    sd = synth.make_dit_state_dict(one, seed=0)
    g = torch.Generator().manual_seed(0)
    d = cfg["dim"]
    x = torch.randn(1, L, d, generator=g)
    ctx = torch.randn(1, ctx_len, d, generator=g)
    t_mod = torch.randn(1, 6, d, generator=g) * 0.1
    f = max(1, L // 1560)
    ang = torch.rand(L, 64, dtype=torch.float64)
    with torch.no_grad():
        t0 = time.perf_counter()
        O.dit_block(sd, 0, x, ctx, t_mod, ang, one)
        dt = time.perf_counter() - t0
        if dt < max_seconds / 3:   # a second run if it is cheap enough (first run includes allocator warm-up)
            t0 = time.perf_counter()
            O.dit_block(sd, 0, x, ctx, t_mod, ang, one)
            dt = min(dt, time.perf_counter() - t0)
    return dt, cores


def run_reference_arm(args, wl_name, cfg, f, h, w, ctx_len, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = f * (h // 2) * (w // 2)
    times = []
    budget_s = 170.0        # keep the whole arm within a few minutes on any host
    t_start = time.perf_counter()
    warm = args.warmup
    for i in range(args.warmup + args.steps):
        dt, cores = cpu_block_baseline(cfg, L, ctx_len)
        if i == 0 and dt > 8.0:
            warm = min(warm, 1)          # a >8 s sample: one untimed warm-up is all the budget allows
        if i >= warm:
            times.append(dt)
        if len(times) >= args.steps or (times and time.perf_counter() - t_start > budget_s):
            break
    t_blk = sum(times) / len(times)
    step_s = t_blk * cfg["num_layers"] * 2            # one denoise step = 2 forwards x num_layers blocks (+ negligible rest)
    value = f / (CLIP_STEPS * step_s)
    line = {"impl": "reference", "metric": "denoised latent frames/sec (81fx480p, 50 steps)", "value": value,
            "unit": "latent_frames/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "description": desc, "tokens": L},
            "cpu_baseline": {"value": value, "unit": "latent_frames/s", "cores": cores, "kind": "port",
                             "sample": f"1 DiT block forward at L={L} per step sample ({t_blk:.2f} s), extrapolated x{cfg['num_layers']} blocks x2 CFG forwards x{CLIP_STEPS} steps"},
            "e2e": {"value": value, "unit": "latent_frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def vae_clip_leg(dev, f, h, w, ms_per_step):
    """Times WanVideoVAE.decode / .encode (public API, random-init weights) at the bench clip shape and folds them into a
    whole-clip figure: one SVI clip = encode(81 frames) + 50 CFG steps + decode(21 latent frames)."""
    from diffsynth.models.wan_video_vae import WanVideoVAE
    from tools import synth_vae
    vae = WanVideoVAE().eval()
    vae.load_state_dict(synth_vae.make_vae_state_dict(seed=0))
    vae.to(dev)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 16, f, h, w, generator=g).to(dev)
    frames = 4 * (f - 1) + 1
    video = (torch.rand(3, frames, 8 * h, 8 * w, generator=g) * 2 - 1).to(dev)
    out = {}
    for name, fn in (("decode", lambda: vae.decode(z, device=dev)), ("encode", lambda: vae.encode([video], device=dev))):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(b)
    clip_s = (out["encode"] + out["decode"] + CLIP_STEPS * ms_per_step) / 1e3
    return {"vae_decode_ms": out["decode"], "vae_encode_ms": out["encode"], "denoise_s": CLIP_STEPS * ms_per_step / 1e3,
            "clip_s": clip_s, "clips_per_hour": 3600.0 / clip_s, "pixel_frames_per_s": frames / clip_s,
            "note": "81f x 480x832 clip = VAE encode + 50 CFG steps (extrapolated from the timed steps) + VAE decode; "
                    "the umT5 / CLIP encoders add ~57 ms per clip (2 prompts x 25 ms + 7 ms, tools/enc_bench.py) and are not timed here"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg-parallel", action="store_true", help="N>1: split only the token axis (sp = N); for A/B runs")
    ap.add_argument("--breakdown", action="store_true", help="after the timed run, one extra step with every kernel call "
                    "bracketed by CUDA events; per-call-type sums go to stderr (not part of the JSON line)")
    ap.add_argument("--no-vae", action="store_true", help="skip the clip-boundary VAE leg (decode 21->81 frames, encode 81 frames)")
    args = ap.parse_args()
    from tools import synth
    cfg_name, f, h, w, ctx_len, desc = WORKLOADS[args.workload]
    cfg = getattr(synth, cfg_name)
    if args.impl == "reference":
        run_reference_arm(args, args.workload, cfg, f, h, w, ctx_len, desc)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    sp = None
    plan = "single"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from diffsynth.distributed.sequence_parallel import init_sp_groups
        sp = init_sp_groups(world, rank, cfg_parallel=not args.no_cfg_parallel)
        plan = sp.describe()
    from diffsynth import _native as nv
    from diffsynth.pipelines.svi_video import model_fn_wan_video
    from diffsynth.schedulers.flow_match import FlowMatchScheduler
    from oracle import wan_dit_oracle as O  # noqa: F401  (FLOP formula + cpu_baseline only; never on the product path)

    model = build_model(cfg, dev)
    eng = model.engine(dev)
    sched = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    sched.set_timesteps(CLIP_STEPS, shift=5.0)
    g = torch.Generator(device="cpu").manual_seed(0)
    lat_host = torch.randn(1, 16, f, h, w, generator=g).pin_memory()
    ctx_pos_host = torch.randn(1, ctx_len, cfg["text_dim"], generator=torch.Generator().manual_seed(1)).pin_memory()
    ctx_neg_host = torch.randn(1, ctx_len, cfg["text_dim"], generator=torch.Generator().manual_seed(2)).pin_memory()
    lat = lat_host.to(dev).clone()
    ctx_pos, ctx_neg = ctx_pos_host.to(dev), ctx_neg_host.to(dev)
    cp, cn = eng.context_state(ctx_pos), eng.context_state(ctx_neg)
    v_c, v_u = torch.empty_like(lat), torch.empty_like(lat)
    L = f * (h // 2) * (w // 2)

    def step(i):
        k = i % CLIP_STEPS
        t = float(sched.timesteps[k])
        sigma = float(sched.sigmas[k])
        nxt = float(sched.sigmas[k + 1]) if k + 1 < CLIP_STEPS else 0.0
        if sp is None:
            eng.forward(lat, t, cp, out=v_c)
            eng.forward(lat, t, cn, out=v_u)
            eng.k.cfg_euler_step(lat, v_c, v_u, CFG_SCALE, sigma, nxt)
        else:
            sp.cfg_parallel_step(eng, lat, t, cp, cn, v_c, v_u, CFG_SCALE, sigma, nxt)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync()
    eng.attn_events = []      # (start, end) CUDA events around every self-attention launch (roofline leg)
    launches0 = eng.k.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = eng.k.launches - launches0
    attn_ms = [a.elapsed_time(b) for a, b in eng.attn_events]
    eng.attn_events = None
    # CPU time to ENQUEUE one step, measured on an empty launch queue (inside the timed loop the CPU runs ahead until the
    # driver's queue is full and then only measures back-pressure): must stay well below ms_per_step or the job is
    # launch-bound
    host_t0 = time.perf_counter()
    step(args.warmup + args.steps)
    host_ms = (time.perf_counter() - host_t0) * 1e3
    sync()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = f / (CLIP_STEPS * ms_per_step / 1e3)
    flops_step = 2 * O.dit_forward_flops(cfg, L, ctx_len)

    # ---- e2e leg: same step through the public API (model_fn_wan_video) with HOST buffers: every step copies the
    # latents and both prompt embeddings from pinned host memory and reads the updated latents back.
    # With N > 1 every rank does the same from its own pinned buffers and the step is the public multi-GPU call
    # (SequenceParallelGroup.cfg_parallel_step); the time is the max over ranks.
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty_like(lat_host).pin_memory()
        lat_h = lat_host.clone().pin_memory()

        def e2e_step(i):
            k = i % CLIP_STEPS
            ts = sched.timesteps[k].reshape(1)
            sigma = float(sched.sigmas[k])
            nxt = float(sched.sigmas[k + 1]) if k + 1 < CLIP_STEPS else 0.0
            x = lat_h.to(dev, non_blocking=True)
            c1 = ctx_pos_host.to(dev, non_blocking=True)
            c2 = ctx_neg_host.to(dev, non_blocking=True)
            if sp is None:
                vc = model_fn_wan_video(model, x, ts, c1)
                vu = model_fn_wan_video(model, x, ts, c2)
                nv.cfg_euler_step(x, vc, vu, CFG_SCALE, sigma, nxt)
            else:
                both = sp.cfg_groups == 1                     # CFG-parallel ranks need only their own branch's prompt
                cpx = eng.context_state(c1) if both or sp.cfg_idx == 0 else None
                cnx = eng.context_state(c2) if both or sp.cfg_idx == 1 else None
                sp.cfg_parallel_step(eng, x, float(ts[0]), cpx, cnx, v_c, v_u, CFG_SCALE, sigma, nxt)
            out_host.copy_(x, non_blocking=True)
            torch.cuda.current_stream().synchronize()     # the caller consumes the host result every step
            lat_h.copy_(out_host)

        for i in range(max(1, args.warmup // 2)):
            e2e_step(i)
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.steps):
            e2e_step(i)
        b.record()
        sync()
        e2e_ms = a.elapsed_time(b) / args.steps
        if world > 1:
            tt = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = tt.item()
        h2d = lat_host.numel() * 4 + 2 * ctx_pos_host.numel() * 4
        api = ("diffsynth.pipelines.svi_video.model_fn_wan_video x2 + svi_cfg_euler_step, pinned host buffers" if sp is None else
               "SequenceParallelGroup.cfg_parallel_step (WanDiTEngine.context_state + forward per rank), pinned host buffers per rank")
        e2e = {"value": f / (CLIP_STEPS * e2e_ms / 1e3), "unit": "latent_frames/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": lat_host.numel() * 4, "api": api}

    if args.breakdown and world == 1:
        eng.k.events = []
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        step(0)
        b1.record()
        bd = eng.k.breakdown()
        tot = b0.elapsed_time(b1)
        print(f"breakdown of one step ({tot:.1f} ms wall, {sum(v[0] for v in bd.values()):.1f} ms inside kernels)", file=sys.stderr)
        for tag, (ms_, n) in sorted(bd.items(), key=lambda kv: -kv[1][0]):
            print(f"  {ms_:9.2f} ms {100 * ms_ / tot:6.2f}%  n={n:4d}  avg={1e3 * ms_ / n:9.1f} us  {tag}", file=sys.stderr)

    # ---- clip leg ("for cfg-4, end-to-end clips/hour with VAE included"): the VAE work one SVI clip adds
    # around the 50 denoising steps — encode of the 81-frame conditioning video and decode of the 21 denoised latent frames.
    clip = None
    if not args.no_vae and world == 1 and args.workload == "cfg2":
        clip = vae_clip_leg(dev, f, h, w, ms_per_step)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    Ll = L if sp is None else sp.local_rows(L)
    attn_flops = 4.0 * Ll * L * cfg["dim"]       # algorithmic FLOPs of one self-attention launch (4 L_q L_k d)
    roof = None
    if attn_ms:
        a_ms = sum(attn_ms) / len(attn_ms)
        ach = attn_flops / (a_ms * 1e-3) / 1e12
        traffic = None   # DRAM bytes per launch from the committed ncu --set full capture (single-GPU shape only)
        tpath = os.path.join(ROOT, "profiles", "attn_traffic.json")
        if sp is None and args.workload == "cfg2" and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("bytes_per_launch")
        roof = {"kernel": "attn_fwd_kernel (self-attention)", "bound": "tensor", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "launches_timed": len(attn_ms), "avg_ms": a_ms,
                "share_of_step": sum(attn_ms) / ms}
    line = {"metric": "denoised latent frames/sec (81fx480p, 50 steps)", "value": value, "unit": "latent_frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "tokens": L, "text_tokens": ctx_len,
                       "cfg_scale": CFG_SCALE, "clip_steps": CLIP_STEPS, "parallelism": plan,
                       "kv_exchange": None if sp is None or sp.sp_size == 1 else
                       ("NVLink peer push (copy engine) consumed by flag-gated attention" if sp._peer is not None else "NCCL all-gather"),
                       "l2_policy": "inputs larger than L2 (weights 2.8 GB + activations per step)",
                       "numerics": "bf16 operands, fp32 accumulate/residual/norm/softmax"},
            "dit_tflops": flops_step / (ms_per_step * 1e-3) / 1e12,
            "dit_tflops_frac_of_peak": flops_step / (ms_per_step * 1e-3) / 1e12 / (peak * world),
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms, "clocks": clocks, "roofline": roof, "e2e": e2e, "clip": clip}
    if not args.no_cpu_baseline and world == 1:
        t_blk, cores = cpu_block_baseline(cfg, L, ctx_len)
        step_s = t_blk * cfg["num_layers"] * 2
        line["cpu_baseline"] = {"value": f / (CLIP_STEPS * step_s), "unit": "latent_frames/s", "cores": cores, "kind": "port",
                                "sample": f"1 DiT block forward at L={L} ({t_blk:.2f} s fp32), extrapolated x{cfg['num_layers']} blocks x2 CFG forwards x{CLIP_STEPS} steps"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
