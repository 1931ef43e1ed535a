"""WanVideoPipeline — the plain Wan T2V/I2V pipeline surface (reference diffsynth/pipelines/wan_video.py:25-286).

Same hot loop as SVIVideoPipeline (shared implementation); differences follow the reference: scalar
``cfg_scale`` and ``encode_image`` (single first frame, zero padding) instead of ``encode_images_adaptive``.
The pose / replace variants of that file (:411-1583) are out of scope.
"""
import torch
from tqdm import tqdm

from .svi_video import SVIVideoPipeline, model_fn_wan_video  # noqa: F401  (re-exported like the reference)


class WanVideoPipeline(SVIVideoPipeline):
    @staticmethod
    def from_model_manager(model_manager, torch_dtype=None, device=None, use_usp=False):
        device = model_manager.device if device is None else device
        torch_dtype = model_manager.torch_dtype if torch_dtype is None else torch_dtype
        pipe = WanVideoPipeline(device=device, torch_dtype=torch_dtype)
        pipe.fetch_models(model_manager)
        if use_usp:
            pipe.enable_usp()
        return pipe

    def encode_image(self, image, num_frames, height, width):
        """reference wan_video.py:157-172: mask of the first frame + VAE latents of [image, zeros...]."""
        dev = self.device
        im = self.preprocess_image(image.resize((width, height))).to(device=dev, dtype=torch.float32)
        clip_context = self.image_encoder.encode_image([im])
        msk = torch.ones(1, num_frames, height // 8, width // 8, device=dev)
        msk[:, 1:] = 0
        msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
        msk = msk.view(1, msk.shape[1] // 4, 4, height // 8, width // 8).transpose(1, 2)[0]
        vae_input = torch.concat([im.transpose(0, 1), torch.zeros(3, num_frames - 1, height, width, device=dev)], dim=1)
        y = self.vae.encode([vae_input], device=dev)[0]
        y = torch.concat([msk, y.to(torch.float32)]).unsqueeze(0)
        return {"clip_feature": clip_context.to(self.torch_dtype), "y": y.to(self.torch_dtype)}

    @torch.no_grad()
    def __call__(self, prompt, negative_prompt="", input_image=None, input_video=None, denoising_strength=1.0, seed=None,
                 rand_device="cpu", height=480, width=832, num_frames=81, cfg_scale=5.0, num_inference_steps=50,
                 sigma_shift=5.0, tiled=True, tile_size=(30, 52), tile_stride=(15, 26), tea_cache_l1_thresh=None,
                 tea_cache_model_id="", progress_bar_cmd=tqdm, progress_bar_st=None):
        height, width = self.check_resize_height_width(height, width)
        if num_frames % 4 != 1:
            num_frames = (num_frames + 2) // 4 * 4 + 1
            print(f"Only `num_frames % 4 != 1` is acceptable. We round it up to {num_frames}.")
        tiler_kwargs = {"tiled": tiled, "tile_size": tile_size, "tile_stride": tile_stride}
        self.scheduler.set_timesteps(num_inference_steps, denoising_strength=denoising_strength, shift=sigma_shift)
        noise = self.generate_noise((1, 16, (num_frames - 1) // 4 + 1, height // 8, width // 8), seed=seed,
                                    device=rand_device, dtype=torch.float32)
        latents = noise.to(dtype=self.torch_dtype, device=self.device).to(torch.float32)
        if input_video is not None:
            vid = torch.stack(self.preprocess_images(input_video), dim=2).to(dtype=torch.float32, device=self.device)
            latents = self.scheduler.add_noise(self.encode_video(vid, **tiler_kwargs).to(torch.float32), latents,
                                               timestep=self.scheduler.timesteps[0])
        pos = self.encode_prompt(prompt, positive=True)
        neg = self.encode_prompt(negative_prompt, positive=False) if cfg_scale != 1.0 else pos
        image_emb = {}
        if input_image is not None and self.image_encoder is not None:
            image_emb = self.encode_image(input_image, num_frames, height, width)
        bar = (lambda r: progress_bar_cmd(r)) if progress_bar_cmd is not None else (lambda r: r)
        from .svi_video import TeaCache
        mk = lambda: (TeaCache(num_inference_steps, rel_l1_thresh=tea_cache_l1_thresh, model_id=tea_cache_model_id)
                      if tea_cache_l1_thresh is not None else None)     # reference wan_video.py:261-262
        latents = self.denoise_latents(latents, pos["context"], neg["context"], image_emb.get("clip_feature"),
                                       image_emb.get("y"), cfg_scale, bar, self.sp_group(), tea_cache_posi=mk(),
                                       tea_cache_nega=mk())
        frames = self.decode_video(latents, **tiler_kwargs)
        return self.tensor2video(frames[0])
