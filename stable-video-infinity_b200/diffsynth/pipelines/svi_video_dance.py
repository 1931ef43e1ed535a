"""SVIDanceVideoPipeline — SVI with a pose-video condition (reference diffsynth/pipelines/svi_video_dance.py).

Differences to SVIVideoPipeline, as in the reference: a pose stem (``dwpose_embedding``, built from the extra-module
keys of the SVI-Dance LoRA file, :254-275) turns ``humanpose_data`` into a token-space condition (:524-530) that
``model_fn_wan_video`` adds to the patch embedding of the conditional branch — and of the unconditional branch too when
``cond_wo_pose`` is set (:414-445).  Everything else (conditioning, CFG loop, VAE) is inherited.
"""
import torch
from PIL import Image
from tqdm import tqdm

from ..models.dwpose_embedding import DWPoseEmbeddingEngine, make_dwpose_embedding, pose_condition
from .svi_video import SVIVideoPipeline, TeaCache


class SVIDanceVideoPipeline(SVIVideoPipeline):
    def __init__(self, device="cuda", torch_dtype=torch.float16, tokenizer_path=None, is_test=False):
        super().__init__(device=device, torch_dtype=torch_dtype, tokenizer_path=tokenizer_path, is_test=is_test)
        self.dwpose_embedding = None
        self._pose_engine = None

    def fetch_models(self, model_manager):
        super().fetch_models(model_manager)
        if self.is_test:
            sd = {k.split("dwpose_embedding.")[1]: v for k, v in model_manager.state_dict_new_module.items()
                  if "dwpose_embedding" in k}
            if sd:
                self.load_pose_stem(sd)

    def load_pose_stem(self, state_dict):
        """state_dict: keys `N.weight` / `N.bias` of the 7-conv stem (strict, like the reference :275)."""
        dim = state_dict["12.weight"].shape[0]
        self.dwpose_embedding = make_dwpose_embedding(dim=dim)
        self.dwpose_embedding.load_state_dict(state_dict, strict=True)
        self._pose_engine = None

    @staticmethod
    def from_model_manager(model_manager, torch_dtype=None, device=None, use_usp=False, is_test=False):
        device = model_manager.device if device is None else device
        torch_dtype = model_manager.torch_dtype if torch_dtype is None else torch_dtype
        pipe = SVIDanceVideoPipeline(device=device, torch_dtype=torch_dtype, is_test=is_test)
        pipe.fetch_models(model_manager)
        if use_usp:
            pipe.enable_usp()
        return pipe

    def encode_pose(self, humanpose_data):
        if self.dwpose_embedding is None:
            raise RuntimeError("svi_b200: no pose stem loaded (dwpose_embedding.* keys of the SVI-Dance LoRA file)")
        if self._pose_engine is None:
            self._pose_engine = DWPoseEmbeddingEngine(self.dwpose_embedding, self.device)
        cond = pose_condition(self._pose_engine, humanpose_data)
        return cond.to(torch.bfloat16).to(torch.float32)          # the reference rounds the stem output to bf16 (:527)

    @torch.no_grad()
    def __call__(self, prompt, negative_prompt="", input_image=None, input_video=None, denoising_strength=1.0, seed=None,
                 rand_device="cpu", height=480, width=832, num_frames=81, cfg_scale=5.0, num_inference_steps=50,
                 sigma_shift=5.0, tiled=True, tile_size=(30, 52), tile_stride=(15, 26), tea_cache_l1_thresh=None,
                 tea_cache_model_id="", progress_bar_cmd=tqdm, humanpose_data=None, random_ref_frame=None,
                 use_controlnet=False, cond_wo_pose=False, args=None):
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
        neg = self.encode_prompt(negative_prompt, positive=False)
        image_emb = {}
        if input_image is not None and self.image_encoder is not None:
            ref_img = Image.fromarray(random_ref_frame.clone().cpu().numpy())
            if not isinstance(input_image, list):
                input_image = [input_image]
            image_emb = self.encode_images_adaptive(input_image, ref_img, num_frames, height, width, use_first_aug=False,
                                                    ref_pad_cfg=args.ref_pad_cfg, ref_pad_num=args.ref_pad_num)
        condition = self.encode_pose(humanpose_data) if humanpose_data is not None else None
        mk = lambda: (TeaCache(num_inference_steps, rel_l1_thresh=tea_cache_l1_thresh, model_id=tea_cache_model_id)
                      if tea_cache_l1_thresh is not None else None)
        scale = cfg_scale["text"] if isinstance(cfg_scale, dict) else cfg_scale
        bar = (lambda r: progress_bar_cmd(r)) if progress_bar_cmd is not None else (lambda r: r)
        latents = self.denoise_latents(latents, pos["context"], neg["context"], image_emb.get("clip_feature"),
                                       image_emb.get("y"), scale, bar, self.sp_group(), mk(), mk(),
                                       add_condition_posi=condition, add_condition_nega=condition if cond_wo_pose else None)
        frames = self.decode_video(latents, **tiler_kwargs)
        return self.tensor2video(frames[0])
