"""Flow-matching sigma schedule + Euler update (host logic; reference diffsynth/schedulers/flow_match.py).

The schedule is 50 scalars computed once per clip on the CPU; the per-step update of the latents runs in the
fused CFG+Euler kernel (``svi_cfg_euler_step``), for which this class supplies (sigma, sigma_next).
``step()`` keeps the reference's tensor semantics for callers that use the scheduler directly.
"""
import torch


class FlowMatchScheduler:
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, inverse_timesteps=False, extra_one_step=False, reverse_sigmas=False):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        self.sigma_max = sigma_max
        self.sigma_min = sigma_min
        self.inverse_timesteps = inverse_timesteps
        self.extra_one_step = extra_one_step
        self.reverse_sigmas = reverse_sigmas
        self.set_timesteps(num_inference_steps)

    def _sigmas(self, n, denoising_strength, shift):
        """reference flow_match.py:31-44: linspace, optional flip, time shift s -> shift*s/(1+(shift-1)s)."""
        start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            s = torch.linspace(start, self.sigma_min, n + 1)[:-1]
        else:
            s = torch.linspace(start, self.sigma_min, n)
        if self.inverse_timesteps:
            s = torch.flip(s, dims=[0])
        s = shift * s / (1 + (shift - 1) * s)
        if self.reverse_sigmas:
            s = 1 - s
        return s

    def get_timesteps(self, num_inference_steps, denoising_strength=1, shift=3.0):
        return self._sigmas(num_inference_steps, denoising_strength, shift) * self.num_train_timesteps

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False, shift=None):
        if shift is not None:
            self.shift = shift
        self.sigmas = self._sigmas(num_inference_steps, denoising_strength, self.shift)
        self.timesteps = self.sigmas * self.num_train_timesteps
        if training:  # reference :45-50 (bell-shaped loss weights)
            x = self.timesteps
            y = torch.exp(-2 * ((x - num_inference_steps / 2) / num_inference_steps) ** 2)
            y = y - y.min()
            self.linear_timesteps_weights = y * (num_inference_steps / y.sum())

    def _index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.detach().cpu()
        return int(torch.argmin((self.timesteps - timestep).abs()))

    def sigma_pair(self, timestep, to_final=False, self_corr=False):
        """(sigma, sigma_next) of the Euler step at `timestep` (reference :53-62)."""
        i = self._index(timestep)
        sigma = float(self.sigmas[i])
        if to_final or i + 1 >= len(self.timesteps):
            nxt = 1.0 if (self.inverse_timesteps or self.reverse_sigmas or self_corr) else 0.0
        else:
            nxt = float(self.sigmas[i + 1])
        return sigma, nxt

    def step(self, model_output, timestep, sample, to_final=False, **kwargs):
        sigma, nxt = self.sigma_pair(timestep, to_final, kwargs.get("self_corr", False))
        return sample + model_output * (nxt - sigma)

    def return_to_timestep(self, timestep, sample, sample_stablized):
        return (sample - sample_stablized) / float(self.sigmas[self._index(timestep)])

    def add_noise(self, original_samples, noise, timestep):
        sigma = float(self.sigmas[self._index(timestep)])
        return (1 - sigma) * original_samples + sigma * noise

    def training_target(self, sample, noise, timestep):
        return noise - sample

    def training_weight(self, timestep):
        return self.linear_timesteps_weights[self._index(timestep)]
