"""Prompt pre-processing base (reference diffsynth/prompters/base_prompter.py:39-70): refiner / extender hooks."""
import torch


class BasePrompter:
    def __init__(self):
        self.refiners = []
        self.extenders = []

    def load_prompt_refiners(self, model_manager, refiner_classes=()):
        for cls in refiner_classes:
            self.refiners.append(cls.from_model_manager(model_manager))

    def load_prompt_extenders(self, model_manager, extender_classes=()):
        for cls in extender_classes:
            self.extenders.append(cls.from_model_manager(model_manager))

    @torch.no_grad()
    def process_prompt(self, prompt, positive=True):
        if isinstance(prompt, list):
            return [self.process_prompt(p, positive=positive) for p in prompt]
        for refiner in self.refiners:
            prompt = refiner(prompt, positive=positive)
        return prompt

    @torch.no_grad()
    def extend_prompt(self, prompt, positive=True):
        out = dict(prompt=prompt)
        for extender in self.extenders:
            out = extender(out)
        return out
