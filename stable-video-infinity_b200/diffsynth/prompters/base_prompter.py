"""Hook points of prompt pre-processing (reference diffsynth/prompters/base_prompter.py:39-70).

A prompter owns two ordered hook lists.  *Refiners* are callables `(text, positive=bool) -> text` applied to every prompt
string (lists of prompts are handled element-wise); *extenders* are callables `dict -> dict` that may add keys next to
`prompt`.  Both are built from a ModelManager through their own `from_model_manager`.  The SVI pipelines use neither by
default, so for them `process_prompt` is the identity.
"""


class BasePrompter:
    def __init__(self):
        self.refiners, self.extenders = [], []

    # -- registration -------------------------------------------------------------------------------
    @staticmethod
    def _build(model_manager, classes):
        return [cls.from_model_manager(model_manager) for cls in classes]

    def load_prompt_refiners(self, model_manager, refiner_classes=()):
        self.refiners += self._build(model_manager, refiner_classes)

    def load_prompt_extenders(self, model_manager, extender_classes=()):
        self.extenders += self._build(model_manager, extender_classes)

    # -- application --------------------------------------------------------------------------------
    def process_prompt(self, prompt, positive=True):
        if isinstance(prompt, (list, tuple)):
            return [self.process_prompt(one, positive=positive) for one in prompt]
        text = prompt
        for hook in self.refiners:
            text = hook(text, positive=positive)
        return text

    def extend_prompt(self, prompt, positive=True):
        bundle = {"prompt": prompt}
        for hook in self.extenders:
            bundle = hook(bundle)
        return bundle
