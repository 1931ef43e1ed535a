"""Prompt encoding: WanPrompter (tokenizer + native umT5 encoder), reference diffsynth/prompters/."""
from .base_prompter import BasePrompter
from .wan_prompter import HuggingfaceTokenizer, WanPrompter

__all__ = ["BasePrompter", "HuggingfaceTokenizer", "WanPrompter"]
