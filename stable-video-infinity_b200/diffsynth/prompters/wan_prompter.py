"""WanPrompter — prompt string -> umT5 embedding [1, 512, 4096] (reference diffsynth/prompters/wan_prompter.py).

Tokenisation is the HuggingFace umT5 tokenizer found next to the text-encoder checkpoint
(`<dirname(T5 path)>/google/umt5-xxl`, reference svi_video.py:249) with `padding='max_length'`, `truncation=True`,
`max_length=512`, special tokens added and 'whitespace' cleaning (:93-103); the encoder output past the prompt length
is zero-filled (:106-108).  The encoder itself is `models.wan_video_text_encoder.WanTextEncoder` on the native kernels.
"""
import html
import re

import torch

from .base_prompter import BasePrompter

try:                      # the reference imports ftfy unconditionally; it only matters for mojibake in prompts
    import ftfy
    _fix_text = ftfy.fix_text
except ImportError:       # not in this image: plain prompts are unaffected
    _fix_text = lambda s: s


def clean_prompt(text, lower=False):
    """'whitespace' / 'lower' cleaning of the reference tokenizer wrapper (:12-21, :74-81): repair mojibake (ftfy, when
    installed), undo doubly escaped HTML entities, collapse runs of whitespace."""
    text = html.unescape(html.unescape(_fix_text(text)))
    text = re.sub(r"\s+", " ", text.strip()).strip()
    return text.lower() if lower else text


class HuggingfaceTokenizer:
    """Fixed-length front end of a HuggingFace tokenizer directory (reference :35-82): prompts are cleaned, padded /
    truncated to `seq_len` tokens, and returned as int64 ids (and the attention mask on request)."""

    CLEAN_MODES = (None, "whitespace", "lower")

    def __init__(self, name, seq_len=None, clean=None, **kwargs):
        if clean not in self.CLEAN_MODES:
            raise ValueError(f"clean must be one of {self.CLEAN_MODES}, got {clean!r}")
        from transformers import AutoTokenizer
        self.name, self.seq_len, self.clean = name, seq_len, clean
        self.tokenizer = AutoTokenizer.from_pretrained(name, **kwargs)
        self.vocab_size = self.tokenizer.vocab_size

    def __call__(self, sequence, return_mask=False, **tokenizer_kwargs):
        texts = [sequence] if isinstance(sequence, str) else list(sequence)
        if self.clean is not None:
            texts = [clean_prompt(t, lower=self.clean == "lower") for t in texts]
        options = dict(return_tensors="pt")
        if self.seq_len is not None:
            options.update(padding="max_length", truncation=True, max_length=self.seq_len)
        options.update(tokenizer_kwargs)
        batch = self.tokenizer(texts, **options)
        return (batch.input_ids, batch.attention_mask) if return_mask else batch.input_ids


class WanPrompter(BasePrompter):
    def __init__(self, tokenizer_path=None, text_len=512):
        super().__init__()
        self.text_len = text_len
        self.text_encoder = None
        self.tokenizer = None
        self.fetch_tokenizer(tokenizer_path)

    def fetch_tokenizer(self, tokenizer_path=None):
        if tokenizer_path is not None:
            self.tokenizer = HuggingfaceTokenizer(name=tokenizer_path, seq_len=self.text_len, clean="whitespace")

    def fetch_models(self, text_encoder=None):
        self.text_encoder = text_encoder

    def encode_ids(self, ids, mask, device="cuda"):
        """ids / mask int64 [B, text_len] -> [B, text_len, dim]; rows past each prompt's length are zero (:104-108)."""
        outs = []
        for i in range(ids.shape[0]):
            emb = self.text_encoder(ids[i:i + 1].to(device), mask[i:i + 1].to(device))
            n = int(mask[i].gt(0).sum())
            emb[:, n:] = 0
            outs.append(emb)
        return torch.cat(outs, dim=0)

    def encode_prompt(self, prompt, positive=True, device="cuda"):
        if self.tokenizer is None or self.text_encoder is None:
            raise RuntimeError("svi_b200: WanPrompter needs a tokenizer directory (<T5 dir>/google/umt5-xxl) and a text encoder")
        prompt = self.process_prompt(prompt, positive=positive)
        ids, mask = self.tokenizer(prompt, return_mask=True, add_special_tokens=True)
        return self.encode_ids(ids, mask, device)

    def __call__(self, prompt, positive=True):
        return self.encode_prompt(prompt, positive=positive, device=next(self.text_encoder.parameters()).device)
