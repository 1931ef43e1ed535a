"""Prompt encoding (umT5-XXL) is outside the hot-path scope; see pipelines.svi_video.encode_prompt."""
