"""Host-side image helpers the harness star-imports (reference utils/image_process.py:39-70,173-204).
Pure CPU / PIL code at the clip boundary; nothing here is on the GPU hot path."""
import os

import numpy as np
from PIL import Image


def tensor_to_pil(x):
    """[H,W,C] tensor in [0,1] (or uint8) -> PIL (reference :20-36)."""
    arr = x.cpu().numpy()
    if arr.dtype != np.uint8:
        arr = (np.clip(arr, 0, 1) * 255).astype(np.uint8)
    return Image.fromarray(arr)


def calculate_dimensions(image_input, max_width=640):
    """(height, width): width-limited, aspect preserved, floored to multiples of 16 (reference :39-70)."""
    img = image_input if isinstance(image_input, Image.Image) else Image.open(image_input)
    w0, h0 = img.size
    if w0 <= max_width:
        w, h = w0, h0
    else:
        w, h = max_width, int(max_width * (h0 / w0))
    return (h // 16) * 16, (w // 16) * 16


def find_reference_image(ref_image_root):
    """frame.jpg > frame.png > first jpg/jpeg > first png (reference :173-204)."""
    for name in ("frame.jpg", "frame.png"):
        p = os.path.join(ref_image_root, name)
        if os.path.exists(p):
            return p
    files = os.listdir(ref_image_root)
    for exts in ((".jpg", ".jpeg", ".JPG", ".JPEG"), (".png", ".PNG")):
        for ext in exts:
            for f in files:
                if f.endswith(ext):
                    return os.path.join(ref_image_root, f)
    raise FileNotFoundError(f"No reference image (jpg/png) found in {ref_image_root}")
