"""save_video / VideoData (reference diffsynth/data/video.py:138-143): pure host I/O at the clip boundary.

The reference writes through imageio+ffmpeg; imageio is not installed in this image, so the writer is chosen
at call time: imageio if importable, else OpenCV, else a directory of PNG frames (never silently dropped)."""
import os

import numpy as np
from PIL import Image


def save_frames(frames, save_path):
    os.makedirs(save_path, exist_ok=True)
    for i, fr in enumerate(frames):
        fr.save(os.path.join(save_path, f"{i}.png"))


def save_video(frames, save_path, fps, quality=9, ffmpeg_params=None):
    try:
        import imageio
        w = imageio.get_writer(save_path, fps=fps, quality=quality, ffmpeg_params=ffmpeg_params)
        for fr in frames:
            w.append_data(np.array(fr))
        w.close()
        return save_path
    except ImportError:
        pass
    try:
        import cv2
        h, w_ = np.array(frames[0]).shape[:2]
        vw = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w_, h))
        for fr in frames:
            vw.write(cv2.cvtColor(np.array(fr), cv2.COLOR_RGB2BGR))
        vw.release()
        return save_path
    except ImportError:
        out_dir = os.path.splitext(save_path)[0] + "_frames"
        print(f"save_video: no video encoder available (imageio / cv2 missing); writing PNG frames to {out_dir}")
        save_frames(frames, out_dir)
        return out_dir


class StreamingVideoWriter:
    """Appends clips to ONE output as they are produced instead of re-encoding everything generated so far after every clip
    (the reference harness calls save_video(all_frames) per clip, test_svi.py:483 — quadratic in the number of clips).
    Same writer choice as save_video: imageio, else OpenCV, else a directory of PNG frames.

        w = StreamingVideoWriter(path, fps=16); w.append(clip_frames); ...; w.close()
    """

    def __init__(self, save_path, fps, quality=9, ffmpeg_params=None):
        self.save_path, self.fps, self.count = save_path, fps, 0
        self._w, self._kind = None, None
        try:
            import imageio
            self._w, self._kind = imageio.get_writer(save_path, fps=fps, quality=quality, ffmpeg_params=ffmpeg_params), "imageio"
        except ImportError:
            try:
                import cv2  # noqa: F401
                self._kind = "cv2"
            except ImportError:
                self._kind = "png"
                self.save_path = os.path.splitext(save_path)[0] + "_frames"
                os.makedirs(self.save_path, exist_ok=True)

    def append(self, frames):
        for fr in frames:
            if self._kind == "imageio":
                self._w.append_data(np.array(fr))
            elif self._kind == "cv2":
                import cv2
                a = np.array(fr)
                if self._w is None:
                    self._w = cv2.VideoWriter(self.save_path, cv2.VideoWriter_fourcc(*"mp4v"), self.fps, (a.shape[1], a.shape[0]))
                self._w.write(cv2.cvtColor(a, cv2.COLOR_RGB2BGR))
            else:
                fr.save(os.path.join(self.save_path, f"{self.count}.png"))
            self.count += 1
        return self.count

    def close(self):
        if self._kind == "imageio" and self._w is not None:
            self._w.close()
        elif self._kind == "cv2" and self._w is not None:
            self._w.release()
        self._w = None
        return self.save_path


class VideoData:
    """Minimal image-folder / frame-list reader (reference data/video.py VideoData)."""

    def __init__(self, video_file=None, image_folder=None, height=None, width=None, **kwargs):
        if video_file is not None:
            raise NotImplementedError("video decoding needs imageio/ffmpeg, which is not part of the hot path")
        names = sorted(os.listdir(image_folder), key=lambda n: (len(n), n))
        self.files = [os.path.join(image_folder, n) for n in names if n.lower().endswith((".png", ".jpg", ".jpeg"))]
        self.height, self.width = height, width

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        im = Image.open(self.files[i]).convert("RGB")
        if self.height is not None and self.width is not None:
            im = im.resize((self.width, self.height))
        return im
