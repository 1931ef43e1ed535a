from .video import StreamingVideoWriter, VideoData, save_frames, save_video
