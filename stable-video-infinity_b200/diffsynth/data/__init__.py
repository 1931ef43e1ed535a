from .video import VideoData, save_frames, save_video
