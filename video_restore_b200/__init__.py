"""B200-native per-frame restoration hot path of video-restore."""
