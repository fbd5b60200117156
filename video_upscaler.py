#!/usr/bin/env python
"""Drop-in entry point with the reference's name and flag surface (reference video_upscaler.py:629-759).
The per-frame restoration runs on hand-written sm_100a kernels (video_restore_b200/); see DESIGN.md."""
import sys

from video_restore_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
