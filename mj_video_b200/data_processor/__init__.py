"""Mirror of the reference's `scripts/data_processor/__init__.py` exports that lie on the scoring path."""
from .data import (ClipDecodePool, find_closest_aspect_ratio, get_index, load_video, open_video,  # noqa: F401
                   preprocess_frames, sample_clip)
