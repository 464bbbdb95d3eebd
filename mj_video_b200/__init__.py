"""mj_video_b200 — B200-native (sm_100a) implementation of MJ-VIDEO's batched reward-scoring forward pass.

Drop-in names mirror the reference (aiming-lab/MJ-Video):
  mj_video_b200.model.InternVLChatRewardModeling / InternVLChatRewardModelingConfig / prepare_chat_input
      (reference: scripts/model/__init__.py:1-2)
  mj_video_b200.data_processor.load_video
      (reference: scripts/data_processor/__init__.py:1-2)
All compute goes through the C-ABI library libmjv_b200.so (include/mjv_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
