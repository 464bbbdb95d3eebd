"""Mirror of the reference's `scripts/model/__init__.py:1-2` exports."""
from .internvl2 import prepare_chat_input  # noqa: F401
from .moe_reward import (CustomOutput, GatingNetwork, InternVLChatRewardModeling,  # noqa: F401
                         InternVLChatRewardModelingConfig, find_token_for_gating, token_pattern)
