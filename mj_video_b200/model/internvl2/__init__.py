from .configuration import InternLM2Config, InternVisionConfig, InternVLChatConfig, InternVLChatRewardModelingConfig  # noqa: F401
from .conversation import get_conv_template  # noqa: F401
from .modeling_internvl_chat import InternVLChatModel, prepare_chat_input  # noqa: F401
