"""Prompt templates used by `prepare_chat_input` (reference: scripts/model/internvl2/conversation.py:240-249, 330-391).
Only the MPT-style templates the InternVL2 family uses are provided; the prompt is `system + sep + (role + message +
sep)*`, with the last (assistant) role left open."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

_SYS_ZH = "你是由上海人工智能实验室联合商汤科技开发的书生多模态大模型，英文名叫InternVL, 是一个有用无害的人工智能助手。"
_SYS_25 = "你是书生·万象，英文名是InternVL，是由上海人工智能实验室、清华大学及多家合作单位联合开发的多模态大语言模型。"


@dataclass
class Conversation:
    name: str
    system_template: str = "{system_message}"
    system_message: str = ""
    roles: tuple = ("USER", "ASSISTANT")
    messages: list = field(default_factory=list)
    sep: str = "\n"
    stop_str: str | None = None

    def get_prompt(self) -> str:
        ret = self.system_template.format(system_message=self.system_message) + self.sep
        for role, message in self.messages:
            if message:
                if type(message) is tuple:
                    message = message[0]
                ret += role + message + self.sep
            else:
                ret += role
        return ret

    def set_system_message(self, system_message: str):
        self.system_message = system_message

    def append_message(self, role: str, message):
        self.messages.append([role, message])

    def copy(self) -> "Conversation":
        return copy.deepcopy(self)


conv_templates: dict[str, Conversation] = {}


def register_conv_template(template: Conversation, override: bool = False):
    if not override:
        assert template.name not in conv_templates, f"{template.name} has been registered."
    conv_templates[template.name] = template


def get_conv_template(name: str) -> Conversation:
    return conv_templates[name].copy()


register_conv_template(Conversation(name="Hermes-2", system_template="<|im_start|>system\n{system_message}",
                                    system_message=_SYS_ZH, roles=("<|im_start|>user\n", "<|im_start|>assistant\n"),
                                    sep="<|im_end|>", stop_str="<|endoftext|>"))
register_conv_template(Conversation(name="internlm2-chat", system_template="<|im_start|>system\n{system_message}",
                                    system_message=_SYS_ZH, roles=("<|im_start|>user\n", "<|im_start|>assistant\n"),
                                    sep="<|im_end|>"))
register_conv_template(Conversation(name="phi3-chat", system_template="<|system|>\n{system_message}",
                                    system_message=_SYS_ZH, roles=("<|user|>\n", "<|assistant|>\n"), sep="<|end|>"))
register_conv_template(Conversation(name="internvl2_5", system_template="<|im_start|>system\n{system_message}",
                                    system_message=_SYS_25, roles=("<|im_start|>user\n", "<|im_start|>assistant\n"),
                                    sep="<|im_end|>\n"))
