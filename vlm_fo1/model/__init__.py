from .fo1_model import Fo1ForCausalLM as OmChatQwen25VLForCausalLM  # the name callers of the reference import  # noqa: F401
from .fo1_model import Fo1ForCausalLM  # noqa: F401
