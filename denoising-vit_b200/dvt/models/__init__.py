from .vit_wrapper import *  # noqa: F401,F403
from .vit_wrapper import MODEL_LIST, PretrainedViTWrapper  # noqa: F401
