from .neural_feature_field import *  # noqa: F401,F403
from .offline_denoiser import *  # noqa: F401,F403
from .vit_wrapper import *  # noqa: F401,F403
from .neural_feature_field import NeuralFeatureField  # noqa: F401
from .offline_denoiser import SingleImageDenoiser  # noqa: F401
from .vit_wrapper import MODEL_LIST, PretrainedViTWrapper  # noqa: F401
from .online_denoiser import CenterPadding, Denoiser  # noqa: F401,E402
