from .views import RandomResizedCropFlip, SingleImageDataset  # noqa: F401
from .gpu_views import GpuViewGenerator, sample_view_params  # noqa: F401,E402
