from .views import RandomResizedCropFlip, SingleImageDataset  # noqa: F401
