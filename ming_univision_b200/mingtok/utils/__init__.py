from .processor import CenterCropProcessor, ResizeProcessor, tensor_to_pil  # noqa: F401
