"""GPU image transforms of the MingTok path (see processor.py): the reference's `mingtok.utils` exports
`CenterCropProcessor`; the no-crop variant and the tensor -> PIL conversion live here too."""
from .processor import CenterCropProcessor, ResizeProcessor, tensor_to_pil

__all__ = ["CenterCropProcessor", "ResizeProcessor", "tensor_to_pil"]
