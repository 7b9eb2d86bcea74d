"""Mirror of the reference's `CLIP.clip` surface used by the hot path (CLIP/clip/__init__.py)."""
from .model import available_models  # noqa: F401
from .tokenizer import tokenize  # noqa: F401
from .wrapper import ClipGradcam, ClipWrapper, saliency_configs  # noqa: F401
