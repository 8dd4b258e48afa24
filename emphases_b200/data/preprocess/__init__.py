from .core import from_audio
from . import mels
