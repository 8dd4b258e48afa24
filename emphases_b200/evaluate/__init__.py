from .core import corpus, datasets
from . import metrics
from .metrics import Metrics
