from . import preprocess
