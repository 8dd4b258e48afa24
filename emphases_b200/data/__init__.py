from . import preprocess
from .collate import collate
from .sampler import LengthDataset, Sampler, buckets, sampler
