from .abstract_stein_sampler import AbstractSteinSampler
from .stein_sampler import SteinSampler
