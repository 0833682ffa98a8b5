"""tf.contrib.distributions.{Normal, Gamma}.log_prob as graph expressions.

Reference call sites: examples/linear_regression/main.py:25,31;
examples/logistic_regression/main.py:33-37,47-48; examples/regression_neural_network/main.py:52-72.
Formulas are TensorFlow 1.12's (tensorflow/python/ops/distributions/{normal,gamma}.py):
  Normal(loc, scale).log_prob(x) = -0.5 ((x - loc)/scale)^2 - (0.5 log(2 pi) + log scale)
  Gamma(concentration a, rate b).log_prob(x) = (a - 1) log x - b x - (lgamma(a) - a log b)
"""
import math

import tensorflow as tf


class Normal:
    def __init__(self, loc, scale, validate_args=False, allow_nan_stats=True, name="Normal"):
        self.loc = tf.convert_to_tensor(loc)
        self.scale = tf.convert_to_tensor(scale)

    def log_prob(self, value, name="log_prob"):
        z = (tf.convert_to_tensor(value) - self.loc) / self.scale
        return -0.5 * tf.square(z) - (0.5 * math.log(2.0 * math.pi) + tf.log(self.scale))

    def prob(self, value, name="prob"):
        return tf.exp(self.log_prob(value))

    def mean(self, name="mean"):
        return self.loc * tf.ones_like(self.scale)

    def stddev(self, name="stddev"):
        return self.scale * tf.ones_like(self.loc)


class Gamma:
    def __init__(self, concentration, rate, validate_args=False, allow_nan_stats=True, name="Gamma"):
        self.concentration = tf.convert_to_tensor(concentration)
        self.rate = tf.convert_to_tensor(rate)

    def log_prob(self, value, name="log_prob"):
        x = tf.convert_to_tensor(value)
        a, b = self.concentration, self.rate
        return (a - 1.0) * tf.log(x) - b * x - (tf.lgamma(a) - a * tf.log(b))

    def prob(self, value, name="prob"):
        return tf.exp(self.log_prob(value))

    def mean(self, name="mean"):
        return self.concentration / self.rate
