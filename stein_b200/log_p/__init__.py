from .base import LogPosterior, Output, Placeholder, Variable
from .linear_regression import LinearRegression
from .logistic_regression import LogisticRegression
from .regression_neural_network import RegressionNeuralNetwork
from .torch_log_p import TorchLogPosterior
from .gaussian import GaussianMixtureTarget
