from .base import LogPosterior, Output, Placeholder, Variable
from .linear_regression import LinearRegression
from .logistic_regression import LogisticRegression
from .regression_neural_network import RegressionNeuralNetwork
from .torch_log_p import TorchLogPosterior
from .graph_log_p import GraphLogPosterior, is_graph_tensor
from .gaussian import GaussianMixtureTarget
