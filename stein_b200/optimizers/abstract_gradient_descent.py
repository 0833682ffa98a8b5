from abc import abstractmethod


class AbstractGradientDescent:
    """Step rule applied to the SVGD direction phi.

    Mirrors stein/optimizers/abstract_gradient_descent.py:4-52: a global
    `learning_rate`, a `decay` factor and an iteration counter `n_iters`;
    subclasses implement `update(phi) -> step` (the step is ADDED to the
    particles, stein/samplers/abstract_stein_sampler.py:126).

    Subclasses defined by this package (Adam, Adagrad) run as one fused CUDA
    kernel inside the sampler; a user-defined subclass is honoured through the
    same `update(phi)` call on host arrays, exactly as in the reference.
    """

    def __init__(self, learning_rate, decay):
        self.learning_rate = learning_rate
        self.decay = decay
        self.n_iters = 0

    @abstractmethod
    def update(self, phi):
        raise NotImplementedError()
