"""stein_b200 -- B200-native SVGD engine with the API of JamesBrofos/Stein.

    from stein_b200.samplers import SteinSampler
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.log_p import LogisticRegression

The same modules are importable as `stein.*` (see the `stein/` alias package at
the repository root) so code written against the reference keeps its imports.
"""
__version__ = "0.1.0"
