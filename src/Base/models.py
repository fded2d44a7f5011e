from gnomix_b200.base import LogisticRegressionBase, CovRSKBase  # noqa: F401  (reference: src/Base/models.py:12,195)
