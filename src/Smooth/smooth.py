from gnomix_b200.smooth import Smoother  # noqa: F401  (reference: src/Smooth/smooth.py:7)
