from gnomix_b200.smooth import XGB_Smoother, CRF_Smoother  # noqa: F401  (reference: src/Smooth/models.py:8,27)
