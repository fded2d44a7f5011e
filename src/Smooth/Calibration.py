from gnomix_b200.calibration import Calibrator  # noqa: F401  (reference: src/Smooth/Calibration.py:19)
